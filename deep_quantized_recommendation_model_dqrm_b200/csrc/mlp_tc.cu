// (a15) QuantLinear contractions on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
// Reference: QuantLinear.forward, quantization_supp/quant_modules_not_quantize_grad.py:209 (F.linear on the
//            integer-valued weights) and its autograd (addmm backward + the STE of SymmetricQuantFunction, qu:363).
//
// Same three fused layer kernels as mlp_fused.cu (fwd / dx / dW+db with their prologues and epilogues), but the
// contraction runs as tcgen05.mma instead of FFMA.  fp32 parity (1e-5 against the fp32 reference) is kept by
// splitting every fp32 operand into TF32 terms inside the kernel while it is staged into shared memory:
//     v = hi + lo (+ 2^-22 |v|),  hi = tf32(v),  lo = tf32(v - hi)
//   * the fake-quantised weights W_int are integers in [-2^(b-1), 2^(b-1)-1]: EXACT in TF32, one term;
//   * fwd  x W_int^t   and  dx  g W_int : (hi + lo) x W        -> 2 MMAs per K-step   ("2xTF32")
//   * dW   g^t x                        : hi.hi + hi.lo + lo.hi -> 3 MMAs per K-step   ("3xTF32")
// All products are exact in the tensor core (11 x 11 significant bits) and accumulate in fp32 in TMEM.
//
// One CTA = one 128 x BN output tile.  Operands are NOT loaded by TMA: every element has to pass through
// registers anyway (the TF32 split; for dx / dW the operand g = dout * act'(out) * s_row is generated on the fly),
// so 256 threads load, split, and store straight into the UMMA canonical no-swizzle K-major core-matrix layout
// (kind::tf32 takes MN-major operands only in the 128B_BASE32B swizzle, so the operands whose K is not contiguous
// in global memory -- W_int in dx, g and x in dW -- are transposed by the store pattern).  Two stages; thread 0
// issues the MMAs of a stage and commits them to that stage's mbarrier, which is what frees the stage for
// re-use (register prefetch of the next K-chunk overlaps the MMAs).  The epilogue reads the accumulators with
// tcgen05.ld (thread = tile row) and applies bias / per-row scale / activation / STE division.
// Small batches leave few tiles: K is split across a thread-block cluster (<= 8 CTAs) and the partial tiles are
// reduced through distributed shared memory in fixed rank order -- deterministic, no atomics, no workspace.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "common.cuh"

namespace dqrm {
namespace tc {

constexpr int BM = 128, BK = 32, kThreads = 256;
constexpr int A_TILE_BYTES = BM * BK * 4;                       // 16 KiB per TF32 term
// Pipeline shape: NS shared-memory stages, P K-chunks of global loads in flight per thread (register sets).  The
// kernels are bound by L2 -> SM latency (one chunk per CTA in flight gave 2 us per chunk = 2.4 TB/s over the chip),
// so P is what sets the throughput; the stage count only has to cover the MMAs that still read older stages.
__host__ __device__ constexpr int stage_bytes(int mode, int bn) { return 2 * A_TILE_BYTES + (mode == 2 ? 2 : 1) * bn * BK * 4; }
__host__ __device__ constexpr int num_stages(int mode, int bn) { return 3; }
__host__ __device__ constexpr int prefetch_depth(int mode, int bn) { return mode == 2 ? 2 : 3; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// A commit that never arrives must abort the kernel (sticky launch error on the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" 8 x 16 B core matrices), version 1 (sm_100):
// start address, leading-dimension byte offset, stride-dimension byte offset, each >> 4.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::tf32 instruction descriptor: D = f32, A = B = tf32, majors, N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  lo = make_float4(to_tf32(__fsub_rn(v.x, hi.x)), to_tf32(__fsub_rn(v.y, hi.y)), to_tf32(__fsub_rn(v.z, hi.z)),
                   to_tf32(__fsub_rn(v.w, hi.w)));
}

// Global loads as PREDICATED instructions with the zero default written before them (inline asm, volatile): no
// branch and no merge of two code paths follows a load, so nothing consumes a loaded register until the chunk is
// stored -- the loads of P chunks really are in flight together -- and the compiler cannot sink them towards their
// use.  (A C++ "in range ? load : 0" was compiled into branches with register moves right behind each load: every
// load's latency was exposed, measured as 50 % long-scoreboard stalls.)  L1::no_allocate: with ~150-190 KB of the
// SM's 256 KB configured as shared memory the L1 is a few tens of KB, and allocating loads can only be in flight
// for as many lines as it has (measured: 1.45 TB/s over the chip, 12 long-scoreboard stalls per issue); every
// sector is consumed by the one instruction that loads it, so there is nothing to cache.
__device__ __forceinline__ float ldg_pred(const float* p, bool ok) {
  float v;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.b32 %0, 0;\n\t@q ld.global.nc.L1::no_allocate.f32 %0, [%1];\n\t}"
               : "=f"(v) : "l"(p), "r"((int)ok));
  return v;
}
__device__ __forceinline__ float4 ldg4_pred(const float* p, bool ok) {
  float4 v;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
               "@q ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "r"((int)ok));
  return v;
}
// 4 consecutive elements of row `row` starting at column `col` (col % 4 == 0) of a row-major [nrows, ncols] matrix;
// out-of-range elements read as 0.  VEC: the matrix allows 16-byte loads (ld % 4 == 0, base 16-byte aligned; then
// ncols -- K or a K-slice end, multiples of 4 -- never cuts a vector).
template <bool VEC>
__device__ __forceinline__ float4 ld4(const float* __restrict__ p, int row, int col, int nrows, int ncols, int ld) {
  const float* q = p + (long long)row * ld + col;
  const bool in = row < nrows;
  if constexpr (VEC) {
    return ldg4_pred(q, in && col < ncols);
  } else {
    return make_float4(ldg_pred(q, in && col < ncols), ldg_pred(q + 1, in && col + 1 < ncols),
                       ldg_pred(q + 2, in && col + 2 < ncols), ldg_pred(q + 3, in && col + 3 < ncols));
  }
}

__device__ __forceinline__ float act_bwd(float dout, float out, int act) {
  if (act == 1) return out > 0.0f ? dout : 0.0f;                        // threshold_backward
  if (act == 2) return dout * ((1.0f - out) * out);                     // sigmoid_backward
  return dout;
}
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == 1) return fmaxf(z, 0.0f);
  if (act == 2) return 1.0f / (1.0f + expf(-z));
  return z;
}

// MODE 0 (fwd): C[b, n]  = act((sum_k x[b,k] W[n,k] + b_int[n]) * s[n])     M = batch, N = out, K = in
// MODE 1 (dx) : C[b, i]  = sum_o g[b,o] W[o,i],  g = dout * act'(out) * s    M = batch, N = in,  K = out
// MODE 2 (dW) : C[o, i] (+)= (sum_b g[b,o] x[b,i]) / s[o] ; db[o] (+)= (sum_b g[b,o]) / s[o]
//                                                                            M = out,   N = in,  K = batch
template <int MODE, int BN, bool CL, bool VEC>
__global__ void __launch_bounds__(kThreads)
linear_tc_kernel(const float* __restrict__ x, const float* __restrict__ W_int, const float* __restrict__ b_int,
                 const float* __restrict__ s_row, const float* __restrict__ dout, const float* __restrict__ out,
                 float* __restrict__ C, float* __restrict__ db, int batch, int out_f, int in_f, int act, int kc,
                 int accumulate) {
  namespace cg = cooperative_groups;
  constexpr int B_TILE_BYTES = BN * BK * 4;
  constexpr int B_TERMS = MODE == 2 ? 2 : 1;
  constexpr int STAGE_BYTES = stage_bytes(MODE, BN);
  constexpr int NS = num_stages(MODE, BN), P = prefetch_depth(MODE, BN);
  static_assert(STAGE_BYTES == 2 * A_TILE_BYTES + B_TERMS * B_TILE_BYTES, "stage layout");
  constexpr int RED_LD = BN + 1;                                       // padded partial-tile row (bank-conflict free)
  static_assert(BM * RED_LD * 4 <= 2 * stage_bytes(MODE, BN), "partial tile must fit in the stage buffers");
  constexpr uint32_t IDESC = umma_idesc(BM, BN, 0, 0);                 // both operands K-major

  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_stage[NS];
  __shared__ uint32_t s_tmem;
  __shared__ float db_s[4][BM];
  __shared__ float db_cta[BM];

  int S = 1, rank = 0;
  if constexpr (CL) {
    cg::cluster_group cluster = cg::this_cluster();
    S = (int)cluster.num_blocks();
    rank = (int)cluster.block_rank();
  }
  const int M = MODE == 2 ? out_f : batch;
  const int N = MODE == 0 ? out_f : in_f;
  const int K = MODE == 0 ? in_f : (MODE == 1 ? out_f : batch);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kbeg = rank * kc, kend = min(K, kbeg + kc);
  const int nchunks = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(&bar_stage[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = s_tmem;

  // ---- operand staging -------------------------------------------------------------------------------------------
  // Every operand is staged K-major (the only no-swizzle layout kind::tf32 accepts for either major-ness is K-major;
  // MN-major TF32 needs the 128B_BASE32B swizzle), i.e. as [K/4][rows][4 floats]: a 128-byte core matrix = 8 rows x
  // 4 consecutive k.
  //  * operand already K-contiguous in global memory (x and W_int in fwd, g in dx): "V" staging -- 16-byte slots,
  //    8 consecutive lanes = 8 rows of one K-chunk: a warp reads 8 global rows x 64 B and writes 4 whole core
  //    matrices with conflict-free 16-byte stores.  Slot i of a thread: row r0 + 32 i, K-chunk kq.
  //  * operand contiguous along M/N in global memory (W_int in dx, g and x in dW, whose K is the batch): "T" staging
  //    -- scalar slots, a warp covers 8 consecutive m x 4 consecutive k: 4 fully used 32-byte sectors per load
  //    instruction, and the transposing 4-byte shared stores hit 32 distinct banks.  Slot i of a thread in a tile
  //    of R rows: m = mt + 64 (i & 1), k = 4 (i >> 1) + k4 for R = 128;  m = mt, k = 4 i + k4 for R = 64.
  // The kernels are instruction-issue bound in this staging code (first version: 2000 instructions per warp and
  // chunk), so everything that does not depend on the chunk -- row / column predicates, base pointers, shared-memory
  // offsets (compile-time multiples of the slot index), s_row factors -- is computed once, before the K loop.
  constexpr int NAV = BM * (BK / 4) / kThreads;                        // V slots per thread: A (4)
  constexpr int NBV = BN * (BK / 4) / kThreads;                        //                     B (BN / 32)
  constexpr int NAT = BM * BK / kThreads;                              // T slots per thread: A (16)
  constexpr int NBT = BN * BK / kThreads;                              //                     B (BN / 8)
  struct RegSet {                                                      // one K-chunk of this thread's global loads
    float4 va0[MODE == 2 ? 1 : NAV], va1[MODE == 1 ? NAV : 1], vb[MODE == 0 ? NBV : 1];
    float ta0[MODE == 2 ? NAT : 1], ta1[MODE == 2 ? NAT : 1], tb[MODE == 0 ? 1 : NBT];
  };
  RegSet rs[P];
  const bool do_db = MODE == 2 && db != nullptr && blockIdx.x == 0;
  float db_part[2] = {0.f, 0.f};
  const int lane_m8 = lane & 7, lane_k4 = lane >> 3;
  const int v_r0 = (tid & 7) + ((tid >> 6) << 3), v_kq = (tid >> 3) & 7;  // V: row of slot 0, K-chunk of every slot
  const int t_m = warp * 8 + lane_m8;                                    // T: row (m or n) of slot 0
  const int ldx = in_f, ldg = out_f;
  // A operand
  const float* gA0 = nullptr; const float* gA1 = nullptr;               // chunk-0 addresses of slot 0
  unsigned a_ok = 0;                                                     // bit i: slot row i (V) / column half i (T) in range
  float a_s[2] = {0.f, 0.f};                                             // MODE 2: s_row of this thread's two channels
  if (MODE == 0) {
    gA0 = x + (long long)(m0 + v_r0) * ldx + kbeg + v_kq * 4;
#pragma unroll
    for (int i = 0; i < NAV; ++i) a_ok |= (unsigned)(m0 + v_r0 + 32 * i < M) << i;
  } else if (MODE == 1) {
    gA0 = dout + (long long)(m0 + v_r0) * ldg + kbeg + v_kq * 4;
    gA1 = out + (long long)(m0 + v_r0) * ldg + kbeg + v_kq * 4;
#pragma unroll
    for (int i = 0; i < NAV; ++i) a_ok |= (unsigned)(m0 + v_r0 + 32 * i < M) << i;
  } else {
    gA0 = dout + (long long)(kbeg + lane_k4) * ldg + m0 + t_m;
    gA1 = out + (long long)(kbeg + lane_k4) * ldg + m0 + t_m;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const bool ok = m0 + t_m + 64 * hh < M;
      a_ok |= (unsigned)ok << hh;
      a_s[hh] = ok ? __ldg(s_row + m0 + t_m + 64 * hh) : 0.f;
    }
  }
  // B operand
  const float* gB0 = nullptr;
  unsigned b_ok = 0;
  if (MODE == 0) {
    gB0 = W_int + (long long)(n0 + v_r0) * ldx + kbeg + v_kq * 4;
#pragma unroll
    for (int i = 0; i < NBV; ++i) b_ok |= (unsigned)(n0 + v_r0 + 32 * i < N) << i;
  } else {
    gB0 = (MODE == 1 ? W_int : x) + (long long)(kbeg + lane_k4) * ldx + n0 + t_m;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) b_ok |= (unsigned)(n0 + t_m + 64 * hh < N) << hh;
  }
  // T slot i of a tile of R rows: column half and K-quad
  auto t_half = [](int i, int R) { return R == 128 ? (i & 1) : 0; };
  auto t_quad = [](int i, int R) { return R == 128 ? (i >> 1) : i; };

  // loads of the chunk that starts `kofs` elements after kbeg (kofs + kbeg = k0)
  auto load_chunk = [&](RegSet& R, int kofs) {
    const int k0 = kbeg + kofs;
    if constexpr (MODE != 2) {
      const bool kok = k0 + v_kq * 4 < kend;                            // (VEC: K and the slices are multiples of 4)
#pragma unroll
      for (int i = 0; i < NAV; ++i) {
        const bool ok = ((a_ok >> i) & 1u) && kok;
        const long long o = (long long)(32 * i) * (MODE == 0 ? ldx : ldg) + kofs;
        if constexpr (VEC) {
          R.va0[i] = ldg4_pred(gA0 + o, ok);
          if (MODE == 1) R.va1[i] = ldg4_pred(gA1 + o, ok);
        } else {
          const int kk = k0 + v_kq * 4;
          R.va0[i] = make_float4(ldg_pred(gA0 + o, ok), ldg_pred(gA0 + o + 1, ok && kk + 1 < kend),
                                 ldg_pred(gA0 + o + 2, ok && kk + 2 < kend), ldg_pred(gA0 + o + 3, ok && kk + 3 < kend));
          if (MODE == 1)
            R.va1[i] = make_float4(ldg_pred(gA1 + o, ok), ldg_pred(gA1 + o + 1, ok && kk + 1 < kend),
                                   ldg_pred(gA1 + o + 2, ok && kk + 2 < kend), ldg_pred(gA1 + o + 3, ok && kk + 3 < kend));
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NAT; ++i) {                                    // g[b = k, o = m]
        const int hh = t_half(i, BM), qd = t_quad(i, BM);
        const bool ok = ((a_ok >> hh) & 1u) && (k0 + qd * 4 + lane_k4 < kend);
        const long long o = (long long)(kofs + qd * 4) * ldg + 64 * hh;
        R.ta0[i] = ldg_pred(gA0 + o, ok);
        R.ta1[i] = ldg_pred(gA1 + o, ok);
      }
    }
    if constexpr (MODE == 0) {
      const bool kok = k0 + v_kq * 4 < kend;
#pragma unroll
      for (int i = 0; i < NBV; ++i) {
        const bool ok = ((b_ok >> i) & 1u) && kok;
        const long long o = (long long)(32 * i) * ldx + kofs;
        if constexpr (VEC) {
          R.vb[i] = ldg4_pred(gB0 + o, ok);
        } else {
          const int kk = k0 + v_kq * 4;
          R.vb[i] = make_float4(ldg_pred(gB0 + o, ok), ldg_pred(gB0 + o + 1, ok && kk + 1 < kend),
                                ldg_pred(gB0 + o + 2, ok && kk + 2 < kend), ldg_pred(gB0 + o + 3, ok && kk + 3 < kend));
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NBT; ++i) {                                    // W_int[o = k, i = n] / x[b = k, i = n]
        const int hh = t_half(i, BN), qd = t_quad(i, BN);
        const bool ok = ((b_ok >> hh) & 1u) && (k0 + qd * 4 + lane_k4 < kend);
        R.tb[i] = ldg_pred(gB0 + (long long)(kofs + qd * 4) * ldx + 64 * hh, ok);
      }
    }
  };

  // TF32 terms by masking (the tensor core reads the top 19 bits): hi = v & ~0x1fff, lo = (v - hi) & ~0x1fff.
  // Truncation instead of cvt.rna: 3 full-rate instructions per element; the dropped remainder is < 2^-21 |v|.
  auto tf32_hi = [](float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); };
  const uint32_t sA_v = (uint32_t)(v_kq * (BM * 16) + v_r0 * 16);        // V slot 0 offsets inside a tile (slot i: + 512 i)
  const uint32_t sB_v = (uint32_t)(v_kq * (BN * 16) + v_r0 * 16);
  const uint32_t sT = (uint32_t)(t_m * 16 + lane_k4 * 4);                // T slot 0 (slot i: + quad * R*16 + half * 1024)

  auto store_chunk = [&](const RegSet& R, int stage, int kofs) {
    unsigned char* a_hi = smem + stage * STAGE_BYTES;
    unsigned char* a_lo = a_hi + A_TILE_BYTES;
    unsigned char* b_hi = a_lo + A_TILE_BYTES;
    unsigned char* b_lo = b_hi + B_TILE_BYTES;
    if constexpr (MODE != 2) {
      float4 sv = make_float4(1.f, 1.f, 1.f, 1.f);
      if (MODE == 1) {                                                   // s_row[o] of this thread's K-chunk (0 beyond K)
        const int o = kbeg + kofs + v_kq * 4;
        sv.x = o + 0 < kend ? __ldg(s_row + o + 0) : 0.f; sv.y = o + 1 < kend ? __ldg(s_row + o + 1) : 0.f;
        sv.z = o + 2 < kend ? __ldg(s_row + o + 2) : 0.f; sv.w = o + 3 < kend ? __ldg(s_row + o + 3) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < NAV; ++i) {
        float4 v = R.va0[i];
        if (MODE == 1) {                                                 // g = dout * act'(out) * s_row[o]
          v.x = __fmul_rn(act_bwd(v.x, R.va1[i].x, act), sv.x); v.y = __fmul_rn(act_bwd(v.y, R.va1[i].y, act), sv.y);
          v.z = __fmul_rn(act_bwd(v.z, R.va1[i].z, act), sv.z); v.w = __fmul_rn(act_bwd(v.w, R.va1[i].w, act), sv.w);
        }
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 lo = make_float4(__fsub_rn(v.x, hi.x), __fsub_rn(v.y, hi.y), __fsub_rn(v.z, hi.z), __fsub_rn(v.w, hi.w));
        *reinterpret_cast<float4*>(a_hi + sA_v + 512 * i) = hi;
        *reinterpret_cast<float4*>(a_lo + sA_v + 512 * i) = lo;          // (the MMA ignores lo's low 13 bits itself)
      }
    } else {
#pragma unroll
      for (int i = 0; i < NAT; ++i) {
        const int hh = t_half(i, BM), qd = t_quad(i, BM);
        const float g = __fmul_rn(act_bwd(R.ta0[i], R.ta1[i], act), a_s[hh]);
        if (do_db) db_part[hh] = __fadd_rn(db_part[hh], g);              // column sums of g
        const float hi = tf32_hi(g);
        *reinterpret_cast<float*>(a_hi + sT + qd * (BM * 16) + hh * 1024) = hi;
        *reinterpret_cast<float*>(a_lo + sT + qd * (BM * 16) + hh * 1024) = __fsub_rn(g, hi);
      }
    }
    if constexpr (MODE == 0) {
#pragma unroll
      for (int i = 0; i < NBV; ++i) *reinterpret_cast<float4*>(b_hi + sB_v + 512 * i) = R.vb[i];   // small integers: exact
    } else {
#pragma unroll
      for (int i = 0; i < NBT; ++i) {
        const int hh = t_half(i, BN), qd = t_quad(i, BN);
        const uint32_t off = sT + qd * (BN * 16) + hh * 1024;
        if (MODE == 1) {
          *reinterpret_cast<float*>(b_hi + off) = R.tb[i];
        } else {
          const float hi = tf32_hi(R.tb[i]);
          *reinterpret_cast<float*>(b_hi + off) = hi;
          *reinterpret_cast<float*>(b_lo + off) = __fsub_rn(R.tb[i], hi);
        }
      }
    }
  };

  // ---- main loop: NS stages, the commit of a stage's MMAs frees it ------------------------------------------------
  // The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of
  // accumulations (measured: 6e-5 of the largest output after K = 8192).  So an accumulator only ever collects
  // kFlush K-chunks (K = 128: <= 48 MMAs); then every thread moves its part of the tile into fp32 REGISTER
  // accumulators with round-to-nearest adds and the next group starts a fresh TMEM accumulator.
  constexpr int kFlush = 4;
  const int q = warp & 3, h = warp >> 2;                                 // TMEM lane quarter, column half
  const int row = q * 32 + lane;                                         // tile row owned by this thread
  constexpr int CW = BN / 2;                                             // columns per warp
  float acc[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) acc[i] = 0.0f;
#pragma unroll
  for (int p = 0; p < P; ++p)
    if (p < nchunks) load_chunk(rs[p], p * BK);
  for (int c0 = 0; c0 < nchunks; c0 += P) {
#pragma unroll
    for (int p = 0; p < P; ++p) {                                        // register set p <-> chunks c0 + p (static index)
      const int c = c0 + p;
      if (c < nchunks) {                                                 // CTA-uniform
        const int st = c % NS, use = c / NS;
        if (use > 0) mbar_wait(&bar_stage[st], (use - 1) & 1);           // MMAs of chunk c-NS have read this stage
        store_chunk(rs[p], st, c * BK);
        fence_proxy_async();                                             // generic-proxy stores -> visible to the MMA
        __syncthreads();                                                 // (also: every thread's flush loads are done)
        if (c + P < nchunks) load_chunk(rs[p], (c + P) * BK);     // P chunks in flight while the MMAs run
        if (tid == 0) {
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + st * STAGE_BYTES), a_lo = a_hi + A_TILE_BYTES;
          const uint32_t b_hi = a_lo + A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
          // K-major tile of R rows: LBO (K-chunk stride) = R*16, SBO (8-row group stride) = 128; one K=8 step = 2
          // chunks, so the next K-step starts R*32 bytes further
          constexpr uint32_t A_LBO = BM * 16, B_LBO = BN * 16;
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t dah = umma_desc(a_hi + j * BM * 32, A_LBO, 128), dal = umma_desc(a_lo + j * BM * 32, A_LBO, 128);
            const uint64_t dbh = umma_desc(b_hi + j * BN * 32, B_LBO, 128);
            umma_tf32(tmem_d, dah, dbh, IDESC, ((c % kFlush) | j) != 0); // first MMA of a group overwrites
            umma_tf32(tmem_d, dal, dbh, IDESC, 1);
            if (MODE == 2) umma_tf32(tmem_d, dah, umma_desc(b_lo + j * BN * 32, B_LBO, 128), IDESC, 1);
          }
          umma_commit(&bar_stage[st]);
        }
        if ((c + 1) % kFlush == 0 || c + 1 == nchunks) {                 // CTA-uniform
          mbar_wait(&bar_stage[st], use & 1);                            // MMAs execute in order: the group is complete
          tc_fence_after();
#pragma unroll
          for (int cb = 0; cb < CW; cb += 32) {
            float v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * CW + cb), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[cb + i] = __fadd_rn(acc[cb + i], v[i]);
          }
          tc_fence_before();                                             // ordered before the next group's first MMA by
        }                                                                // the __syncthreads of the next iteration
      }
    }
  }

  // ---- epilogue ------------------------------------------------------------------------------------------------------
  auto epilogue = [&](int m, int n, float v) {
    if (m >= M || n >= N) return;
    if (MODE == 0) {
      const float z = __fmul_rn(__fadd_rn(v, b_int ? __ldg(b_int + n) : 0.0f), __ldg(s_row + n));
      C[(long long)m * out_f + n] = act_fwd(z, act);
    } else if (MODE == 1) {
      C[(long long)m * in_f + n] = v;
    } else {
      float* dst = C + (long long)m * in_f + n;
      const float gq = __fdiv_rn(v, __ldg(s_row + m));
      *dst = accumulate ? __fadd_rn(*dst, gq) : gq;
    }
  };
  if (do_db) {                                                           // thread: channels t_m and t_m + 64, its k4
    db_s[lane_k4][t_m] = db_part[0];
    db_s[lane_k4][t_m + 64] = db_part[1];
  }
  float* red = reinterpret_cast<float*>(smem);                           // [BM][RED_LD] partial tile (split-K only)
  if (S == 1) {
    // thread = tile row: its CW consecutive outputs go out as 16-byte stores when the row pitch allows
    const int ldc = MODE == 0 ? out_f : in_f, m = m0 + row;
    const bool vec_c = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15u) == 0;
    if (m < M) {
      float sm = 0.0f;
      if (MODE == 2) sm = __ldg(s_row + m);
#pragma unroll
      for (int i = 0; i < CW; i += 4) {
        const int n = n0 + h * CW + i;
        if (vec_c && n + 3 < N) {
          float4 r;
          if (MODE == 0) {
            const float4 bb = b_int ? __ldg(reinterpret_cast<const float4*>(b_int + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 ss = __ldg(reinterpret_cast<const float4*>(s_row + n));
            r.x = act_fwd(__fmul_rn(__fadd_rn(acc[i + 0], bb.x), ss.x), act);
            r.y = act_fwd(__fmul_rn(__fadd_rn(acc[i + 1], bb.y), ss.y), act);
            r.z = act_fwd(__fmul_rn(__fadd_rn(acc[i + 2], bb.z), ss.z), act);
            r.w = act_fwd(__fmul_rn(__fadd_rn(acc[i + 3], bb.w), ss.w), act);
          } else if (MODE == 1) {
            r = make_float4(acc[i + 0], acc[i + 1], acc[i + 2], acc[i + 3]);
          } else {
            r = make_float4(__fdiv_rn(acc[i + 0], sm), __fdiv_rn(acc[i + 1], sm), __fdiv_rn(acc[i + 2], sm),
                            __fdiv_rn(acc[i + 3], sm));
            if (accumulate) {
              const float4 o4 = *reinterpret_cast<const float4*>(C + (long long)m * ldc + n);
              r.x = __fadd_rn(o4.x, r.x); r.y = __fadd_rn(o4.y, r.y); r.z = __fadd_rn(o4.z, r.z); r.w = __fadd_rn(o4.w, r.w);
            }
          }
          *reinterpret_cast<float4*>(C + (long long)m * ldc + n) = r;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) epilogue(m, n + u, acc[i + u]);
        }
      }
    }
  } else {
    __syncthreads();                                                     // the stage buffers are free: all MMAs retired
#pragma unroll
    for (int i = 0; i < CW; ++i) red[row * RED_LD + h * CW + i] = acc[i];
  }
  __syncthreads();
  if (do_db && tid < BM) {
    float t = db_s[0][tid];
#pragma unroll
    for (int j = 1; j < 4; ++j) t = __fadd_rn(t, db_s[j][tid]);
    db_cta[tid] = t;
  }
  if (S == 1) {
    if (do_db && tid < BM && m0 + tid < M) {
      const float gq = __fdiv_rn(db_cta[tid], __ldg(s_row + m0 + tid));
      db[m0 + tid] = accumulate ? __fadd_rn(db[m0 + tid], gq) : gq;
    }
  } else if constexpr (CL) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    const int rows_per = BM / S;                                         // S in {2, 4, 8}
    for (int e = tid; e < rows_per * BN; e += kThreads) {
      const int r = rank * rows_per + e / BN, cc = e % BN;
      float part[8];
#pragma unroll
      for (int p = 0; p < 8; ++p)                                        // all DSMEM loads in flight, then a fixed-order sum
        part[p] = p < S ? cluster.map_shared_rank(red, p)[r * RED_LD + cc] : 0.0f;
      float v = part[0];
#pragma unroll
      for (int p = 1; p < 8; ++p) if (p < S) v = __fadd_rn(v, part[p]);
      epilogue(m0 + r, n0 + cc, v);
    }
    if (do_db && rank == 0 && tid < BM && m0 + tid < M) {
      float t = db_cta[tid];
      for (int p = 1; p < S; ++p) t = __fadd_rn(t, cluster.map_shared_rank(db_cta, p)[tid]);
      const float gq = __fdiv_rn(t, __ldg(s_row + m0 + tid));
      db[m0 + tid] = accumulate ? __fadd_rn(db[m0 + tid], gq) : gq;
    }
    cluster.sync();                                                      // peers must not exit while being read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(BN) : "memory");
}

static int pick_split(int tiles, int K) {
  static const int max_s = [] { const char* e = getenv("DQRM_MLP_MAX_CLUSTER"); int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
  int S = 1;
  while (S < max_s && tiles * S < 128 && K / (S * 2) >= BK) S *= 2;     // every K-slice keeps >= 1 full chunk
  return S;
}

template <int MODE, int BN>
static int launch(const float* x, const float* W_int, const float* b_int, const float* s_row, const float* dout,
                  const float* out, float* C, float* db, int batch, int out_f, int in_f, int act, int accumulate,
                  cudaStream_t st) {
  const int M = MODE == 2 ? out_f : batch;
  const int N = MODE == 0 ? out_f : in_f;
  const int K = MODE == 0 ? in_f : (MODE == 1 ? out_f : batch);
  const int gx = (N + BN - 1) / BN, gy = (M + BM - 1) / BM;
  const int S = pick_split(gx * gy, K);
  int kc = (K + S - 1) / S;
  kc = ((kc + BK - 1) / BK) * BK;
  const size_t smem = (size_t)num_stages(MODE, BN) * stage_bytes(MODE, BN);
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  // 16-byte loads need every operand the mode reads along its contiguous dimension to be aligned (the Criteo layer
  // widths 13 / 367 / 415 are not multiples of 4: those layers take the scalar-load instantiation)
  const bool vec = MODE == 0 ? (in_f % 4 == 0 && al16(x) && al16(W_int))
                             : (MODE == 1 ? (out_f % 4 == 0 && al16(dout) && al16(out)) : false);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, gy, S);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = S;
  cfg.attrs = attr;
  cfg.numAttrs = S > 1 ? 1 : 0;
  cudaError_t e;
#define DQRM_TC_LAUNCH(CLV, VECV)                                                                                     \
  do {                                                                                                                \
    auto kern = linear_tc_kernel<MODE, BN, CLV, VECV>;                                                                \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                           \
    if (e == cudaSuccess)                                                                                             \
      e = cudaLaunchKernelEx(&cfg, kern, x, W_int, b_int, s_row, dout, out, C, db, batch, out_f, in_f, act, kc,       \
                             accumulate);                                                                             \
  } while (0)
  if (S > 1) { if (vec) DQRM_TC_LAUNCH(true, true); else DQRM_TC_LAUNCH(true, false); }
  else       { if (vec) DQRM_TC_LAUNCH(false, true); else DQRM_TC_LAUNCH(false, false); }
#undef DQRM_TC_LAUNCH
  if (e != cudaSuccess) { set_error("linear_tc_kernel<%d,%d>: %s", MODE, BN, cudaGetErrorString(e)); return -EIO; }
  return 0;
}

}  // namespace tc

// Entry used by dqrm_linear_fwd / dqrm_linear_bwd (mlp_fused.cu).  BN = 64 when the output is narrow or the grid
// would otherwise be small, else 128.
int launch_gemm_tc(int mode, const float* x, const float* W_int, const float* b_int, const float* s_row,
                   const float* dout, const float* out, float* C, float* db, int batch, int out_f, int in_f, int act,
                   int accumulate, cudaStream_t st) {
  const int M = mode == 2 ? out_f : batch;
  const int N = mode == 0 ? out_f : in_f;
  const bool narrow = N <= 64 || ((N + 127) / 128) * ((M + 127) / 128) < 64;
#define DQRM_TC(MODE)                                                                                                  \
  (narrow ? tc::launch<MODE, 64>(x, W_int, b_int, s_row, dout, out, C, db, batch, out_f, in_f, act, accumulate, st)    \
          : tc::launch<MODE, 128>(x, W_int, b_int, s_row, dout, out, C, db, batch, out_f, in_f, act, accumulate, st))
  if (mode == 0) return DQRM_TC(0);
  if (mode == 1) return DQRM_TC(1);
  return DQRM_TC(2);
#undef DQRM_TC
}

}  // namespace dqrm
