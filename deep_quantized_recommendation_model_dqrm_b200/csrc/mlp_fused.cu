// (a15) fused QuantLinear layers: one fake-quant launch for ALL layers of the model, then one kernel per
// layer forward and two per layer backward, each with its prologue / epilogue fused:
//   fwd : out = act((x W_int^t + b_int) * s_row)                       qm:209 + the ReLU / Sigmoid that follows
//   dx  : dx  = g W_int,      g = dout * act'(out) * s_row             autograd of qm:209
//   dw  : dW += (g^t x) / s_row ;  db += (sum_b g) / s_row             STE of SymmetricQuantFunction (qu:363)
// Reference: QuantLinear.forward, quantization_supp/quant_modules_not_quantize_grad.py:105-211.
//
// The reference step spends ~95 launches here (per layer: 4 reductions, 6 pointwise, addmm, mul, relu;
// backward: 3 GEMM/GEMV, 5 pointwise, a column reduction, 2 AccumulateGrad adds).  These layers are tiny
// (batch 128: 33 MFLOP for the largest, 0.95 MFLOP/sample in total) -- launch latency, not FLOPs, so the
// win is the launch count: 1 + 7 + 13 launches.  fp32 FFMA on purpose: parity is 1e-5 against an fp32
// reference (allow_tf32 = False there) and one layer is far below a tcgen05 tile per SM.  No atomics and
// no split-K: the summation order is fixed, so data-parallel replicas stay bit-identical.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "common.cuh"

namespace dqrm {

constexpr int kMaxLayers = 16;

struct MlpLayers {
  const float* W[kMaxLayers];
  const float* b[kMaxLayers];
  float* W_int[kMaxLayers];
  float* b_int[kMaxLayers];
  float* s[kMaxLayers];
  int out_f[kMaxLayers];
  int in_f[kMaxLayers];
  int row_begin[kMaxLayers + 1];
  int num_layers;
};

// one warp per weight row over all layers
__global__ void __launch_bounds__(256)
mlp_fakequant_all_kernel(const __grid_constant__ MlpLayers L, int bits) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= L.row_begin[L.num_layers]) return;
  int l = 0;
  while (row >= L.row_begin[l + 1]) ++l;
  const int r = row - L.row_begin[l];
  const int in_f = L.in_f[l];
  const float* w = L.W[l] + (long long)r * in_f;
  unsigned m = 0u;
  for (int i = lane; i < in_f; i += 32) m = max(m, abs_bits(w[i]));
  m = warp_max_u32(m);
  const float s = scale_of(__uint_as_float(m), bits);
  const float inv = __fdiv_rn(1.0f, s);
  const float hi = qmax_of(bits), lo = -hi - 1.0f;
  float* q = L.W_int[l] + (long long)r * in_f;
  for (int i = lane; i < in_f; i += 32) q[i] = quant_code(w[i], inv, lo, hi);
  if (lane == 0) {
    L.s[l][r] = s;
    if (L.b[l]) L.b_int[l][r] = quant_code(L.b[l][r], inv, lo, hi);
  }
}

enum Act { kActNone = 0, kActRelu = 1, kActSigmoid = 2 };

__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == kActRelu) return fmaxf(z, 0.0f);                           // torch.relu == clamp_min(0)
  if (act == kActSigmoid) return 1.0f / (1.0f + expf(-z));
  return z;
}
__device__ __forceinline__ float act_bwd(float dout, float out, int act) {
  if (act == kActRelu) return out > 0.0f ? dout : 0.0f;                 // threshold_backward
  if (act == kActSigmoid) return dout * ((1.0f - out) * out);           // sigmoid_backward
  return dout;
}

// Small fp32 GEMM, C[M,N] = sum_k A(m,k) B(k,n):  CTA tile 32 x 64, 128 threads, 4 x 4 outputs per thread.
// MODE 0 (fwd): A = x[M=batch, K=in] (k contiguous)         B(k,n) = W_int[n][k]
// MODE 1 (dx) : A(m,k) = g(batch m, out k)                   B(k,n) = W_int[k][n]   (N = in)
// MODE 2 (dw) : A(m,k) = g(batch k, out m)                   B(k,n) = x[k][n]       (M = out, N = in, K = batch)
//
// At batch 128 a layer has only 16-96 output tiles, far fewer than 148 SMs, and its K loop is the whole
// latency.  So K is split across a THREAD-BLOCK CLUSTER (1,1,S), S <= 8: CTA r of the cluster multiplies
// its K-slice, parks the partial tile in its shared memory, and after a cluster barrier every CTA sums
// one row-slice of the tile over ranks 0..S-1 through distributed shared memory -- a fixed order, so the
// result is deterministic (no atomics, no workspace; replicas stay bit-identical) -- and applies the
// epilogue for that slice.  Inside a CTA the K loop is a register-prefetch double buffer.
constexpr int BM = 32, BN = 64, TM = 4, TN = 4, kGemmThreads = 128;
constexpr int kSliceMin = 64;                                 // a cluster rank's K-slice is at least this long
constexpr int APAD = BM + 4, BPAD = BN + 4;

// CL = false: the same kernel with no cluster instruction at all (S = 1).  A grid that uses clusters was
// measured not to become co-resident with a long-running non-cluster grid (the pipelined table rescan,
// tracker.cu): the step's GEMMs waited for the whole pass.  See DESIGN.md "pipelined rescan".
// BK: K-tile depth (16 or 32).  A "prefetch the whole K-slice into registers before the first FMA" variant
// (DQRM_GEMM_PREFETCH of round 1) was measured SLOWER on B200 (dW 11.0 vs 7.8 us) and is gone.
template <int MODE, bool CL, int BK>
__global__ void __launch_bounds__(kGemmThreads)
linear_gemm_kernel(const float* __restrict__ x, const float* __restrict__ W_int, const float* __restrict__ b_int,
                   const float* __restrict__ s_row, const float* __restrict__ dout, const float* __restrict__ out,
                   float* __restrict__ C, float* __restrict__ db, int batch, int out_f, int in_f, int act, int kc,
                   int accumulate, int serial) {
  namespace cg = cooperative_groups;
  int S = 1, rank = 0;
  if constexpr (CL) {
    cg::cluster_group cluster = cg::this_cluster();
    S = (int)cluster.num_blocks();
    rank = (int)cluster.block_rank();
  }
  constexpr int A_PER_THR = BM * BK / kGemmThreads, B_PER_THR = BK * BN / kGemmThreads;    // BK 16: 4, 8
  __shared__ __align__(16) float As[2][BK][APAD];
  __shared__ __align__(16) float Bs[2][BK][BPAD];                       // also the partial tile red[BM][BN]
  static_assert(2 * BK * BPAD >= BM * BN, "partial tile must fit in Bs");
  const int M = MODE == 2 ? out_f : batch;
  const int N = MODE == 0 ? out_f : in_f;
  const int K = MODE == 0 ? in_f : (MODE == 1 ? out_f : batch);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;                               // thread tile rows ty*4.., cols tx*4..
  int kbeg = rank * kc, kend = min(K, kbeg + kc);

  // Staging loads are BRANCH-FREE: out-of-range elements read a clamped (valid) address and are zeroed by a select
  // afterwards, and the activation derivative is a select too.  With `if (m >= M) return 0` / `out > 0 ? dout : 0`
  // written as control flow, ptxas emitted one reconvergence region per element -- load dout+out, wait, branch, load
  // s_row, wait -- i.e. 8 serialised load latencies per K-tile (ncu source page, profiles/r02_linear_gemm_ncu.txt);
  // now every load of a K-tile is in flight before the first use.  Same arithmetic, same bits.
  struct Raw { float d, o, s; bool ok; };
  auto A_issue = [&](int m, int k) -> Raw {
    Raw r;
    r.ok = m < M && k < kend;
    const int mc = min(m, M - 1), kk = min(k, kend - 1);
    if (MODE == 0) {
      r.d = __ldg(x + (long long)mc * in_f + kk);
      r.o = 0.0f; r.s = 0.0f;
    } else {
      const int bi = MODE == 1 ? mc : kk, oi = MODE == 1 ? kk : mc;      // g(batch bi, out oi) = dout * act'(out) * s_row
      const long long e = (long long)bi * out_f + oi;
      r.d = __ldg(dout + e);
      r.o = __ldg(out + e);
      r.s = __ldg(s_row + oi);
    }
    return r;
  };
  auto A_finish = [&](const Raw& r) -> float {
    if (MODE == 0) return r.ok ? r.d : 0.0f;
    const float relu = r.o > 0.0f ? r.d : 0.0f;                           // threshold_backward
    const float sigm = __fmul_rn(r.d, __fmul_rn(__fsub_rn(1.0f, r.o), r.o));   // sigmoid_backward
    const float g = act == kActRelu ? relu : (act == kActSigmoid ? sigm : r.d);
    return r.ok ? __fmul_rn(g, r.s) : 0.0f;
  };
  auto B_issue = [&](int k, int n, bool& ok) -> float {
    ok = k < kend && n < N;
    const int kk = min(k, kend - 1), nc = min(n, N - 1);
    if (MODE == 0) return __ldg(W_int + (long long)nc * in_f + kk);
    if (MODE == 1) return __ldg(W_int + (long long)kk * in_f + nc);
    return __ldg(x + (long long)kk * in_f + nc);
  };
  // element -> (row, k) mappings chosen so that consecutive threads read consecutive addresses
  auto a_coord = [&](int e, int& am, int& ak) {
    if (MODE == 2) { am = e & (BM - 1); ak = e / BM; }                  // g(k, m): contiguous along m (out)
    else           { ak = e & (BK - 1); am = e / BK; }                  // contiguous along k
  };
  auto b_coord = [&](int e, int& bk, int& bn) {
    if (MODE == 0) { bk = e & (BK - 1); bn = e / BK; }                  // W_int[n][k]: contiguous along k
    else           { bn = e & (BN - 1); bk = e / BN; }                  // contiguous along n
  };

  // MODE 2: the A tiles ARE g, so the bias gradient (column sums of g over the batch) is accumulated for
  // free from the elements each thread stages; only the first column of tiles needs it
  const bool do_db = MODE == 2 && db != nullptr && blockIdx.x == 0;
  float db_part = 0.0f;
  float a_reg[A_PER_THR], b_reg[B_PER_THR];
  auto fetch = [&](int k0) {
    Raw ar[A_PER_THR];
    bool bok[B_PER_THR];
#pragma unroll
    for (int i = 0; i < A_PER_THR; ++i) { int am, ak; a_coord(tid + i * kGemmThreads, am, ak); ar[i] = A_issue(m0 + am, k0 + ak); }
#pragma unroll
    for (int i = 0; i < B_PER_THR; ++i) { int bk, bn; b_coord(tid + i * kGemmThreads, bk, bn); b_reg[i] = B_issue(k0 + bk, n0 + bn, bok[i]); }
#pragma unroll
    for (int i = 0; i < A_PER_THR; ++i) a_reg[i] = A_finish(ar[i]);
#pragma unroll
    for (int i = 0; i < B_PER_THR; ++i) b_reg[i] = bok[i] ? b_reg[i] : 0.0f;
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER_THR; ++i) {
      int am, ak; a_coord(tid + i * kGemmThreads, am, ak); As[buf][ak][am] = a_reg[i];
      if (MODE == 2) db_part = __fadd_rn(db_part, a_reg[i]);            // row am == tid & 31 for every i
    }
#pragma unroll
    for (int i = 0; i < B_PER_THR; ++i) { int bk, bn; b_coord(tid + i * kGemmThreads, bk, bn); Bs[buf][bk][bn] = b_reg[i]; }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  {
    // serial > 1 (CL = false, forward only): ONE CTA walks the `serial` K-slices that a cluster of that size would
    // have split between its CTAs, one after the other, and adds the slice sums in rank order -- the same bits as the
    // cluster launch, without a cluster launch (which does not become co-resident beside a long-running grid: the
    // bottom MLP that runs next to the table scan uses this, DQRM_LINEAR_FFMA_SERIAL)
    const int slices = (!CL && serial > 1) ? serial : 1;
    float tot[TM][TN];
    for (int sl = 0; sl < slices; ++sl) {
      if (slices > 1) { kbeg = sl * kc; kend = min(K, kbeg + kc); }
      if (kbeg < kend) {
        fetch(kbeg);
        stash(0);
        __syncthreads();
        int buf = 0;
        for (int k0 = kbeg; k0 < kend; k0 += BK) {
          const bool more = k0 + BK < kend;
          if (more) fetch(k0 + BK);                                     // loads in flight during the FMAs below
#pragma unroll
          for (int k = 0; k < BK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
            const float a[TM] = {av.x, av.y, av.z, av.w}, b[TN] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
              for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
          }
          if (more) stash(buf ^ 1);
          __syncthreads();
          buf ^= 1;
        }
      }
      if (slices > 1) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            tot[i][j] = sl == 0 ? acc[i][j] : __fadd_rn(tot[i][j], acc[i][j]);
            acc[i][j] = sl + 1 == slices ? tot[i][j] : 0.0f;
          }
      }
    }
  }

  float* red = &Bs[0][0][0];                                            // [BM][BN]
  __shared__ float db_slice[kGemmThreads / BM][BM];
  __shared__ float db_cta[BM];
  if (do_db) {                                                           // CTA-level: 4 k-interleaved partials per row
    db_slice[tid / BM][tid & (BM - 1)] = db_part;
    __syncthreads();
    if (tid < BM) {
      float t = db_slice[0][tid];
#pragma unroll
      for (int sl = 1; sl < kGemmThreads / BM; ++sl) t = __fadd_rn(t, db_slice[sl][tid]);
      db_cta[tid] = t;
    }
  }
  auto epilogue = [&](int m, int n, float v) {
    if (m >= M || n >= N) return;
    if (MODE == 0) {
      const float z = __fmul_rn(__fadd_rn(v, b_int ? b_int[n] : 0.0f), s_row[n]);
      C[(long long)m * out_f + n] = act_fwd(z, act);
    } else if (MODE == 1) {
      C[(long long)m * in_f + n] = v;
    } else {
      float* dst = C + (long long)m * in_f + n;                          // straight into the grad arena
      const float gq = __fdiv_rn(v, s_row[m]);
      *dst = accumulate ? __fadd_rn(*dst, gq) : gq;
    }
  };
  if (S == 1) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) epilogue(m0 + ty * TM + i, n0 + tx * TN + j, acc[i][j]);
    if (do_db) {
      __syncthreads();
      if (tid < BM && m0 + tid < M) {
        const float gq = __fdiv_rn(db_cta[tid], s_row[m0 + tid]);
        db[m0 + tid] = accumulate ? __fadd_rn(db[m0 + tid], gq) : gq;
      }
    }
  } else if constexpr (CL) {
    cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
    for (int i = 0; i < TM; ++i)
      *reinterpret_cast<float4*>(&red[(ty * TM + i) * BN + tx * TN]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    cluster.sync();
    const int rows_per = BM / S;                                        // S in {2,4,8}
    for (int e = tid; e < rows_per * BN; e += kGemmThreads) {
      const int r = rank * rows_per + e / BN, c = e % BN;
      float part[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)                                        // all DSMEM loads in flight, then a fixed-order sum
        part[q] = q < S ? cluster.map_shared_rank(red, q)[r * BN + c] : 0.0f;
      float v = part[0];
#pragma unroll
      for (int q = 1; q < 8; ++q) if (q < S) v = __fadd_rn(v, part[q]);
      epilogue(m0 + r, n0 + c, v);
    }
    if (do_db && rank == 0 && tid < BM && m0 + tid < M) {
      float t = db_cta[tid];
      for (int q = 1; q < S; ++q) t = __fadd_rn(t, cluster.map_shared_rank(db_cta, q)[tid]);
      const float gq = __fdiv_rn(t, s_row[m0 + tid]);
      db[m0 + tid] = accumulate ? __fadd_rn(db[m0 + tid], gq) : gq;
    }
    cluster.sync();                                                      // peers must not exit while being read
  }
}

// cluster size: enough K-slices to put >= ~128 CTAs on the chip, each slice >= kSliceMin long
// (DQRM_MLP_MAX_CLUSTER caps it -- a diagnostic knob: 1 = no cluster launch at all)
static int pick_split(int tiles, int K) {
  static const int max_s = [] { const char* e = getenv("DQRM_MLP_MAX_CLUSTER"); int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
  int S = 1;
  while (S < max_s && tiles * S < 128 && K / (S * 2) >= kSliceMin / 2) S *= 2;
  return S;
}

template <int MODE>
static int launch_gemm(const float* x, const float* W_int, const float* b_int, const float* s_row, const float* dout,
                       const float* out, float* C, float* db, int batch, int out_f, int in_f, int act, int accumulate,
                       cudaStream_t st, bool serial_slices = false) {
  const int M = MODE == 2 ? out_f : batch;
  const int N = MODE == 0 ? out_f : in_f;
  const int K = MODE == 0 ? in_f : (MODE == 1 ? out_f : batch);
  const int gx = (N + BN - 1) / BN, gy = (M + BM - 1) / BM;
  int S = pick_split(gx * gy, K);
  int kc = (K + S - 1) / S;
  kc = ((kc + 31) / 32) * 32;                                            // a multiple of either K-tile depth: same slices, same bits
  const int serial = (serial_slices && MODE == 0) ? S : 1;              // the same slices, walked by one CTA
  if (serial > 1) S = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, gy, S);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = S;
  cfg.attrs = attr;
  cfg.numAttrs = S > 1 ? 1 : 0;
  // K-tile depth: 32 halves the number of dependent load round trips of a slice (DQRM_GEMM_BK=16 for comparison)
  static const int bk = [] { const char* e = getenv("DQRM_GEMM_BK"); return (e && atoi(e) == 16) ? 16 : 32; }();
  cudaError_t e;
#define DQRM_GEMM_LAUNCH(BKV)                                                                                        \
  do {                                                                                                               \
    if (S > 1) {                                                                                                     \
      e = cudaLaunchKernelEx(&cfg, linear_gemm_kernel<MODE, true, BKV>, x, W_int, b_int, s_row, dout, out, C, db,    \
                             batch, out_f, in_f, act, kc, accumulate, 1);                                            \
    } else {                                                                                                         \
      linear_gemm_kernel<MODE, false, BKV><<<cfg.gridDim, cfg.blockDim, 0, st>>>(x, W_int, b_int, s_row, dout, out,  \
                                                                                 C, db, batch, out_f, in_f, act, kc, \
                                                                                 accumulate, serial);                \
      e = cudaGetLastError();                                                                                        \
    }                                                                                                                \
  } while (0)
  if (bk == 16) DQRM_GEMM_LAUNCH(16);
  else DQRM_GEMM_LAUNCH(32);
#undef DQRM_GEMM_LAUNCH
  if (e != cudaSuccess) { set_error("linear_gemm_kernel<%d>: %s", MODE, cudaGetErrorString(e)); return -EIO; }
  return 0;
}

// mlp_tc.cu: the same three layer kernels with the contraction on tcgen05 tensor cores (TF32 split, fp32 parity)
int launch_gemm_tc(int mode, const float* x, const float* W_int, const float* b_int, const float* s_row,
                   const float* dout, const float* out, float* C, float* db, int batch, int out_f, int in_f, int act,
                   int accumulate, cudaStream_t st);

// path: DQRM_LINEAR_AUTO / _FFMA / _TC.  AUTO: tensor cores from DQRM_MLP_TC_MIN_BATCH rows (default 256) -- below
// that a layer is a handful of tiles and the cluster split-K FFMA kernel's shorter prologue wins -- and only for layers
// whose weight matrix is at least DQRM_MLP_TC_MIN_DIM (default 32) in both directions: the 13-wide first bottom layer and
// the 1-wide last top layer would pad a 128 x 64 x 32 tensor-core tile with 60-98 % zeros (batch 2048: 18.5 us for
// 2048 x 512 x 13, 12.2 us for 2048 x 1 x 256 on the tensor-core kernel; profiles/r02_timeline_n1_batch2048.txt).
// The weight gradient contracts over the batch: from 1024 rows on the FFMA kernel's 8-way cluster split leaves slices too
// long (512 x 13 x 2048: 28 us against 20 us), so it stays on the tensor cores whatever the layer's shape.
static bool use_tc(int path, int batch, int out_f, int in_f, bool weight_grad = false) {
  static const int min_batch = [] { const char* e = getenv("DQRM_MLP_TC_MIN_BATCH"); return e ? atoi(e) : 256; }();
  static const int min_dim = [] { const char* e = getenv("DQRM_MLP_TC_MIN_DIM"); return e ? atoi(e) : 32; }();
  if (path == DQRM_LINEAR_FFMA || path == DQRM_LINEAR_FFMA_SERIAL) return false;
  if (path == DQRM_LINEAR_TC) return true;
  return batch >= min_batch && ((out_f >= min_dim && in_f >= min_dim) || (weight_grad && batch >= 1024));
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_mlp_fakequant_all(int num_layers, const float* const* W, const float* const* b,
                                      const int32_t* out_features, const int32_t* in_features, int bits,
                                      float* const* W_int, float* const* b_int, float* const* scale_row, void* stream) {
  DQRM_REQUIRE(num_layers >= 1 && num_layers <= kMaxLayers, -E2BIG, "mlp_fakequant_all: num_layers=%d (max %d)", num_layers, kMaxLayers);
  DQRM_REQUIRE(W && out_features && in_features && W_int && scale_row, -EINVAL, "mlp_fakequant_all: null argument");
  DQRM_REQUIRE(bits >= 2 && bits <= 16, -EINVAL, "mlp_fakequant_all: bits=%d outside [2,16]", bits);
  MlpLayers L;
  L.num_layers = num_layers;
  int rows = 0;
  for (int l = 0; l < num_layers; ++l) {
    DQRM_REQUIRE(W[l] && W_int[l] && scale_row[l] && out_features[l] >= 1 && in_features[l] >= 1, -EINVAL,
                 "mlp_fakequant_all: layer %d malformed", l);
    L.W[l] = W[l]; L.b[l] = b ? b[l] : nullptr; L.W_int[l] = W_int[l]; L.b_int[l] = b_int ? b_int[l] : nullptr;
    DQRM_REQUIRE((L.b[l] == nullptr) == (L.b_int[l] == nullptr), -EINVAL, "mlp_fakequant_all: layer %d b/b_int mismatch", l);
    L.s[l] = scale_row[l]; L.out_f[l] = out_features[l]; L.in_f[l] = in_features[l];
    L.row_begin[l] = rows;
    rows += out_features[l];
  }
  L.row_begin[num_layers] = rows;
  mlp_fakequant_all_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(L, bits);
  DQRM_LAUNCH_CHECK("mlp_fakequant_all_kernel");
  return 0;
}

extern "C" int dqrm_linear_fwd(const float* x, const float* W_int, const float* b_int, const float* scale_row,
                               int batch, int out_features, int in_features, int act, float* out, int path, void* stream) {
  DQRM_REQUIRE(x && W_int && scale_row && out, -EINVAL, "linear_fwd: null argument");
  DQRM_REQUIRE(batch >= 1 && out_features >= 1 && in_features >= 1 && act >= 0 && act <= 2, -EINVAL, "linear_fwd: bad shape/act");
  DQRM_REQUIRE(path >= 0 && path <= 3, -EINVAL, "linear_fwd: path=%d", path);
  if (use_tc(path, batch, out_features, in_features))
    return launch_gemm_tc(0, x, W_int, b_int, scale_row, nullptr, nullptr, out, nullptr, batch, out_features, in_features,
                          act, 0, static_cast<cudaStream_t>(stream));
  return launch_gemm<0>(x, W_int, b_int, scale_row, nullptr, nullptr, out, nullptr, batch, out_features, in_features, act,
                        0, static_cast<cudaStream_t>(stream), path == DQRM_LINEAR_FFMA_SERIAL);
}

extern "C" int dqrm_linear_bwd(const float* x, const float* W_int, const float* scale_row, const float* dout,
                               const float* out, int batch, int out_features, int in_features, int act,
                               float* dx, float* dW, float* db, int accumulate, int path, void* stream) {
  DQRM_REQUIRE(x && W_int && scale_row && dout && out && (dW || dx), -EINVAL, "linear_bwd: null argument");
  DQRM_REQUIRE(batch >= 1 && out_features >= 1 && in_features >= 1 && act >= 0 && act <= 2, -EINVAL, "linear_bwd: bad shape/act");
  DQRM_REQUIRE(path >= 0 && path <= 3, -EINVAL, "linear_bwd: path=%d", path);     // (_SERIAL: same as _FFMA here)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tcp = use_tc(path, batch, out_features, in_features);
  const bool tcw = use_tc(path, batch, out_features, in_features, true);
  if (dx) {
    const int rc = tcp ? launch_gemm_tc(1, x, W_int, nullptr, scale_row, dout, out, dx, nullptr, batch, out_features,
                                        in_features, act, 0, st)
                       : launch_gemm<1>(x, W_int, nullptr, scale_row, dout, out, dx, nullptr, batch, out_features,
                                        in_features, act, 0, st);
    if (rc) return rc;
  }
  if (!dW) return 0;                                      // dx only (the caller runs dW on another stream)
  if (tcw)
    return launch_gemm_tc(2, x, W_int, nullptr, scale_row, dout, out, dW, db, batch, out_features, in_features, act,
                          accumulate ? 1 : 0, st);
  return launch_gemm<2>(x, W_int, nullptr, scale_row, dout, out, dW, db, batch, out_features, in_features, act,
                        accumulate ? 1 : 0, st);
}
