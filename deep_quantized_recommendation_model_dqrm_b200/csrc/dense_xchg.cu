// (a11 + the MLP half of a9) the whole dense-gradient exchange of one step as ONE kernel over NVLink peer memory.
//
// Reference: quantize_linear_grad / quantize_bias_grad for every MLP tensor (sgd_quantized_gradients_parallel_comm.py:
// 892-961: local per-channel scale -> all_reduce(scale)/N -> 8-bit codes -> all_reduce(codes)/N) followed by the MLP
// part of weight_update_parallel_comm (:630-663).  28 Gloo all-reduces + ~140 launches per step there; the portable
// form of this library is scale kernel -> all-gather -> quantise -> all-gather -> apply (csrc/mlp.cu, csrc/p2p.cu),
// five dependent launches whose two all-gathers each pay a launch, a grid drain and a flag round trip: 47 us at two
// GPUs and 63 us at eight, all of it on the step's critical tail.
//
// Channels are independent, so no grid-wide step is needed anywhere: CTA b owns a contiguous run of channels (the same
// run on every rank) and only ever talks to CTA b of the peers --
//   1. load grad (+ error compensation) of its elements into shared memory, per-channel max-abs -> local scales,
//      stored into slot[rank] of the scale site of every PEER arena (remote stores);
//   2. wait for the peers' scales; s_bar = (sum_r s_r) * (1/N) in rank order; int8 codes -> slot[rank] of the code
//      site of every peer arena;
//   3. wait for the peers' codes; integer sum of the N code sets in rank order; W += (-lr * (sum/N)) * s_bar (and the
//      error-compensation residual).
// Same arithmetic, same order as dense_grad_scale / dense_grad_quant_gathered / dense_apply_gathered: bit-identical
// parameters (tests/test_gpu_p2p.py; tools/p2p_check.py against the NCCL transport on real GPUs).
//
// Transport: every 8-byte word in a slot carries its payload AND the step's sequence number -- {fp32 scale, u32 seq} or
// {seven int8 codes, u8 seq} -- written with ONE 64-bit store and polled by the reader until the sequence is this
// step's (the idea of NCCL's low-latency protocol).  Payload and flag arrive together, so there is no system-scope
// fence and no separate flag: a wait costs one NVLink one-way trip.  The first version (data stores, st.release.sys flag
// per CTA, ld.acquire poll) measured 4.4-6.4 us per wait at two GPUs (profiles/r02_dense_xchg_phases_n2_fence.txt)
// against ~2.5 us of work per phase.  The codes pay 8/7 of their bytes (0.54 MB per peer at Kaggle shape); a stale word
// always carries the previous step's sequence, so one byte of it is enough.
//
// Single buffering: rank r overwrites its scale words at peer p in step s+1 only after it received p's codes of step s,
// which p computed from those scales; it overwrites its code words only after p's scales of step s+1, which p sends
// after its kernel of step s (the reader of the codes) retired.  Words are read with 64-bit relaxed system-scope loads
// (L2).  All CTAs of the launch must be able to become resident (num_ctas <= 148); a peer that never arrives trips the
// same watchdog / sticky status bit as dqrm_p2p_allgather and this CTA skips its apply.
#include <stdlib.h>

#include "common.cuh"

namespace dqrm {

constexpr int kXThreads = 256;
constexpr int kXMaxWorld = 16;

struct XPeers { unsigned char* base[kXMaxWorld]; };

struct XSites {
  size_t scale_off, scale_stride;     // data area of the scale site (bytes), stride between rank slots: u64 [num_chan]
  size_t code_off, code_stride;       // u64 [cta_word[num_ctas]]: seven codes per word, every CTA's run padded to a word
};

constexpr int kXCodesPerWord = 7;

__device__ __forceinline__ void x_st_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long x_ld_word(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Poll word `idx` of every peer's slot in the LOCAL arena until the top SEQ_BITS of each equal `seq`; all loads of a round
// are in flight together.  The words land in w[r] (w[rank] untouched).  false: the watchdog fired.
template <int SEQ_BITS>
__device__ __forceinline__ bool x_wait_words(const unsigned char* slots, size_t stride, long long idx, int world, int rank,
                                             unsigned seq, long long timeout_cycles, int* status, unsigned long long* w) {
  const unsigned long long want = (unsigned long long)(seq & (unsigned)((1ull << SEQ_BITS) - 1ull));
  bool all = false;
  long long t0 = 0;
  while (!all) {
#pragma unroll
    for (int r = 0; r < kXMaxWorld; ++r)
      if (r < world && r != rank) w[r] = x_ld_word(reinterpret_cast<const unsigned long long*>(slots + (size_t)r * stride) + idx);
    all = true;
#pragma unroll
    for (int r = 0; r < kXMaxWorld; ++r)
      if (r < world && r != rank) all = all && (w[r] >> (64 - SEQ_BITS)) == want;
    if (!all) {
      if (t0 == 0) t0 = clock64();
      if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) { atomicOr(status, DQRM_STATUS_P2P_TIMEOUT); return false; }
      __nanosleep(100);                     // thousands of threads poll: leave the L2 to the kernels running beside this one
    }
  }
  return true;
}

__device__ __forceinline__ void x_stamp(unsigned long long* dbg, int k) {      // phase time stamps (debug hook, NULL = off)
  if (dbg && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[blockIdx.x * 8 + k] = t;
  }
}

__global__ void __launch_bounds__(kXThreads)
dense_exchange_apply_kernel(const __grid_constant__ XPeers peers, const __grid_constant__ XSites st, int world, int rank,
                            float* __restrict__ param, float* __restrict__ grad, float* error_comp,
                            const long long* __restrict__ chan_begin, const int* __restrict__ cta_chan,
                            const int* __restrict__ cta_word, int elem_cap, int chan_cap, int bits, float inv_world,
                            float* __restrict__ scale_mean, unsigned* __restrict__ seq_dev, float neg_lr_arg,
                            const float* __restrict__ lr_dev, long long timeout_cycles, int* status,
                            unsigned long long* dbg) {
  extern __shared__ uint4 x_smem[];
  // g_s fp32 [elem_cap] | sbar_s fp32 [chan_cap] | inv_s fp32 [chan_cap] | cb_s i32 [chan_cap + 4] | qsum_s i16
  // [elem_cap + 32] | chrel_s u16 [elem_cap] | q_s i8 [elem_cap + 32]     (elem_cap % 16 == 0, chan_cap % 4 == 0)
  float* g_s = reinterpret_cast<float*>(x_smem);
  float* sbar_s = g_s + elem_cap;
  float* inv_s = sbar_s + chan_cap;
  int* cb_s = reinterpret_cast<int*>(inv_s + chan_cap);
  short* qsum_s = reinterpret_cast<short*>(cb_s + chan_cap + 4);
  unsigned short* chrel_s = reinterpret_cast<unsigned short*>(qsum_s + elem_cap + 32);
  signed char* q_s = reinterpret_cast<signed char*>(chrel_s + elem_cap);

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = cta_chan[b], c1 = cta_chan[b + 1], nch = c1 - c0;
  const long long w0 = cta_word[b];
  const unsigned seq = seq_dev[b] + 1u;
  const long long e0 = chan_begin[c0], e1 = chan_begin[c1];
  const int n = (int)(e1 - e0), nw = (n + kXCodesPerWord - 1) / kXCodesPerWord;
  const unsigned char* local = peers.base[rank];
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;

  x_stamp(dbg, 0);
  // ---- 1. gradients (+ error compensation) -> shared memory; per-channel max-abs -> local scales -> every peer
  for (int k = tid; k <= nch; k += kXThreads) cb_s[k] = (int)(chan_begin[c0 + k] - e0);
  for (int i0 = tid; i0 < n; i0 += 4 * kXThreads) {
    float w[4], c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int i = i0 + j * kXThreads; w[j] = grad[e0 + (i < n ? i : n - 1)]; }
    if (error_comp) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int i = i0 + j * kXThreads; c[j] = error_comp[e0 + (i < n ? i : n - 1)]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * kXThreads;
      if (i >= n) continue;
      float v = w[j];
      if (error_comp) { v = __fadd_rn(v, c[j]); grad[e0 + i] = v; }      // weight = grad + error_compensation (:899-900)
      g_s[i] = v;
    }
  }
  __syncthreads();
  for (int k = warp; k < nch; k += kXThreads / 32) {
    const int a = cb_s[k], e = cb_s[k + 1];
    unsigned m = 0u;
    for (int i = a + lane; i < e; i += 32) { m = max(m, abs_bits(g_s[i])); chrel_s[i] = (unsigned short)k; }
    m = warp_max_u32(m);
    if (lane == 0) sbar_s[k] = scale_of(__uint_as_float(m), bits);
  }
  __syncthreads();
  for (int k = 1; k < world; ++k) {
    const int p = (rank + k) % world;                                     // every rank starts on a different peer
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(peers.base[p] + st.scale_off + (size_t)rank * st.scale_stride) + c0;
    for (int i = tid; i < nch; i += kXThreads) x_st_u64(dst + i, ((unsigned long long)seq << 32) | __float_as_uint(sbar_s[i]));
  }
  x_stamp(dbg, 1);

  // ---- 2. the peers' scales; mean scale in rank order; int8 codes -> every peer
  for (int k = tid; k < nch; k += kXThreads) {
    unsigned long long sc[kXMaxWorld];
    if (!x_wait_words<32>(local + st.scale_off, st.scale_stride, c0 + k, world, rank, seq, timeout_cycles, status, sc)) s_bad = 1;
    const float own = sbar_s[k];
    float acc = rank == 0 ? own : __uint_as_float((unsigned)sc[0]);       // (static indices only: sc[] stays in registers)
#pragma unroll
    for (int r = 1; r < kXMaxWorld; ++r)
      if (r < world) acc = __fadd_rn(acc, r == rank ? own : __uint_as_float((unsigned)sc[r]));
    const float s_bar = __fmul_rn(acc, inv_world);
    sbar_s[k] = s_bar;
    inv_s[k] = __fdiv_rn(1.0f, s_bar);
    scale_mean[c0 + k] = s_bar;
  }
  __syncthreads();
  x_stamp(dbg, 2);
  {
    const float hi = qmax_of(bits), lo = -hi - 1.0f;
    for (int i = tid; i < kXCodesPerWord * nw; i += kXThreads)
      q_s[i] = i < n ? (signed char)quant_code(g_s[i], inv_s[chrel_s[i]], lo, hi) : (signed char)0;
  }
  __syncthreads();
  for (int k = 1; k < world; ++k) {
    const int p = (rank + k) % world;
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(peers.base[p] + st.code_off + (size_t)rank * st.code_stride) + w0;
    for (int w = tid; w < nw; w += kXThreads) {
      unsigned long long v = (unsigned long long)(seq & 0xffu) << 56;
#pragma unroll
      for (int j = 0; j < kXCodesPerWord; ++j) v |= (unsigned long long)(unsigned char)q_s[kXCodesPerWord * w + j] << (8 * j);
      x_st_u64(dst + w, v);
    }
  }
  x_stamp(dbg, 3);

  // ---- 3. the peers' codes; exact integer sum in rank order; SGD update (and the error-compensation residual)
  for (int w = tid; w < nw; w += kXThreads) {
    unsigned long long cw[kXMaxWorld];
    if (!x_wait_words<8>(local + st.code_off, st.code_stride, w0 + w, world, rank, seq, timeout_cycles, status, cw)) s_bad = 1;
    short acc[kXCodesPerWord];
#pragma unroll
    for (int j = 0; j < kXCodesPerWord; ++j) acc[j] = 0;
#pragma unroll
    for (int r = 0; r < kXMaxWorld; ++r) {
      if (r >= world) continue;
#pragma unroll
      for (int j = 0; j < kXCodesPerWord; ++j)
        acc[j] += r == rank ? (short)q_s[kXCodesPerWord * w + j] : (short)(signed char)((cw[r] >> (8 * j)) & 0xffu);
    }
#pragma unroll
    for (int j = 0; j < kXCodesPerWord; ++j) qsum_s[kXCodesPerWord * w + j] = acc[j];
  }
  __syncthreads();
  x_stamp(dbg, 4);
  if (tid == 0) seq_dev[b] = seq;
  if (s_bad || (*reinterpret_cast<volatile int*>(status) & DQRM_STATUS_P2P_TIMEOUT)) return;   // never apply stale slots
  const float neg_lr = lr_dev ? -(*lr_dev) : neg_lr_arg;
  for (int i0 = tid; i0 < n; i0 += 4 * kXThreads) {
    float p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int i = i0 + j * kXThreads; p[j] = param[e0 + (i < n ? i : n - 1)]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * kXThreads;
      if (i >= n) continue;
      const float s = sbar_s[chrel_s[i]];
      const float g = __fmul_rn((float)qsum_s[i], inv_world);             // all_reduce(SUM) * (1/N)
      const float u = __fmul_rn(__fmul_rn(neg_lr, g), s);                 // (-lr * grad) * s     (:642-643)
      param[e0 + i] = __fadd_rn(p[j], u);
      if (error_comp) error_comp[e0 + i] = __fsub_rn(g_s[i], __fmul_rn(g, s));   // weight - grad_up * s (:926-927)
    }
  }
  x_stamp(dbg, 5);
}

static unsigned long long* g_x_dbg = nullptr;

static long long x_timeout_cycles() {
  static const long long v = [] {
    const char* e = getenv("DQRM_P2P_TIMEOUT_S");
    const double sec = e ? atof(e) : 30.0;
    return sec <= 0.0 ? 0ll : (long long)(sec * 1.965e9);
  }();
  return v;
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_dense_exchange_debug(uint64_t* stamps) {
  g_x_dbg = reinterpret_cast<unsigned long long*>(stamps);
  return 0;
}

extern "C" size_t dqrm_dense_exchange_smem_bytes(int max_cta_elems, int max_cta_chans) {
  const size_t cap = ((size_t)max_cta_elems + 15) / 16 * 16;
  max_cta_chans = (max_cta_chans + 3) & ~3;                       // keeps every section 16-byte aligned
  return cap * 4 + (size_t)max_cta_chans * 12 + 16 + (cap + 32) * 2 + cap * 2 + (cap + 32) + 16;
}

extern "C" int dqrm_dense_exchange_apply(void* const* peer_base, int world, int rank, size_t scale_data_off,
                                         size_t scale_stride_bytes, size_t code_data_off, size_t code_stride_bytes,
                                         float* param, float* grad,
                                         float* error_comp, const int64_t* chan_begin, const int32_t* cta_chan,
                                         const int32_t* cta_word, int num_ctas, int max_cta_elems, int max_cta_chans, int bits, float* scale_mean,
                                         uint32_t* seq, float lr, const float* lr_dev, int32_t* status, void* stream) {
  DQRM_REQUIRE(peer_base && param && grad && chan_begin && cta_chan && cta_word && scale_mean && seq && status, -EINVAL,
               "dense_exchange_apply: null argument");
  DQRM_REQUIRE(world >= 1 && world <= kXMaxWorld && rank >= 0 && rank < world, -EINVAL,
               "dense_exchange_apply: rank %d / world %d", rank, world);
  DQRM_REQUIRE(bits >= 2 && bits <= 8, -EINVAL, "dense_exchange_apply: bits=%d outside [2,8] (int8 payload)", bits);
  DQRM_REQUIRE(num_ctas >= 1 && num_ctas <= kSMs, -EINVAL,
               "dense_exchange_apply: num_ctas=%d outside [1,%d] (every CTA must be resident: they wait for their peers)",
               num_ctas, kSMs);
  DQRM_REQUIRE(max_cta_elems >= 1 && max_cta_chans >= 1 && max_cta_chans < 65536, -EINVAL,
               "dense_exchange_apply: max_cta_elems=%d max_cta_chans=%d", max_cta_elems, max_cta_chans);
  DQRM_REQUIRE(((scale_data_off | scale_stride_bytes | code_data_off | code_stride_bytes) & 15u) == 0,
               -EINVAL, "dense_exchange_apply: site offsets / strides must be multiples of 16");
  XPeers pp;
  for (int r = 0; r < world; ++r) {
    DQRM_REQUIRE(peer_base[r] && (reinterpret_cast<uintptr_t>(peer_base[r]) & 15u) == 0, -EINVAL, "dense_exchange_apply: peer %d base", r);
    pp.base[r] = static_cast<unsigned char*>(peer_base[r]);
  }
  const size_t smem = dqrm_dense_exchange_smem_bytes(max_cta_elems, max_cta_chans);
  DQRM_REQUIRE(smem <= 200 * 1024, -E2BIG, "dense_exchange_apply: %zu bytes of shared memory per CTA (use more CTAs)", smem);
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && smem > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dense_exchange_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DQRM_REQUIRE(e == cudaSuccess, -EIO, "dense_exchange_apply: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    attr_set = smem;
  }
  XSites st{scale_data_off, scale_stride_bytes, code_data_off, code_stride_bytes};
  const int cap = (max_cta_elems + 15) / 16 * 16;
  dense_exchange_apply_kernel<<<num_ctas, kXThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      pp, st, world, rank, param, grad, error_comp, reinterpret_cast<const long long*>(chan_begin), cta_chan, cta_word, cap,
      (max_cta_chans + 3) & ~3, bits, (float)(1.0 / world), scale_mean, seq, -lr, lr_dev, x_timeout_cycles(), status, g_x_dbg);
  DQRM_LAUNCH_CHECK("dense_exchange_apply_kernel");
  return 0;
}
