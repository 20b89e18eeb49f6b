// (a7-3/5, a11, a1-sharded) the step's exchanges as ONE-KERNEL all-gathers over NVLink peer memory.
//
// Reference: the per-table / per-tensor Gloo collectives of quantize_emb_grad, quantize_linear_grad and
// quantize_bias_grad (sgd_quantized_gradients_parallel_comm.py:865,878,913,925,949,957) -- 80 host-staged
// collectives per step.  The NCCL form of this library needs 5 (two all-gathers, three all-reduces); every one
// moves 0.1-500 kB, i.e. it is pure launch + protocol latency (15-40 us each inside a CUDA graph, 8 GPUs).
//
// Here every rank owns one "peer arena" (cudaMalloc + CUDA IPC, mapped by all ranks of the box, NVLink/NVSwitch).
// An exchange site is  ctl[4] | flag[world] | slot[world][slot_bytes]  at the same offsets in every arena.
// The producer kernel of the step writes this rank's contribution into slot[rank] of its OWN arena; then
//   p2p_allgather_kernel:  copy slot[rank] into slot[rank] of every peer's arena with 128-bit remote stores,
//                          fence.sys, last CTA: st.release.sys flag[rank] = seq on every peer, then spin
//                          (ld.acquire.sys) until all local flag[p] >= seq.
// When the kernel retires, slot[0..world) of the local arena holds every rank's data and the consumer kernel
// (pack / merge / quantise / apply / scale) reads it in rank order -- so sums are deterministic and identical on
// all ranks, and the all-reduces become "all-gather + reduce in the consumer" (integer codes travel as int8: 4x
// fewer bytes than the fp32 code all-reduce).  seq comes from a device-side counter, so the launch is CUDA-graph
// replayable with fixed arguments.  Single buffering is safe because a step has >= 2 sites visited in a fixed
// order on one stream: a peer's flag for site B(s) is stored after its consumer of site A(s) has retired, and
// this rank overwrites A only after it has seen B(s) from every peer.
//
// A rank that never shows up would spin forever; the wait gives up after DQRM_P2P_TIMEOUT_S seconds (default 30;
// 0 = wait for ever, like NCCL would) and sets DQRM_STATUS_P2P_TIMEOUT.  That bit is FATAL and sticky: the slots may
// be stale or half written and the single-buffering argument above no longer holds, so the consumers that would
// apply them to the weights (grad_merge_apply, dense_apply_gathered) turn into no-ops while it is set, and the host
// raises at its next status poll (EmbeddingTableGroup.check_status / DenseArena.check_status never clear it).
#include <stdlib.h>

#include "common.cuh"

namespace dqrm {

constexpr int kP2PThreads = 256;
constexpr int kP2PMaxWorld = 16;
static long long p2p_timeout_cycles() {                    // env DQRM_P2P_TIMEOUT_S, default 30 s at ~1.97 GHz; 0 = never
  static const long long v = [] {
    const char* e = getenv("DQRM_P2P_TIMEOUT_S");
    const double sec = e ? atof(e) : 30.0;
    return sec <= 0.0 ? 0ll : (long long)(sec * 1.965e9);
  }();
  return v;
}

struct PeerPtrs { unsigned char* base[kP2PMaxWorld]; };

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ctl = {counter, arrive, -, -} (local), flags = world x u32 (written by peers), data = world x slot_bytes
__global__ void __launch_bounds__(kP2PThreads)
p2p_allgather_kernel(const __grid_constant__ PeerPtrs peers, int world, int rank, size_t ctl_off, size_t flag_off,
                     size_t data_off, size_t slot_bytes, long long timeout_cycles, int* __restrict__ status) {
  unsigned char* local = peers.base[rank];
  unsigned* ctl = reinterpret_cast<unsigned*>(local + ctl_off);
  const unsigned seq = ctl[0] + 1u;          // every CTA reads it before the last CTA (below) bumps it
  const size_t my_off = data_off + (size_t)rank * slot_bytes;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  const uint4* src = reinterpret_cast<const uint4*>(local + my_off);        // slot_bytes is a multiple of 16
  const size_t n16 = slot_bytes >> 4;
  for (size_t i0 = tid; i0 < n16; i0 += 4 * nthr) {        // four 128-bit loads in flight, then the remote stores
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const size_t i = i0 + j * nthr; v[j] = src[i < n16 ? i : n16 - 1]; }
    for (int k = 1; k < world; ++k) {
      const int p = (rank + k) % world;      // every rank starts on a different peer
      uint4* dst = reinterpret_cast<uint4*>(peers.base[p] + my_off);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const size_t i = i0 + j * nthr; if (i < n16) dst[i] = v[j]; }
    }
  }
  // One system-scope fence per CTA, not per thread: the barrier orders every thread's remote stores before thread 0's
  // fence, which is cumulative (the pattern of a cooperative grid barrier); 256 threads x fence.sys cost 2-4 us a piece.
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    s_last = true;
    if (gridDim.x > 1) {
      __threadfence_system();                // release side of the arrive counter, in the thread that bumps it
      s_last = (atomicAdd(&ctl[1], 1u) == gridDim.x - 1);
      __threadfence();                       // acquire side: the other CTAs' fences happened before their arrivals
    }
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < world && threadIdx.x != rank) {
    st_release_sys(reinterpret_cast<unsigned*>(peers.base[threadIdx.x] + flag_off) + rank, seq);   // (release: cumulative)
    const unsigned* f = reinterpret_cast<const unsigned*>(local + flag_off) + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_relaxed_sys(f) - seq) < 0) {
      if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) { atomicOr(status, DQRM_STATUS_P2P_TIMEOUT); break; }
    }
    (void)ld_acquire_sys(f);
  }
  __syncthreads();
  if (threadIdx.x == 0) { ctl[1] = 0u; ctl[0] = seq; }
}

// ---- consumers that reduce the gathered slots in rank order ------------------------------------------------

// (a11) s_bar = (sum_r s_r) * (1/N) in rank order; q = clamp(rint((1/s_bar) * g)) as int8 into this rank's slot
__global__ void __launch_bounds__(256)
dense_grad_quant_gathered_kernel(const float* __restrict__ grad, const long long* __restrict__ chan_begin, int num_chan,
                                 const float* __restrict__ gathered_scales, size_t scale_stride, int world,
                                 float inv_world, int bits, signed char* __restrict__ codes,
                                 float* __restrict__ scale_mean) {
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= num_chan) return;
  const long long a = chan_begin[ch], e = chan_begin[ch + 1];
  float acc = gathered_scales[ch];
  for (int r = 1; r < world; ++r) acc = __fadd_rn(acc, gathered_scales[(size_t)r * scale_stride + ch]);
  const float s_bar = __fmul_rn(acc, inv_world);
  const float inv = __fdiv_rn(1.0f, s_bar);
  const float hi = qmax_of(bits), lo = -hi - 1.0f;
#pragma unroll 4
  for (long long i = a + lane; i < e; i += 32) codes[i] = (signed char)quant_code(grad[i], inv, lo, hi);
  if (lane == 0) scale_mean[ch] = s_bar;
}

// (a11 + a9 MLP part) W += (-lr * ((sum_r q_r) * (1/N))) * s   with the integer codes summed exactly
__global__ void __launch_bounds__(256)
dense_apply_gathered_kernel(float* __restrict__ param, const signed char* __restrict__ gathered_codes,
                            size_t code_stride, int world, const long long* __restrict__ chan_begin, int num_chan,
                            const float* __restrict__ scale_mean, float inv_world, float neg_lr_arg,
                            const float* __restrict__ lr_dev, const float* __restrict__ comp_grad,
                            float* __restrict__ ec_out, const int* __restrict__ status) {
  if (status && (*status & DQRM_STATUS_P2P_TIMEOUT)) return;             // an exchange timed out: never apply stale slots
  const float neg_lr = lr_dev ? -(*lr_dev) : neg_lr_arg;
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= num_chan) return;
  const long long a = chan_begin[ch], e = chan_begin[ch + 1];
  const float s = scale_mean[ch];
  // four strided elements per lane and trip, every load (world codes + the parameter + comp_grad) issued before the
  // first store: with one element per trip the row cost one dependent load round trip per 32 elements (14 us at
  // world 2 for the 0.47 M MLP parameters of the Kaggle shape).  Out-of-range lanes read a clamped address.
  for (long long i0 = a + lane; i0 < e; i0 += 128) {
    long long idx[4];
    bool ok[4];
    int q[4] = {0, 0, 0, 0};
    float p[4], c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < 4; ++j) { idx[j] = i0 + 32 * j; ok[j] = idx[j] < e; idx[j] = ok[j] ? idx[j] : e - 1; }
#pragma unroll 2
    for (int r = 0; r < world; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) q[j] += gathered_codes[(size_t)r * code_stride + idx[j]];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = param[idx[j]];
    if (ec_out) {
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = comp_grad[idx[j]];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!ok[j]) continue;
      const float g = __fmul_rn((float)q[j], inv_world);                  // all_reduce(SUM) * (1/N)
      const float u = __fmul_rn(__fmul_rn(neg_lr, g), s);                 // (-lr * grad) * s     (:642-643)
      param[idx[j]] = __fadd_rn(p[j], u);
      if (ec_out) ec_out[idx[j]] = __fsub_rn(c[j], __fmul_rn(g, s));      // weight - grad_up * s   (:926-927,958-959)
    }
  }
}

// (a1, row-sharded) absmax = max_r absmax_r ; s = max(absmax,1e-8)/n ; inv = 1/s
__global__ void scale_from_absmax_gathered_kernel(int n, const float* __restrict__ gathered, size_t stride, int world,
                                                  int bits, float* __restrict__ absmax, float* __restrict__ scale,
                                                  float* __restrict__ inv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float m = gathered[i];
  for (int r = 1; r < world; ++r) m = fmaxf(m, gathered[(size_t)r * stride + i]);
  const float s = scale_of(m, bits);
  absmax[i] = m;
  scale[i] = s;
  inv[i] = __fdiv_rn(1.0f, s);
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_p2p_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_64) {
  DQRM_REQUIRE(bytes > 0 && dev_ptr && ipc_handle_64, -EINVAL, "p2p_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  DQRM_REQUIRE(e == cudaSuccess, -ENOMEM, "p2p_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(ipc_handle_64), p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("p2p_alloc: %s", cudaGetErrorString(e));
    return -EIO;
  }
  *dev_ptr = p;
  return 0;
}

extern "C" int dqrm_p2p_open(const void* ipc_handle_64, void** dev_ptr) {
  DQRM_REQUIRE(ipc_handle_64 && dev_ptr, -EINVAL, "p2p_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  DQRM_REQUIRE(e == cudaSuccess, -EIO, "p2p_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int dqrm_p2p_close(void* dev_ptr) {
  if (!dev_ptr) return 0;
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  DQRM_REQUIRE(e == cudaSuccess, -EIO, "p2p_close: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int dqrm_p2p_free(void* dev_ptr) {
  if (!dev_ptr) return 0;
  cudaError_t e = cudaFree(dev_ptr);
  DQRM_REQUIRE(e == cudaSuccess, -EIO, "p2p_free: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" size_t dqrm_p2p_site_bytes(int world, size_t slot_bytes) {
  // ctl (16 B) | flags (world x 4, padded to 16) | slots (each padded to 16)
  const size_t flags = ((size_t)world * 4 + 15) & ~(size_t)15;
  const size_t slot = (slot_bytes + 15) & ~(size_t)15;
  return 16 + flags + (size_t)world * slot;
}

extern "C" int dqrm_p2p_site_layout(int world, size_t slot_bytes, size_t* flag_off, size_t* data_off, size_t* slot_stride) {
  DQRM_REQUIRE(world >= 1 && flag_off && data_off && slot_stride, -EINVAL, "p2p_site_layout: bad argument");
  *flag_off = 16;
  *data_off = 16 + (((size_t)world * 4 + 15) & ~(size_t)15);
  *slot_stride = (slot_bytes + 15) & ~(size_t)15;
  return 0;
}

extern "C" int dqrm_p2p_allgather(void* const* peer_base, int world, int rank, size_t site_off, size_t slot_bytes,
                                  int32_t* status, void* stream) {
  DQRM_REQUIRE(peer_base && status, -EINVAL, "p2p_allgather: null argument");
  DQRM_REQUIRE(world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, -EINVAL, "p2p_allgather: rank %d / world %d",
               rank, world);
  DQRM_REQUIRE(slot_bytes >= 1 && (site_off & 15u) == 0, -EINVAL,
               "p2p_allgather: slot_bytes=%zu site_off=%zu (multiple of 16)", slot_bytes, site_off);
  if (world == 1) return 0;
  PeerPtrs pp;
  for (int r = 0; r < world; ++r) {
    DQRM_REQUIRE(peer_base[r] && (reinterpret_cast<uintptr_t>(peer_base[r]) & 15u) == 0, -EINVAL, "p2p_allgather: peer %d base", r);
    pp.base[r] = static_cast<unsigned char*>(peer_base[r]);
  }
  size_t flag_off, data_off, stride;
  dqrm_p2p_site_layout(world, slot_bytes, &flag_off, &data_off, &stride);
  // one 16-byte chunk per thread while the SMs last: remote stores drain slowly per SM (66 kB to seven peers took 18 us
  // from 5 CTAs), so spread them over as many SMs as there are chunks
  long long grid = ceil_div((long long)stride, 16 * kP2PThreads);
  if (grid > kSMs) grid = kSMs;
  if (grid < 1) grid = 1;
  p2p_allgather_kernel<<<(unsigned)grid, kP2PThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      pp, world, rank, site_off, site_off + flag_off, site_off + data_off, stride, p2p_timeout_cycles(), status);
  DQRM_LAUNCH_CHECK("p2p_allgather_kernel");
  return 0;
}

extern "C" int dqrm_dense_grad_quant_gathered(const float* grad, const int64_t* chan_begin, int num_chan,
                                              const float* gathered_scales, size_t scale_stride_elems, int world,
                                              int bits, int8_t* codes, float* scale_mean, void* stream) {
  DQRM_REQUIRE(grad && chan_begin && gathered_scales && codes && scale_mean && num_chan >= 1 && world >= 1, -EINVAL,
               "dense_grad_quant_gathered: bad argument");
  DQRM_REQUIRE(bits >= 2 && bits <= 8, -EINVAL, "dense_grad_quant_gathered: bits=%d outside [2,8] (int8 payload)", bits);
  dense_grad_quant_gathered_kernel<<<(num_chan + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad, reinterpret_cast<const long long*>(chan_begin), num_chan, gathered_scales, scale_stride_elems, world,
      (float)(1.0 / world), bits, reinterpret_cast<signed char*>(codes), scale_mean);
  DQRM_LAUNCH_CHECK("dense_grad_quant_gathered_kernel");
  return 0;
}

extern "C" int dqrm_dense_apply_gathered(float* param, const int8_t* gathered_codes, size_t code_stride_bytes, int world,
                                         const int64_t* chan_begin, int num_chan, const float* scale_mean, float lr,
                                         const float* lr_dev, const float* comp_grad, float* error_comp_out, const int32_t* status,
                                         void* stream) {
  DQRM_REQUIRE(param && gathered_codes && chan_begin && scale_mean && num_chan >= 1 && world >= 1, -EINVAL,
               "dense_apply_gathered: bad argument");
  DQRM_REQUIRE(!error_comp_out || comp_grad, -EINVAL, "dense_apply_gathered: error compensation needs comp_grad");
  dense_apply_gathered_kernel<<<(num_chan + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, reinterpret_cast<const signed char*>(gathered_codes), code_stride_bytes, world,
      reinterpret_cast<const long long*>(chan_begin), num_chan, scale_mean, (float)(1.0 / world), -lr, lr_dev, comp_grad,
      error_comp_out, status);
  DQRM_LAUNCH_CHECK("dense_apply_gathered_kernel");
  return 0;
}

extern "C" int dqrm_scale_from_absmax_gathered(int n_scales, const float* gathered_absmax, size_t stride_elems, int world,
                                               int bits, float* absmax, float* scale, float* inv_scale, void* stream) {
  DQRM_REQUIRE(n_scales >= 1 && gathered_absmax && absmax && scale && inv_scale && world >= 1, -EINVAL,
               "scale_from_absmax_gathered: bad argument");
  DQRM_REQUIRE(bits >= 2 && bits <= 16, -EINVAL, "scale_from_absmax_gathered: bits=%d outside [2,16]", bits);
  scale_from_absmax_gathered_kernel<<<(n_scales + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      n_scales, gathered_absmax, stride_elems, world, bits, absmax, scale, inv_scale);
  DQRM_LAUNCH_CHECK("scale_from_absmax_gathered_kernel");
  return 0;
}
