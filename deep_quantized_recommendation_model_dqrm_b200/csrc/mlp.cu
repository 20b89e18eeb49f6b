// (a15) QuantLinear weight/bias fake-quantisation and (a11) the 8-bit per-channel quantised exchange
// of the dense (MLP) gradients.
// Reference: QuantLinear.forward, quantization_supp/quant_modules_not_quantize_grad.py:125-154
//            symmetric_linear_quantization_params, quantization_supp/quant_utils.py:196-220
//            quantize_linear_grad / quantize_bias_grad, sgd_quantized_gradients_parallel_comm.py:892-961
//            weight_update_parallel_comm (MLP part), sgd_quantized_gradients_parallel_comm.py:630-663
//
// Everything here is "one warp per quantisation channel" over a flat fp32 arena: a weight row is a
// channel, a whole bias vector is one channel.  All 14 MLP tensors of the model are handled by ONE
// launch per phase (scale / quantise / apply) instead of the reference's ~10 ATen launches and two
// Gloo all-reduces per tensor.  The GEMMs themselves stay on cuBLAS (fp32, library call).
#include "common.cuh"

namespace dqrm {

__global__ void __launch_bounds__(256)
linear_fakequant_kernel(const float* __restrict__ W, const float* __restrict__ b, int out_f, int in_f, int bits,
                        float* __restrict__ W_int, float* __restrict__ b_int, float* __restrict__ scale_row) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= out_f) return;
  const float* w = W + (long long)row * in_f;
  unsigned m = 0u;
  for (int i = lane; i < in_f; i += 32) m = max(m, abs_bits(w[i]));
  m = warp_max_u32(m);
  const float s = scale_of(__uint_as_float(m), bits);
  const float inv = __fdiv_rn(1.0f, s);
  const float hi = qmax_of(bits), lo = -hi - 1.0f;
  float* q = W_int + (long long)row * in_f;
  for (int i = lane; i < in_f; i += 32) q[i] = quant_code(w[i], inv, lo, hi);
  if (lane == 0) {
    scale_row[row] = s;
    if (b) b_int[row] = quant_code(b[row], inv, lo, hi);
  }
}

__global__ void __launch_bounds__(256)
dense_grad_scale_kernel(float* __restrict__ grad, const float* __restrict__ ec, const long long* __restrict__ chan_begin,
                        int num_chan, int bits, float* __restrict__ scale_local) {
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= num_chan) return;
  const long long a = chan_begin[ch], e = chan_begin[ch + 1];
  unsigned m = 0u;
  for (long long i0 = a + lane; i0 < e; i0 += 128) {                      // four elements in flight per lane
    long long idx[4];
    bool ok[4];
    float w[4], c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < 4; ++j) { idx[j] = i0 + 32 * j; ok[j] = idx[j] < e; idx[j] = ok[j] ? idx[j] : e - 1; }
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = grad[idx[j]];
    if (ec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = ec[idx[j]];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!ok[j]) continue;
      float v = w[j];
      if (ec) { v = __fadd_rn(v, c[j]); grad[idx[j]] = v; }                // weight = grad + error_compensation  (:899-900,938-939)
      m = max(m, abs_bits(v));
    }
  }
  m = warp_max_u32(m);
  if (lane == 0) scale_local[ch] = scale_of(__uint_as_float(m), bits);
}

__global__ void __launch_bounds__(256)
dense_grad_quant_kernel(const float* __restrict__ grad, const long long* __restrict__ chan_begin, int num_chan,
                        const float* __restrict__ scale_sum, float inv_world, int bits,
                        float* __restrict__ codes, float* __restrict__ scale_mean) {
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= num_chan) return;
  const long long a = chan_begin[ch], e = chan_begin[ch + 1];
  const float s_bar = __fmul_rn(scale_sum[ch], inv_world);
  const float inv = __fdiv_rn(1.0f, s_bar);
  const float hi = qmax_of(bits), lo = -hi - 1.0f;
#pragma unroll 4
  for (long long i = a + lane; i < e; i += 32) codes[i] = quant_code(grad[i], inv, lo, hi);
  if (lane == 0) scale_mean[ch] = s_bar;
}

__global__ void __launch_bounds__(256)
dense_apply_kernel(float* __restrict__ param, const float* __restrict__ code_sum, const long long* __restrict__ chan_begin,
                   int num_chan, const float* __restrict__ scale_mean, float inv_world, float neg_lr_arg,
                   const float* __restrict__ lr_dev, const float* __restrict__ comp_grad, float* __restrict__ ec_out) {
  const float neg_lr = lr_dev ? -(*lr_dev) : neg_lr_arg;
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= num_chan) return;
  const long long a = chan_begin[ch], e = chan_begin[ch + 1];
  const float s = scale_mean ? scale_mean[ch] : 1.0f;
  for (long long i0 = a + lane; i0 < e; i0 += 128) {                      // four elements in flight per lane (see p2p.cu)
    long long idx[4];
    bool ok[4];
    float cs[4], p[4], c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < 4; ++j) { idx[j] = i0 + 32 * j; ok[j] = idx[j] < e; idx[j] = ok[j] ? idx[j] : e - 1; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { cs[j] = code_sum[idx[j]]; p[j] = param[idx[j]]; }
    if (ec_out) {
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = comp_grad[idx[j]];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!ok[j]) continue;
      const float g = __fmul_rn(cs[j], inv_world);                        // all_reduce(SUM) * (1/N)
      float u = __fmul_rn(neg_lr, g);                                     // (-lr * grad) ...
      if (scale_mean) u = __fmul_rn(u, s);                                // ... * s          (:642-643)
      param[idx[j]] = __fadd_rn(p[j], u);
      if (ec_out) ec_out[idx[j]] = __fsub_rn(c[j], __fmul_rn(g, s));      // weight - grad_up * s   (:926-927,958-959)
    }
  }
}

// Single rank: dense_grad_scale + dense_grad_quant + dense_apply in ONE launch (a warp owns a channel from its
// max-abs to its updated parameters; the channel's second read hits L1/L2).  The same operations in the same order
// as the three kernels with inv_world = 1, and the same buffers written (scale_local, scale_mean, codes, grad when
// error compensation is on), so every observable result is bit-identical -- it only removes two launches from the
// tail of the step.  error_comp and error_comp_out may alias (the reference's buffer is read, then overwritten).
__global__ void __launch_bounds__(256)
dense_local_quant_apply_kernel(float* __restrict__ param, float* __restrict__ grad, const float* error_comp,
                               const long long* __restrict__ chan_begin, int num_chan, int bits,
                               float* __restrict__ scale_local, float* __restrict__ codes, float* __restrict__ scale_mean,
                               float neg_lr_arg, const float* __restrict__ lr_dev, float* error_comp_out) {
  const float neg_lr = lr_dev ? -(*lr_dev) : neg_lr_arg;
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= num_chan) return;
  const long long a = chan_begin[ch], e = chan_begin[ch + 1];
  unsigned m = 0u;
  for (long long i0 = a + lane; i0 < e; i0 += 128) {
    long long idx[4];
    bool ok[4];
    float w[4], c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < 4; ++j) { idx[j] = i0 + 32 * j; ok[j] = idx[j] < e; idx[j] = ok[j] ? idx[j] : e - 1; }
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = grad[idx[j]];
    if (error_comp) {
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = error_comp[idx[j]];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!ok[j]) continue;
      float v = w[j];
      if (error_comp) { v = __fadd_rn(v, c[j]); grad[idx[j]] = v; }
      m = max(m, abs_bits(v));
    }
  }
  m = warp_max_u32(m);
  const float s_local = scale_of(__uint_as_float(m), bits);
  const float s_bar = __fmul_rn(s_local, 1.0f);                           // scale_sum * inv_world
  const float inv = __fdiv_rn(1.0f, s_bar);
  const float hi = qmax_of(bits), lo = -hi - 1.0f;
  if (lane == 0) { scale_local[ch] = s_local; scale_mean[ch] = s_bar; }
  for (long long i0 = a + lane; i0 < e; i0 += 128) {                      // each lane re-reads what it wrote above
    long long idx[4];
    bool ok[4];
    float w[4], p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { idx[j] = i0 + 32 * j; ok[j] = idx[j] < e; idx[j] = ok[j] ? idx[j] : e - 1; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { w[j] = grad[idx[j]]; p[j] = param[idx[j]]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!ok[j]) continue;
      const float q = quant_code(w[j], inv, lo, hi);
      codes[idx[j]] = q;
      const float g = __fmul_rn(q, 1.0f);                                 // code_sum * inv_world
      const float u = __fmul_rn(__fmul_rn(neg_lr, g), s_bar);
      param[idx[j]] = __fadd_rn(p[j], u);
      if (error_comp_out) error_comp_out[idx[j]] = __fsub_rn(w[j], __fmul_rn(g, s_bar));
    }
  }
}

__global__ void __launch_bounds__(256)
fake_quant_kernel(const float* __restrict__ x, long long rows, long long cols, const float* __restrict__ scale,
                  int per_row, int bits, float* __restrict__ q, float* __restrict__ dq) {
  const float hi = qmax_of(bits), lo = -hi - 1.0f;
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = per_row ? scale[i / cols] : scale[0];
    const float c = quant_code(x[i], __fdiv_rn(1.0f, s), lo, hi);
    q[i] = c;
    if (dq) dq[i] = __fmul_rn(c, s);
  }
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_fake_quant(const float* x, int64_t rows, int64_t cols, const float* scale, int scale_per_row,
                               int bits, float* q, float* dequant, void* stream) {
  DQRM_REQUIRE(x && scale && q && rows >= 0 && cols >= 0, -EINVAL, "fake_quant: bad argument");
  DQRM_REQUIRE(bits >= 2 && bits <= 16, -EINVAL, "fake_quant: bits=%d outside [2,16]", bits);
  if (rows * cols == 0) return 0;
  long long blocks = ceil_div(rows * cols, 256);
  if (blocks > 8 * kSMs) blocks = 8 * kSMs;
  fake_quant_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, cols, scale, scale_per_row,
                                                                                     bits, q, dequant);
  DQRM_LAUNCH_CHECK("fake_quant_kernel");
  return 0;
}

extern "C" int dqrm_linear_fakequant(const float* W, const float* b, int out_features, int in_features, int bits,
                                     float* W_int, float* b_int, float* scale_row, void* stream) {
  DQRM_REQUIRE(W && W_int && scale_row, -EINVAL, "linear_fakequant: null argument");
  DQRM_REQUIRE((b == nullptr) == (b_int == nullptr), -EINVAL, "linear_fakequant: b and b_int must come together");
  DQRM_REQUIRE(out_features >= 1 && in_features >= 1, -EINVAL, "linear_fakequant: bad shape");
  DQRM_REQUIRE(bits >= 2 && bits <= 16, -EINVAL, "linear_fakequant: bits=%d outside [2,16]", bits);
  linear_fakequant_kernel<<<(out_features + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, b, out_features, in_features, bits, W_int, b_int, scale_row);
  DQRM_LAUNCH_CHECK("linear_fakequant_kernel");
  return 0;
}

extern "C" int dqrm_dense_grad_scale(float* grad, const float* error_comp, const int64_t* chan_begin, int num_chan, int bits,
                                     float* scale_local, void* stream) {
  DQRM_REQUIRE(grad && chan_begin && scale_local && num_chan >= 1, -EINVAL, "dense_grad_scale: bad argument");
  DQRM_REQUIRE(bits >= 2 && bits <= 16, -EINVAL, "dense_grad_scale: bits=%d outside [2,16]", bits);
  dense_grad_scale_kernel<<<(num_chan + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad, error_comp, reinterpret_cast<const long long*>(chan_begin), num_chan, bits, scale_local);
  DQRM_LAUNCH_CHECK("dense_grad_scale_kernel");
  return 0;
}

extern "C" int dqrm_dense_grad_quant(const float* grad, const int64_t* chan_begin, int num_chan,
                                     const float* scale_sum, float inv_world, int bits,
                                     float* codes, float* scale_mean, void* stream) {
  DQRM_REQUIRE(grad && chan_begin && scale_sum && codes && scale_mean && num_chan >= 1, -EINVAL, "dense_grad_quant: bad argument");
  DQRM_REQUIRE(bits >= 2 && bits <= 16, -EINVAL, "dense_grad_quant: bits=%d outside [2,16]", bits);
  dense_grad_quant_kernel<<<(num_chan + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad, reinterpret_cast<const long long*>(chan_begin), num_chan, scale_sum, inv_world, bits, codes, scale_mean);
  DQRM_LAUNCH_CHECK("dense_grad_quant_kernel");
  return 0;
}

extern "C" int dqrm_dense_apply(float* param, const float* code_sum, const int64_t* chan_begin, int num_chan,
                                const float* scale_mean, float inv_world, float lr, const float* lr_dev, const float* comp_grad,
                                float* error_comp_out, void* stream) {
  DQRM_REQUIRE(param && code_sum && chan_begin && num_chan >= 1, -EINVAL, "dense_apply: bad argument");
  DQRM_REQUIRE(!error_comp_out || (comp_grad && scale_mean), -EINVAL, "dense_apply: error compensation needs comp_grad and scale_mean");
  dense_apply_kernel<<<(num_chan + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, code_sum, reinterpret_cast<const long long*>(chan_begin), num_chan, scale_mean, inv_world, -lr, lr_dev,
      comp_grad, error_comp_out);
  DQRM_LAUNCH_CHECK("dense_apply_kernel");
  return 0;
}

extern "C" int dqrm_dense_quant_apply_local(float* param, float* grad, float* error_comp, const int64_t* chan_begin,
                                            int num_chan, int bits, float* scale_local, float* codes, float* scale_mean,
                                            float lr, const float* lr_dev, void* stream) {
  DQRM_REQUIRE(param && grad && chan_begin && scale_local && codes && scale_mean && num_chan >= 1, -EINVAL,
               "dense_quant_apply_local: bad argument");
  DQRM_REQUIRE(bits >= 2 && bits <= 16, -EINVAL, "dense_quant_apply_local: bits=%d outside [2,16]", bits);
  dense_local_quant_apply_kernel<<<(num_chan + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, error_comp, reinterpret_cast<const long long*>(chan_begin), num_chan, bits, scale_local, codes,
      scale_mean, -lr, lr_dev, error_comp);
  DQRM_LAUNCH_CHECK("dense_local_quant_apply_kernel");
  return 0;
}

// ---- BCE loss (mean) and its gradient in one launch ------------------------------------------------------------
// Reference: loss_fn_wrap -> torch.nn.BCELoss(reduction="mean") (dlrm_s_pytorch_comm_grad.py:192-211) followed by
// E.backward() (:1938): ATen runs 8 small kernels for it (bce forward, mean, ones_like, fills, bce backward, /N).
//   loss = mean( (t - 1) * max(log1p(-z), -100) - t * max(log(z), -100) )
//   dz   = ((z - t) / max((1 - z) * z, 1e-12)) * (1/N)                 (the value autograd hands to the sigmoid;
//          ATen: bce backward with grad 1, then the mean's division as a multiply by the fp32 reciprocal)
// One CTA, fixed summation order (deterministic); the loss agrees with ATen's tree reduction to fp32 rounding,
// the gradient is evaluated in ATen's operation order.
namespace dqrm {
__global__ void __launch_bounds__(1024)
bce_loss_grad_kernel(const float* __restrict__ z, const float* __restrict__ t, long long n, float inv_n,
                     float* __restrict__ loss, float* __restrict__ dz) {
  __shared__ float s_part[32];
  float acc = 0.0f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float zi = z[i], ti = t[i];
    const float l1 = fmaxf(log1pf(-zi), -100.0f), l0 = fmaxf(logf(zi), -100.0f);
    acc = __fadd_rn(acc, __fsub_rn(__fmul_rn(__fsub_rn(ti, 1.0f), l1), __fmul_rn(ti, l0)));
    if (dz) dz[i] = __fmul_rn(__fdiv_rn(__fsub_rn(zi, ti), fmaxf(__fmul_rn(__fsub_rn(1.0f, zi), zi), 1e-12f)), inv_n);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, o));
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) *loss = __fmul_rn(v, inv_n);
  }
}
}  // namespace dqrm

extern "C" int dqrm_bce_loss_grad(const float* z, const float* target, int64_t n, float* loss, float* dz, void* stream) {
  DQRM_REQUIRE(z && target && loss && n >= 1, -EINVAL, "bce_loss_grad: bad argument");
  const int threads = n >= 1024 ? 1024 : (int)((n + 31) / 32 * 32);
  dqrm::bce_loss_grad_kernel<<<1, threads, 0, static_cast<cudaStream_t>(stream)>>>(z, target, n, (float)(1.0 / (double)n),
                                                                                  loss, dz);
  DQRM_LAUNCH_CHECK("bce_loss_grad_kernel");
  return 0;
}
