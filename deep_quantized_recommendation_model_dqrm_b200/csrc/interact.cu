// (a14) fused dot interaction: cat + T.T^t + lower-triangle gather + cat, forward and backward.
// Reference: DLRM_Net.interact_features, dlrm_s_pytorch_comm_grad.py:701-725 (bmm :710, li/lj index
// lists rebuilt on the host each call :719-722, gather :723, cat :725).
//
// One CTA (4 warps) per sample.  The (F+1) x D feature tile is staged once in shared memory (row stride
// D+4 so consecutive rows start on different banks); each thread then owns pairs p = tid, tid+128, ...
// and writes its dot products straight into R next to the copied dense features: no [B,27,27] matrix,
// no index tensors, one launch.  27x27x16 per sample is far below a tcgen05 tile (and fp32-exact
// parity rules out TF32), so this is FFMA work bounded by launch latency, not by a roofline.
#include "common.cuh"

namespace dqrm {

__device__ __forceinline__ int pair_index(int i, int j, int off) { return i * (i - 1) / 2 + off * i + j; }

constexpr int kInteractThreads = 128;

// cooperative tile load by the whole CTA: row 0 = x[b], rows 1..F = ly[k][b]
__device__ __forceinline__ void load_tile(float* tile, int stride, const float* __restrict__ x,
                                          const float* __restrict__ ly, long long lts, long long lbs,
                                          long long b, int F, int dim4) {
  const int total = (F + 1) * dim4;
  for (int it = threadIdx.x; it < total; it += kInteractThreads) {
    const int i = it / dim4, c = it - i * dim4;
    const float* src = (i == 0) ? x + b * dim4 * 4 : ly + (long long)(i - 1) * lts + b * lbs;
    *reinterpret_cast<float4*>(tile + i * stride + c * 4) = __ldg(reinterpret_cast<const float4*>(src) + c);
  }
}

// One CTA (4 warps) per sample: at batch 128 that is 128 CTAs, i.e. the whole chip, instead of 16 CTAs of
// warp-per-sample; every phase is a single round of independent loads.
__global__ void __launch_bounds__(kInteractThreads)
interact_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ly, long long lts, long long lbs,
                    long long batch, int F, int dim4, int off, int npairs, float* __restrict__ R) {
  extern __shared__ __align__(16) float smem_f[];
  const int dim = dim4 * 4, stride = dim + 4, nf = F + 1;
  float* tile = smem_f;
  const long long out_w = dim + npairs;
  for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
    __syncthreads();
    load_tile(tile, stride, x, ly, lts, lbs, b, F, dim4);
    __syncthreads();
    float* r = R + b * out_w;
    for (int d = threadIdx.x; d < dim; d += kInteractThreads) r[d] = tile[d];
    // rows i = 1 (or 0 with `itself`) .. nf-1; pair (i, j), j < i + off, lives at pair_index(i, j)
    for (int p = threadIdx.x; p < npairs; p += kInteractThreads) {
      // invert p -> (i, j): i is the largest row with pair_index(i, 0) <= p
      int i = (int)((sqrtf(8.0f * (float)p + 1.0f) + 1.0f) * 0.5f);
      if (off) i = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
      while (pair_index(i + 1, 0, off) <= p) ++i;
      while (pair_index(i, 0, off) > p) --i;
      const int j = p - pair_index(i, 0, off);
      const float4* a = reinterpret_cast<const float4*>(tile + i * stride);
      const float4* c = reinterpret_cast<const float4*>(tile + j * stride);
      float acc = 0.f;
      for (int k = 0; k < dim4; ++k) {
        const float4 u = a[k], v = c[k];
        acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc); acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
      }
      r[dim + p] = acc;
    }
  }
}

// Backward.  `ste_scale` (optional, dev [F]) fuses the straight-through estimator of the QAT EmbeddingBag
// into the epilogue: dly <- (dly * s_k) / s_k  (autograd of qm:393 followed by quant_utils.py:363), so the
// de-duplicating backward can consume it without dividing on its serial fold path.
__global__ void __launch_bounds__(kInteractThreads)
interact_bwd_kernel(const float* __restrict__ x, const float* __restrict__ ly, long long lts, long long lbs,
                    const float* __restrict__ dR, long long batch, int F, int dim4, int off, int npairs,
                    float* __restrict__ dx, float* __restrict__ dly, long long dts, long long dbs,
                    const float* __restrict__ ste_scale) {
  extern __shared__ __align__(16) float smem_f[];
  const int dim = dim4 * 4, stride = dim + 4, nf = F + 1;
  float* tile = smem_f;
  float* dz = tile + nf * stride;
  const long long in_w = dim + npairs;
  for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
    __syncthreads();
    load_tile(tile, stride, x, ly, lts, lbs, b, F, dim4);
    const float* g = dR + b * in_w;
    for (int p = threadIdx.x; p < npairs; p += kInteractThreads) dz[p] = __ldg(g + dim + p);
    __syncthreads();
    for (int it = threadIdx.x; it < nf * dim4; it += kInteractThreads) {
      const int i = it / dim4, c = it - i * dim4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* tcol = tile + c * 4;
      const int rowbase = pair_index(i, 0, off);
      for (int j = 0; j < i; ++j) {                                     // dZ[i][j], j < i
        const float w = dz[rowbase + j];
        const float4 v = *reinterpret_cast<const float4*>(tcol + j * stride);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      if (off) {                                                        // diagonal: d(T_i . T_i) = 2 T_i
        const float w = 2.0f * dz[rowbase + i];
        const float4 v = *reinterpret_cast<const float4*>(tcol + i * stride);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      for (int j = i + 1; j < nf; ++j) {                                // dZ[j][i], j > i
        const float w = dz[pair_index(j, i, off)];
        const float4 v = *reinterpret_cast<const float4*>(tcol + j * stride);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      if (i == 0) {
        acc.x += __ldg(g + c * 4); acc.y += __ldg(g + c * 4 + 1); acc.z += __ldg(g + c * 4 + 2); acc.w += __ldg(g + c * 4 + 3);
        reinterpret_cast<float4*>(dx + b * dim)[c] = acc;
      } else {
        if (ste_scale) {
          const float s = __ldg(ste_scale + (i - 1));
          acc.x = __fdiv_rn(__fmul_rn(acc.x, s), s); acc.y = __fdiv_rn(__fmul_rn(acc.y, s), s);
          acc.z = __fdiv_rn(__fmul_rn(acc.z, s), s); acc.w = __fdiv_rn(__fmul_rn(acc.w, s), s);
        }
        reinterpret_cast<float4*>(dly + (long long)(i - 1) * dts + b * dbs)[c] = acc;
      }
    }
  }
}

static int interact_cfg(int F, int dim, int itself, int* npairs, size_t* smem_fwd, size_t* smem_bwd) {
  const int nf = F + 1, off = itself ? 1 : 0;
  *npairs = nf * (nf - 1) / 2 + off * nf;
  const size_t tile = (size_t)nf * (dim + 4) * sizeof(float);
  *smem_fwd = tile;
  *smem_bwd = tile + (size_t)((*npairs + 3) & ~3) * sizeof(float);
  return (*smem_bwd <= 227 * 1024) ? 0 : -E2BIG;
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_interact_fwd(const float* x, const float* ly, int64_t ly_table_stride, int64_t ly_bag_stride,
                                 int64_t batch, int num_tables, int dim, int itself, float* R, void* stream) {
  DQRM_REQUIRE(x && ly && R, -EINVAL, "interact_fwd: null argument");
  DQRM_REQUIRE(num_tables >= 1 && num_tables < 256, -EINVAL, "interact_fwd: num_tables=%d", num_tables);
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "interact_fwd: dim=%d", dim);
  DQRM_REQUIRE(ly_table_stride % 4 == 0 && ly_bag_stride % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(ly)) & 15u) == 0,
               -EINVAL, "interact_fwd: inputs must be 16-byte aligned with strides multiple of 4");
  if (batch <= 0) return 0;
  int npairs;
  size_t sf, sb;
  DQRM_REQUIRE(interact_cfg(num_tables, dim, itself, &npairs, &sf, &sb) == 0, -E2BIG,
               "interact_fwd: feature tile does not fit shared memory");
  if (sf > 48 * 1024) cudaFuncSetAttribute(interact_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sf);
  long long grid = batch < 16ll * kSMs ? batch : 16ll * kSMs;
  interact_fwd_kernel<<<(unsigned)grid, kInteractThreads, sf, static_cast<cudaStream_t>(stream)>>>(
      x, ly, ly_table_stride, ly_bag_stride, batch, num_tables, dim / 4, itself ? 1 : 0, npairs, R);
  DQRM_LAUNCH_CHECK("interact_fwd_kernel");
  return 0;
}

extern "C" int dqrm_interact_bwd(const float* x, const float* ly, int64_t ly_table_stride, int64_t ly_bag_stride,
                                 const float* dR, int64_t batch, int num_tables, int dim, int itself,
                                 float* dx, float* dly, int64_t dly_table_stride, int64_t dly_bag_stride,
                                 const float* ste_scale, void* stream) {
  DQRM_REQUIRE(x && ly && dR && dx && dly, -EINVAL, "interact_bwd: null argument");
  DQRM_REQUIRE(num_tables >= 1 && num_tables < 256, -EINVAL, "interact_bwd: num_tables=%d", num_tables);
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "interact_bwd: dim=%d", dim);
  DQRM_REQUIRE(ly_table_stride % 4 == 0 && ly_bag_stride % 4 == 0 && dly_table_stride % 4 == 0 && dly_bag_stride % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(ly) | reinterpret_cast<uintptr_t>(dx) |
                     reinterpret_cast<uintptr_t>(dly)) & 15u) == 0,
               -EINVAL, "interact_bwd: buffers must be 16-byte aligned with strides multiple of 4");
  if (batch <= 0) return 0;
  int npairs;
  size_t sf, sb;
  DQRM_REQUIRE(interact_cfg(num_tables, dim, itself, &npairs, &sf, &sb) == 0, -E2BIG,
               "interact_bwd: feature tile does not fit shared memory");
  if (sb > 48 * 1024) cudaFuncSetAttribute(interact_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
  long long grid = batch < 16ll * kSMs ? batch : 16ll * kSMs;
  interact_bwd_kernel<<<(unsigned)grid, kInteractThreads, sb, static_cast<cudaStream_t>(stream)>>>(
      x, ly, ly_table_stride, ly_bag_stride, dR, batch, num_tables, dim / 4, itself ? 1 : 0, npairs, dx, dly,
      dly_table_stride, dly_bag_stride, ste_scale);
  DQRM_LAUNCH_CHECK("interact_bwd_kernel");
  return 0;
}
