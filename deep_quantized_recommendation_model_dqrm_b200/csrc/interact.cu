// (a14) fused dot interaction: cat + T.T^t + lower-triangle gather + cat, forward and backward.
// Reference: DLRM_Net.interact_features, dlrm_s_pytorch_comm_grad.py:701-725 (bmm :710, li/lj index
// lists rebuilt on the host each call :719-722, gather :723, cat :725).
//
// One warp per sample.  The (F+1) x D feature tile is staged once in shared memory (row stride D+4 so
// consecutive rows start on different banks); each lane then owns pairs p = lane, lane+32, ... and
// writes its dot products straight into R next to the copied dense features: no [B,27,27] matrix,
// no index tensors, one launch.  27x27x16 per sample is far below a tcgen05 tile (and fp32-exact
// parity rules out TF32), so this is FFMA work bounded by launch latency, not by a roofline.
#include "common.cuh"

namespace dqrm {

__device__ __forceinline__ int pair_index(int i, int j, int off) { return i * (i - 1) / 2 + off * i + j; }

__device__ __forceinline__ void load_tile(float* tile, int stride, const float* __restrict__ x,
                                          const float* __restrict__ ly, long long lts, long long lbs,
                                          long long b, int F, int dim4, int lane) {
  const int total = (F + 1) * dim4;
  for (int it = lane; it < total; it += 32) {
    const int i = it / dim4, c = it - i * dim4;
    const float* src = (i == 0) ? x + b * dim4 * 4 : ly + (long long)(i - 1) * lts + b * lbs;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c);
    *reinterpret_cast<float4*>(tile + i * stride + c * 4) = v;
  }
}

__global__ void __launch_bounds__(256)
interact_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ly, long long lts, long long lbs,
                    long long batch, int F, int dim4, int off, int npairs, float* __restrict__ R) {
  extern __shared__ __align__(16) float smem_f[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dim = dim4 * 4, stride = dim + 4;
  const int nf = F + 1;
  unsigned short* pairs = reinterpret_cast<unsigned short*>(smem_f);         // [npairs][2]
  float* tile = smem_f + ((npairs * 2 * (int)sizeof(unsigned short) + 15) / 16) * 4 + warp * nf * stride;
  for (int i = threadIdx.x; i < nf; i += blockDim.x)
    for (int j = 0; j < i + off; ++j) {
      const int p = pair_index(i, j, off);
      pairs[2 * p] = (unsigned short)i;
      pairs[2 * p + 1] = (unsigned short)j;
    }
  __syncthreads();
  const long long out_w = dim + npairs;
  for (long long b = (long long)blockIdx.x * warps + warp; b < batch; b += (long long)gridDim.x * warps) {
    __syncwarp();
    load_tile(tile, stride, x, ly, lts, lbs, b, F, dim4, lane);
    __syncwarp();
    float* r = R + b * out_w;
    for (int d = lane; d < dim; d += 32) r[d] = tile[d];
    for (int p = lane; p < npairs; p += 32) {
      const float4* a = reinterpret_cast<const float4*>(tile + pairs[2 * p] * stride);
      const float4* c = reinterpret_cast<const float4*>(tile + pairs[2 * p + 1] * stride);
      float acc = 0.f;
      for (int k = 0; k < dim4; ++k) {
        const float4 u = a[k], v = c[k];
        acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc); acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
      }
      r[dim + p] = acc;
    }
  }
}

__global__ void __launch_bounds__(256)
interact_bwd_kernel(const float* __restrict__ x, const float* __restrict__ ly, long long lts, long long lbs,
                    const float* __restrict__ dR, long long batch, int F, int dim4, int off, int npairs,
                    float* __restrict__ dx, float* __restrict__ dly, long long dts, long long dbs) {
  extern __shared__ __align__(16) float smem_f[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dim = dim4 * 4, stride = dim + 4;
  const int nf = F + 1;
  const int per_warp = nf * stride + ((npairs + 3) & ~3);
  float* tile = smem_f + warp * per_warp;
  float* dz = tile + nf * stride;
  const long long in_w = dim + npairs;
  for (long long b = (long long)blockIdx.x * warps + warp; b < batch; b += (long long)gridDim.x * warps) {
    __syncwarp();
    load_tile(tile, stride, x, ly, lts, lbs, b, F, dim4, lane);
    const float* g = dR + b * in_w;
    for (int p = lane; p < npairs; p += 32) dz[p] = __ldg(g + dim + p);
    __syncwarp();
    for (int it = lane; it < nf * dim4; it += 32) {
      const int i = it / dim4, c = it - i * dim4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < nf; ++j) {
        float w;
        if (j < i) w = dz[pair_index(i, j, off)];
        else if (j > i) w = dz[pair_index(j, i, off)];
        else if (off) w = 2.0f * dz[pair_index(i, i, off)];
        else continue;
        const float4 v = *reinterpret_cast<const float4*>(tile + j * stride + c * 4);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      if (i == 0) {
        acc.x += __ldg(g + c * 4); acc.y += __ldg(g + c * 4 + 1); acc.z += __ldg(g + c * 4 + 2); acc.w += __ldg(g + c * 4 + 3);
        reinterpret_cast<float4*>(dx + b * dim)[c] = acc;
      } else {
        reinterpret_cast<float4*>(dly + (long long)(i - 1) * dts + b * dbs)[c] = acc;
      }
    }
  }
}

static int interact_cfg(int F, int dim, int itself, int* npairs, int* warps, size_t* smem_fwd, size_t* smem_bwd) {
  const int nf = F + 1, off = itself ? 1 : 0;
  *npairs = nf * (nf - 1) / 2 + off * nf;
  const size_t tile = (size_t)nf * (dim + 4) * sizeof(float);
  int w = 8;
  while (w > 1 && w * (tile + ((*npairs + 3) & ~3) * sizeof(float)) > 96 * 1024) w >>= 1;
  *warps = w;
  *smem_fwd = (((size_t)*npairs * 4 + 15) / 16) * 16 + w * tile;
  *smem_bwd = w * (tile + ((*npairs + 3) & ~3) * sizeof(float));
  return (*smem_fwd <= 227 * 1024 && *smem_bwd <= 227 * 1024) ? 0 : -E2BIG;
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
  int npairs, warps;
  size_t sf, sb;
  DQRM_REQUIRE(interact_cfg(num_tables, dim, itself, &npairs, &warps, &sf, &sb) == 0, -E2BIG,
               "interact_fwd: feature tile does not fit shared memory");
  if (sf > 48 * 1024) cudaFuncSetAttribute(interact_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sf);
  long long grid = ceil_div(batch, warps);
  if (grid > 8 * kSMs) grid = 8 * kSMs;
  interact_fwd_kernel<<<(unsigned)grid, warps * 32, sf, static_cast<cudaStream_t>(stream)>>>(
      x, ly, ly_table_stride, ly_bag_stride, batch, num_tables, dim / 4, itself ? 1 : 0, npairs, R);
  DQRM_LAUNCH_CHECK("interact_fwd_kernel");
  return 0;
}

extern "C" int dqrm_interact_bwd(const float* x, const float* ly, int64_t ly_table_stride, int64_t ly_bag_stride,
                                 const float* dR, int64_t batch, int num_tables, int dim, int itself,
                                 float* dx, float* dly, int64_t dly_table_stride, int64_t dly_bag_stride,
                                 void* stream) {
  DQRM_REQUIRE(x && ly && dR && dx && dly, -EINVAL, "interact_bwd: null argument");
  DQRM_REQUIRE(num_tables >= 1 && num_tables < 256, -EINVAL, "interact_bwd: num_tables=%d", num_tables);
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "interact_bwd: dim=%d", dim);
  DQRM_REQUIRE(ly_table_stride % 4 == 0 && ly_bag_stride % 4 == 0 && dly_table_stride % 4 == 0 && dly_bag_stride % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(ly) | reinterpret_cast<uintptr_t>(dx) |
                     reinterpret_cast<uintptr_t>(dly)) & 15u) == 0,
               -EINVAL, "interact_bwd: buffers must be 16-byte aligned with strides multiple of 4");
  if (batch <= 0) return 0;
  int npairs, warps;
  size_t sf, sb;
  DQRM_REQUIRE(interact_cfg(num_tables, dim, itself, &npairs, &warps, &sf, &sb) == 0, -E2BIG,
               "interact_bwd: feature tile does not fit shared memory");
  if (sb > 48 * 1024) cudaFuncSetAttribute(interact_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
  long long grid = ceil_div(batch, warps);
  if (grid > 8 * kSMs) grid = 8 * kSMs;
  interact_bwd_kernel<<<(unsigned)grid, warps * 32, sb, static_cast<cudaStream_t>(stream)>>>(
      x, ly, ly_table_stride, ly_bag_stride, dR, batch, num_tables, dim / 4, itself ? 1 : 0, npairs, dx, dly,
      dly_table_stride, dly_bag_stride);
  DQRM_LAUNCH_CHECK("interact_bwd_kernel");
  return 0;
}
