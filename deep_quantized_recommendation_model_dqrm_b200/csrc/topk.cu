// (a8) top-k row sparsification of the de-duplicated embedding gradients (north_star extension).
// The DQRM path has no top-k; score and selection follow the only top-k in the
// reference tree, training_imagenet_speedup.py:138 (score = ||g_row||^2 / cols) and :149 (torch.topk).
//
// One CTA per table: scores -> 64-bit keys (~score_bits << 32 | position) -> bitonic sort in shared
// memory (largest score first, ties by lower row id) -> keep flags for the first k -> block scan ->
// in-place compaction of (rows, sums) in waves (a survivor only ever moves to a lower slot).
#include "common.cuh"

namespace dqrm {

template <int COLS>
__global__ void __launch_bounds__(1024)
grad_topk_kernel(int dim4, int group, float* __restrict__ grad_sums, int* __restrict__ uniq_rows,
                 int* __restrict__ uniq_count, long long capacity, int topk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_warp_tot[32];
  const int t = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int U = uniq_count[t];
  if (topk >= U) return;                                  // identity (block-uniform)
  int n = 2;
  while (n < U) n <<= 1;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  int* newpos = reinterpret_cast<int*>(keys + n);         // keep flag, then destination slot (-1 = dropped)
  float* sums = grad_sums + (long long)t * capacity * dim4 * 4;
  int* rows = uniq_rows + (long long)t * capacity;
  const int dim = dim4 * 4;

  for (int j = tid; j < n; j += nthr) {
    unsigned long long key = ~0ull;
    if (j < U) {
      const float* g = sums + (long long)j * dim;
      float sq = 0.f;
      for (int d = 0; d < dim; ++d) sq = __fadd_rn(sq, __fmul_rn(g[d], g[d]));
      const float score = __fdiv_rn(sq, (float)dim);
      key = ((unsigned long long)(~__float_as_uint(score)) << 32) | (unsigned)j;
    }
    keys[j] = key;
    newpos[j] = 0;
  }
  __syncthreads();
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < (n >> 1); i += nthr) {
        const int lo = 2 * i - (i & (j - 1)), hi = lo + j;
        const unsigned long long x = keys[lo], y = keys[hi];
        if ((x > y) == ((lo & k) == 0)) { keys[lo] = y; keys[hi] = x; }
      }
      __syncthreads();
    }
  for (int i = tid; i < topk; i += nthr) newpos[(unsigned)keys[i]] = 1;
  __syncthreads();

  // exclusive scan of the keep flags over [0, U)
  const int chunk = (U + nthr - 1) / nthr;
  const int c0 = min(tid * chunk, U), c1 = min(c0 + chunk, U);
  int cnt = 0;
  for (int i = c0; i < c1; ++i) cnt += newpos[i];
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += v;
  }
  if ((tid & 31) == 31) s_warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    const int nw = (nthr + 31) >> 5;
    const int v = tid < nw ? s_warp_tot[tid] : 0;
    int w = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, w, d);
      if (tid >= d) w += u;
    }
    s_warp_tot[tid] = w - v;
  }
  __syncthreads();
  int pos = s_warp_tot[tid >> 5] + incl - cnt;
  for (int i = c0; i < c1; ++i) newpos[i] = newpos[i] ? pos++ : -1;
  __syncthreads();

  // in-place compaction, one wave of rows at a time
  const int lane = tid % group, gpb = nthr / group;
  for (int wb = 0; wb < U; wb += gpb) {
    const int j = wb + tid / group;
    const int dst = j < U ? newpos[j] : -1;
    float4 v[COLS];
    int row = 0;
    if (dst >= 0) {
      row = rows[j];
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const int col = lane + c * group;
        if (col < dim4) v[c] = reinterpret_cast<const float4*>(sums + (long long)j * dim)[col];
      }
    }
    __syncthreads();
    if (dst >= 0) {
      if (lane == 0) rows[dst] = row;
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const int col = lane + c * group;
        if (col < dim4) reinterpret_cast<float4*>(sums + (long long)dst * dim)[col] = v[c];
      }
    }
    __syncthreads();
  }
  if (tid == 0) uniq_count[t] = topk;
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_grad_topk(int num_tables, int dim, float* grad_sums, int32_t* uniq_rows, int32_t* uniq_count,
                              int64_t capacity, int64_t topk, void* stream) {
  DQRM_REQUIRE(grad_sums && uniq_rows && uniq_count, -EINVAL, "grad_topk: null argument");
  DQRM_REQUIRE(num_tables >= 1, -EINVAL, "grad_topk: num_tables=%d", num_tables);
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "grad_topk: dim=%d", dim);
  DQRM_REQUIRE(topk >= 1, -EINVAL, "grad_topk: topk=%lld must be >= 1", (long long)topk);
  if (topk >= capacity) return 0;                       // identity
  DQRM_REQUIRE(capacity <= DQRM_BWD_CTA_MAX_LOOKUPS, -E2BIG, "grad_topk: capacity %lld > %d", (long long)capacity,
               DQRM_BWD_CTA_MAX_LOOKUPS);
  int n = 2;
  while (n < capacity) n <<= 1;
  int threads = n / 2 < 128 ? 128 : (n / 2 > 1024 ? 1024 : n / 2);
  const size_t smem = (size_t)n * 12;
  const RowLanes rl = row_lanes(dim);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define DQRM_TOPK(COLS)                                                                                   \
  do {                                                                                                    \
    auto kern = grad_topk_kernel<COLS>;                                                                   \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    kern<<<num_tables, threads, smem, st>>>(dim / 4, rl.group, grad_sums, uniq_rows, uniq_count, capacity, (int)topk); \
  } while (0)
  if (rl.cols == 1) DQRM_TOPK(1); else if (rl.cols == 2) DQRM_TOPK(2); else DQRM_TOPK(4);
#undef DQRM_TOPK
  DQRM_LAUNCH_CHECK("grad_topk_kernel");
  return 0;
}
