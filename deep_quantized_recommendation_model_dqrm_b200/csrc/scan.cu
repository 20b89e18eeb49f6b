// (a1) per-table max-abs scan + scale, all tables in one launch.
// Reference: symmetric_linear_quantization_param_two, quantization_supp/quant_utils.py:141-194.
//
// HBM-bound streaming reduction: N*D*4 bytes read exactly once.  Persistent grid
// (kSMs * kCtasPerSm CTAs), each CTA walks 16 KiB tiles round-robin; each thread
// keeps kVec independent 128-bit streaming loads in flight (128 KiB per SM),
// folds |x| as an integer max on the IEEE bit pattern (monotone for
// non-negative floats), and the CTA publishes one atomicMax per table it
// touched.  The last CTA to finish turns the per-table maxima into
// (absmax, scale, 1/scale) and re-zeros the workspace, so the scale never
// visits the host (the reference pays a D2H sync per table, quant_utils.py:191).
#include <stdlib.h>
#include "common.cuh"

namespace dqrm {

constexpr int kScanThreads = 256;
constexpr int kScanVec = 4;                                   // float4 loads in flight per thread
constexpr int kScanTile = kScanThreads * kScanVec * 4;        // floats per tile (16 KiB)
constexpr int kScanCtasPerSm = 8;                             // launch bound (most that can be resident)
constexpr int kScanCtasDefault = 4;                           // launched: 64 KiB in flight per SM; measured 6.93 TB/s vs 6.82 at 8

struct ScanArgs {
  const float* w[DQRM_MAX_TABLES];
  long long elems[DQRM_MAX_TABLES];
  int tile_begin[DQRM_MAX_TABLES + 1];
  int num_tables;
};

__global__ void __launch_bounds__(kScanThreads, kScanCtasPerSm)
table_absmax_kernel(const __grid_constant__ ScanArgs a, unsigned* __restrict__ acc, unsigned* __restrict__ counter,
                    float* __restrict__ absmax_out, float* __restrict__ scale_out, float* __restrict__ inv_out,
                    int bits, int total_tiles) {
  __shared__ unsigned s_max;
  __shared__ bool s_last;
  unsigned m = 0u;
  int cur = 0;
  auto flush = [&]() {
    unsigned bm = block_max_u32(m, &s_max);
    if (threadIdx.x == 0 && bm) atomicMax(&acc[cur], bm);
    m = 0u;
  };
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    while (tile >= a.tile_begin[cur + 1]) {       // block-uniform: tiles are visited in table order
      flush();
      ++cur;
    }
    const long long base = (long long)(tile - a.tile_begin[cur]) * kScanTile;
    const long long remain = a.elems[cur] - base;
    const float* p = a.w[cur] + base;
    if (remain >= kScanTile) {
      const float4* p4 = reinterpret_cast<const float4*>(p) + threadIdx.x;
      float4 v[kScanVec];
#pragma unroll
      for (int j = 0; j < kScanVec; ++j) v[j] = ld_stream_f4(p4 + j * kScanThreads);
#pragma unroll
      for (int j = 0; j < kScanVec; ++j) m = max(m, abs_bits4(v[j]));
    } else {
      const long long n4 = remain >> 2;
      const float4* p4 = reinterpret_cast<const float4*>(p);
      for (long long i = threadIdx.x; i < n4; i += kScanThreads) m = max(m, abs_bits4(ld_stream_f4(p4 + i)));
      for (long long i = (n4 << 2) + threadIdx.x; i < remain; i += kScanThreads) m = max(m, abs_bits(p[i]));
    }
  }
  flush();

  __threadfence();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    for (int k = threadIdx.x; k < a.num_tables; k += kScanThreads) {
      const float am = __uint_as_float(atomicExch(&acc[k], 0u));
      absmax_out[k] = am;
      if (scale_out) {
        const float s = scale_of(am, bits);
        scale_out[k] = s;
        inv_out[k] = __fdiv_rn(1.0f, s);
      }
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

__global__ void scale_from_absmax_kernel(int n, const float* __restrict__ absmax, int bits,
                                         float* __restrict__ scale, float* __restrict__ inv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = scale_of(absmax[i], bits);
    scale[i] = s;
    if (inv) inv[i] = __fdiv_rn(1.0f, s);
  }
}

}  // namespace dqrm

using namespace dqrm;

extern "C" size_t dqrm_scan_workspace_bytes(int num_tables) {
  return sizeof(unsigned) * (size_t)(num_tables + 1);
}

extern "C" int dqrm_table_absmax_scale(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                       int bits, int shard_rank, int shard_world,
                                       float* absmax, float* scale, float* inv_scale,
                                       void* workspace, void* stream) {
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "table_absmax_scale: num_tables=%d", num_tables);
  DQRM_REQUIRE(weight && rows && absmax && workspace, -EINVAL, "table_absmax_scale: null argument");
  DQRM_REQUIRE(dim >= 1, -EINVAL, "table_absmax_scale: dim=%d", dim);
  DQRM_REQUIRE((scale == nullptr) == (inv_scale == nullptr), -EINVAL, "table_absmax_scale: scale/inv_scale must both be set or both NULL");
  DQRM_REQUIRE(!scale || (bits >= 2 && bits <= 16), -EINVAL, "table_absmax_scale: bits=%d outside [2,16]", bits);
  DQRM_REQUIRE(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world, -EINVAL,
               "table_absmax_scale: shard %d/%d", shard_rank, shard_world);
  ScanArgs a;
  a.num_tables = num_tables;
  long long tiles = 0;
  for (int k = 0; k < num_tables; ++k) {
    DQRM_REQUIRE(rows[k] >= 0, -EINVAL, "table_absmax_scale: rows[%d]<0", k);
    // contiguous balanced row shard (same split as get_my_slice, dlrm_s_pytorch_comm_grad.py:993-997)
    const long long q = rows[k] / shard_world, r = rows[k] % shard_world;
    const long long lo = shard_rank * q + (shard_rank < r ? shard_rank : r);
    const long long hi = lo + q + (shard_rank < r ? 1 : 0);
    a.w[k] = weight[k] + lo * dim;
    a.elems[k] = (hi - lo) * dim;
    DQRM_REQUIRE((reinterpret_cast<uintptr_t>(a.w[k]) & 15u) == 0 || a.elems[k] == 0, -EINVAL,
                 "table_absmax_scale: table %d shard base not 16-byte aligned (dim %% 4 != 0?)", k);
    a.tile_begin[k] = (int)tiles;
    tiles += ceil_div(a.elems[k], kScanTile);
    DQRM_REQUIRE(tiles < (1ll << 31), -E2BIG, "table_absmax_scale: too many tiles");
  }
  a.tile_begin[num_tables] = (int)tiles;
  unsigned* acc = static_cast<unsigned*>(workspace);
  unsigned* counter = acc + num_tables;
  // resident CTAs per SM: 8 fills every thread slot; fewer leave room for kernels of other streams to run beside the
  // pass (the bottom MLP does not depend on it) -- 46 KiB in flight per SM already covers HBM latency x bandwidth
  static const int ctas_per_sm = [] {
    const char* e = getenv("DQRM_SCAN_CTAS_PER_SM");
    const int v = e ? atoi(e) : kScanCtasDefault;
    return v < 1 ? 1 : (v > kScanCtasPerSm ? kScanCtasPerSm : v);
  }();
  long long grid = tiles < (long long)kSMs * ctas_per_sm ? tiles : (long long)kSMs * ctas_per_sm;
  if (grid < 1) grid = 1;
  // (Measured, not adopted: asking for the max shared-memory carve-out so that kernels needing shared memory could
  // become resident beside the pass -- an SM's L1/shared split only changes while the SM is empty -- cuts the pass to
  // 5.5 TB/s, and the co-running latency-bound kernels crawl under the saturated memory system: DESIGN.md 5b.)
  table_absmax_kernel<<<(unsigned)grid, kScanThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      a, acc, counter, absmax, scale, inv_scale, bits, (int)tiles);
  DQRM_LAUNCH_CHECK("table_absmax_kernel");
  return 0;
}

extern "C" int dqrm_scale_from_absmax(int n_scales, const float* absmax, int bits, float* scale, float* inv_scale,
                                      void* stream) {
  DQRM_REQUIRE(n_scales >= 1 && absmax && scale, -EINVAL, "scale_from_absmax: bad argument");
  DQRM_REQUIRE(bits >= 2 && bits <= 16, -EINVAL, "scale_from_absmax: bits=%d outside [2,16]", bits);
  scale_from_absmax_kernel<<<(n_scales + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      n_scales, absmax, bits, scale, inv_scale);
  DQRM_LAUNCH_CHECK("scale_from_absmax_kernel");
  return 0;
}
