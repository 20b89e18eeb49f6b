// Bit-packed INT4 tables (two codes per byte) and the gather / dequantise / sum-pool kernel that reads them
// (north_star kernel 2; SURVEY.md section 8 f-4 "packed INT4 checkpoint / inference export").
//
//   pack : code = clamp(rint((1/s_k) * w), -8, 7) per element with the table's scale s_k (the same arithmetic
//          as SymmetricQuantFunction, quant_utils.py:322-346); element d of a row lives in byte d/2, low nibble
//          for even d.  fp32 [N, D] (4 B/elem) -> [N, D/2] bytes: 2.16 GB -> 0.27 GB at Kaggle shape, the model
//          size the paper reports (Table 3).
//   fwd  : out[b] = s_k * sum_{l in bag b} code(row_l)   -- "dequantised sum pooling".  The integer sum is exact
//          and order-free, so long bags are reduced across the warp with shuffles.  For bags of length 1 (every
//          Criteo lookup) this equals the QAT forward q*s bit for bit, because Q(pooled) == Q(row); for longer
//          bags it is the serving semantics (quantise rows, then pool), NOT the training semantics (pool, then
//          quantise, quant_modules_not_quantize_grad.py:367,378) -- which is why training keeps fp32 rows.
//
// Layout per lane: 16 codes = 8 bytes = one 64-bit load; a row of D codes is D/16 consecutive lanes.
#include "common.cuh"

namespace dqrm {

struct PackedSet {
  const unsigned char* p[DQRM_MAX_TABLES];   // packed tables, [rows_k, dim/2] bytes
  long long rows[DQRM_MAX_TABLES];
  long long idx_begin[DQRM_MAX_TABLES + 1];
  int num_tables;
};

struct PackArgs {
  const float* w[DQRM_MAX_TABLES];
  unsigned char* p[DQRM_MAX_TABLES];
  long long chunks[DQRM_MAX_TABLES];         // rows_k * dim / 16
  long long chunk_begin[DQRM_MAX_TABLES + 1];
  int num_tables;
};

// one thread per 16 consecutive elements: 4 x 128-bit loads, one 64-bit store
__global__ void __launch_bounds__(256)
table_pack_int4_kernel(const __grid_constant__ PackArgs a, const float* __restrict__ inv_scale) {
  const long long total = a.chunk_begin[a.num_tables];
  int t = 0;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
    while (c >= a.chunk_begin[t + 1]) ++t;               // chunks are visited in table order per thread
    const long long lc = c - a.chunk_begin[t];
    const float inv = inv_scale[t];
    const float4* src = reinterpret_cast<const float4*>(a.w[t]) + lc * 4;
    unsigned lo = 0u, hi = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = ld_stream_f4(src + j);
      const unsigned q0 = (unsigned)((int)quant_code(v.x, inv, -8.f, 7.f)) & 15u;
      const unsigned q1 = (unsigned)((int)quant_code(v.y, inv, -8.f, 7.f)) & 15u;
      const unsigned q2 = (unsigned)((int)quant_code(v.z, inv, -8.f, 7.f)) & 15u;
      const unsigned q3 = (unsigned)((int)quant_code(v.w, inv, -8.f, 7.f)) & 15u;
      const unsigned h = q0 | (q1 << 4) | (q2 << 8) | (q3 << 12);
      if (j < 2) lo |= h << (16 * j); else hi |= h << (16 * (j - 2));
    }
    reinterpret_cast<uint2*>(a.p[t])[lc] = make_uint2(lo, hi);
  }
}

__device__ __forceinline__ void add_codes16(int (&acc)[16], uint2 v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += ((int)(v.x << (28 - 4 * i))) >> 28;        // sign-extend nibble i
    acc[8 + i] += ((int)(v.y << (28 - 4 * i))) >> 28;
  }
}

// WARP_BAG = false: a group of R = dim/16 lanes owns a bag and walks its lookups (short bags, Criteo).
// WARP_BAG = true : a whole warp owns a bag: 32/R lookups in flight per step, exact integer shuffle reduction.
template <bool WARP_BAG>
__global__ void __launch_bounds__(256)
embbag_fwd_int4_kernel(const __grid_constant__ PackedSet ts, int R, const long long* __restrict__ indices,
                       const long long* __restrict__ offsets, long long bags, const float* __restrict__ scale,
                       float* __restrict__ out, long long out_ts, long long out_bs, int* __restrict__ status) {
  const int group = WARP_BAG ? 32 : R;
  const int lane = threadIdx.x % group;
  const int col = lane % R, slot = lane / R, slots = group / R;
  const long long groups_per_block = blockDim.x / group;
  const long long total = (long long)ts.num_tables * bags;
  int bad = 0;
  for (long long gb = blockIdx.x * groups_per_block + threadIdx.x / group; gb < total;
       gb += (long long)gridDim.x * groups_per_block) {
    const int t = (int)(gb / bags);
    const long long b = gb - (long long)t * bags;
    const long long L = ts.idx_begin[t + 1] - ts.idx_begin[t];
    const long long* idx = indices + ts.idx_begin[t];
    const long long* off = offsets + (long long)t * bags;
    long long start = off[b], end = (b + 1 < bags) ? off[b + 1] : L;
    if (start < 0 || end > L || start > end) {
      bad |= DQRM_STATUS_OFFSET_ORDER;
      start = start < 0 ? 0 : (start > L ? L : start);
      end = end > L ? L : (end < start ? start : end);
    }
    const long long nrows = ts.rows[t];
    const uint2* P = reinterpret_cast<const uint2*>(ts.p[t]);
    int acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    for (long long l = start + slot; l < end; l += 4ll * slots) {
      uint2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                       // 4 independent gathers in flight per lane
        const long long lu = l + (long long)u * slots;
        long long r = lu < end ? idx[lu] : -1;
        if (lu < end && (r < 0 || r >= nrows)) { bad |= DQRM_STATUS_INDEX_RANGE; r = r < 0 ? 0 : nrows - 1; }
        v[u] = r >= 0 ? __ldg(P + r * R + col) : make_uint2(0x0u, 0x0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) add_codes16(acc, v[u]);
    }
    if (WARP_BAG) {
      for (int d = R; d < 32; d <<= 1)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d);
    }
    if (slot == 0) {
      const float s = scale[t];
      float* dst = out + (long long)t * out_ts + b * out_bs + col * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<float4*>(dst)[i] = make_float4(__fmul_rn((float)acc[4 * i], s), __fmul_rn((float)acc[4 * i + 1], s),
                                                        __fmul_rn((float)acc[4 * i + 2], s), __fmul_rn((float)acc[4 * i + 3], s));
    }
  }
  if (bad) atomicOr(status, bad);
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_table_pack_int4(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                    const float* inv_scale, uint8_t* const* packed, void* stream) {
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "table_pack_int4: num_tables=%d", num_tables);
  DQRM_REQUIRE(weight && rows && inv_scale && packed, -EINVAL, "table_pack_int4: null argument");
  DQRM_REQUIRE(dim >= 16 && dim % 16 == 0 && dim <= 512, -EINVAL, "table_pack_int4: dim=%d must be a multiple of 16", dim);
  PackArgs a;
  a.num_tables = num_tables;
  long long tot = 0;
  for (int k = 0; k < num_tables; ++k) {
    DQRM_REQUIRE(weight[k] && packed[k] && rows[k] >= 0, -EINVAL, "table_pack_int4: table %d malformed", k);
    DQRM_REQUIRE(((reinterpret_cast<uintptr_t>(weight[k]) & 15u) | (reinterpret_cast<uintptr_t>(packed[k]) & 7u)) == 0,
                 -EINVAL, "table_pack_int4: table %d alignment (fp32 16 B, packed 8 B)", k);
    a.w[k] = weight[k]; a.p[k] = packed[k];
    a.chunks[k] = rows[k] * dim / 16;
    a.chunk_begin[k] = tot;
    tot += a.chunks[k];
  }
  a.chunk_begin[num_tables] = tot;
  if (tot == 0) return 0;
  long long grid = ceil_div(tot, 256);
  if (grid > 8ll * kSMs) grid = 8ll * kSMs;
  table_pack_int4_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, inv_scale);
  DQRM_LAUNCH_CHECK("table_pack_int4_kernel");
  return 0;
}

extern "C" int dqrm_embbag_fwd_int4(int num_tables, const uint8_t* const* packed, const int64_t* rows, int dim,
                                    const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin,
                                    int64_t bags, const float* scale, float* out, int64_t out_table_stride,
                                    int64_t out_bag_stride, int32_t* status, void* stream) {
  DQRM_REQUIRE(packed && rows && indices && offsets && idx_begin && scale && out && status, -EINVAL, "embbag_fwd_int4: null argument");
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "embbag_fwd_int4: num_tables=%d", num_tables);
  DQRM_REQUIRE(dim >= 16 && dim % 16 == 0 && dim <= 512, -EINVAL, "embbag_fwd_int4: dim=%d must be a multiple of 16", dim);
  const int R = dim / 16;
  DQRM_REQUIRE((R & (R - 1)) == 0, -EINVAL, "embbag_fwd_int4: dim/16 must be a power of two");
  DQRM_REQUIRE(out_bag_stride % 4 == 0 && out_table_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
               -EINVAL, "embbag_fwd_int4: out must be 16-byte aligned with strides multiple of 4");
  PackedSet ts;
  ts.num_tables = num_tables;
  long long lookups = 0;
  for (int k = 0; k < num_tables; ++k) {
    DQRM_REQUIRE(packed[k] && (reinterpret_cast<uintptr_t>(packed[k]) & 7u) == 0, -EINVAL, "embbag_fwd_int4: table %d", k);
    DQRM_REQUIRE(idx_begin[k + 1] >= idx_begin[k], -EINVAL, "embbag_fwd_int4: idx_begin not monotone");
    ts.p[k] = packed[k]; ts.rows[k] = rows[k]; ts.idx_begin[k] = idx_begin[k];
    lookups += idx_begin[k + 1] - idx_begin[k];
  }
  ts.idx_begin[num_tables] = idx_begin[num_tables];
  if (bags <= 0) return 0;
  const long long total = (long long)num_tables * bags;
  const bool warp_bag = lookups >= 8 * total;            // average bag length >= 8: reduce across the warp
  const int group = warp_bag ? 32 : R;
  long long grid = ceil_div(total, 256 / group);
  if (grid > 16ll * kSMs) grid = 16ll * kSMs;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* ip = reinterpret_cast<const long long*>(indices);
  const long long* op = reinterpret_cast<const long long*>(offsets);
  if (warp_bag) embbag_fwd_int4_kernel<true><<<(unsigned)grid, 256, 0, st>>>(ts, R, ip, op, bags, scale, out, out_table_stride, out_bag_stride, status);
  else embbag_fwd_int4_kernel<false><<<(unsigned)grid, 256, 0, st>>>(ts, R, ip, op, bags, scale, out, out_table_stride, out_bag_stride, status);
  DQRM_LAUNCH_CHECK("embbag_fwd_int4_kernel");
  return 0;
}
