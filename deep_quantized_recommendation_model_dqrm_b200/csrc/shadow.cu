// Packed-INT4 shadow rows in the TRAINING forward (north_star kernel 2).
// Reference semantics: QuantEmbeddingBagTwo.forward, quant_modules_not_quantize_grad.py:367,378,393 -- pool in fp32,
// quantise the POOLED vector, dequantise.  For a bag of length 1 (every Criteo lookup, dlrm_data_pytorch.py:342-343)
// Q(pooled) == Q(row), so the forward can read the row's 4-bit codes instead of its fp32 values -- D/2 bytes instead
// of 4D per lookup -- and still produce the same bits, PROVIDED the codes were computed with the scale this forward
// uses.  The fp32 master rows stay authoritative (updates, multi-hot bags, scale changes); the shadow is a cache:
//
//   shadow[t]        [rows_t, D/2] bytes, element d of a row in byte d/2, low nibble for even d (as int4.cu)
//   shadow_scale[t]  the table scale (fp32 bits) the shadow of table t was encoded with
//   refresh          after every scale computation: tables whose scale changed BIT-WISE are re-encoded from the
//                    fp32 rows (one extra pass over that table; for the large tables the max-abs element is almost
//                    never among the ~batch rows an update touches, the 3-row tables change every step and cost
//                    nothing), the others keep their shadow
//   update_rows      after the row update of a step (merge_apply / sgd_rows): the touched rows are re-encoded with the
//                    scale of that step, so the shadow stays valid as long as the next scan returns the same scale
//   forward          per (table, bag): bag length 1 and shadow_scale[t] == scale[t] -> codes from the shadow row, out =
//                    code * s; anything else -> the fp32 path of embbag_fwd.cu (same arithmetic, same bits)
//
// Small tables (packed size <= kStageBytes) are first copied into shared memory with 128-bit loads and served from
// there: at batch 128-8192 the 3..1460-row Criteo tables receive all their lookups from a few KB.
#include "common.cuh"

namespace dqrm {

struct ShadowSet {
  const float* w[DQRM_MAX_TABLES];
  unsigned char* p[DQRM_MAX_TABLES];
  long long rows[DQRM_MAX_TABLES];
  long long idx_begin[DQRM_MAX_TABLES + 1];
  int num_tables;
};

__device__ __forceinline__ unsigned pack4(float4 v, float inv) {          // 4 codes -> 16 bits
  const unsigned q0 = (unsigned)((int)quant_code(v.x, inv, -8.f, 7.f)) & 15u;
  const unsigned q1 = (unsigned)((int)quant_code(v.y, inv, -8.f, 7.f)) & 15u;
  const unsigned q2 = (unsigned)((int)quant_code(v.z, inv, -8.f, 7.f)) & 15u;
  const unsigned q3 = (unsigned)((int)quant_code(v.w, inv, -8.f, 7.f)) & 15u;
  return q0 | (q1 << 4) | (q2 << 8) | (q3 << 12);
}

// flags[t] = shadow of table t is stale (its scale changed bit-wise); shadow_scale <- scale
__global__ void shadow_flags_kernel(int T, const float* __restrict__ scale, float* __restrict__ shadow_scale,
                                    int* __restrict__ flags) {
  const int t = threadIdx.x;
  if (t >= T) return;
  const bool stale = __float_as_uint(scale[t]) != __float_as_uint(shadow_scale[t]);
  flags[t] = stale;
  if (stale) shadow_scale[t] = scale[t];
}

// re-encode every stale table: one thread per 16 consecutive elements (4 x 128-bit streaming loads, one 64-bit store)
__global__ void __launch_bounds__(256)
shadow_refresh_kernel(const __grid_constant__ ShadowSet ts, int dim, const float* __restrict__ inv_scale,
                      const int* __restrict__ flags) {
  for (int t = 0; t < ts.num_tables; ++t) {
    if (!flags[t]) continue;                                               // grid-uniform
    const float inv = inv_scale[t];
    const long long chunks = ts.rows[t] * dim / 16;
    const float4* W = reinterpret_cast<const float4*>(ts.w[t]);
    uint2* P = reinterpret_cast<uint2*>(ts.p[t]);
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += (long long)gridDim.x * blockDim.x) {
      const unsigned h0 = pack4(ld_stream_f4(W + c * 4 + 0), inv), h1 = pack4(ld_stream_f4(W + c * 4 + 1), inv);
      const unsigned h2 = pack4(ld_stream_f4(W + c * 4 + 2), inv), h3 = pack4(ld_stream_f4(W + c * 4 + 3), inv);
      P[c] = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
    }
  }
}

// re-encode the rows a step updated (row lists: the gathered exchange slots of all ranks, or the local unique rows)
__global__ void __launch_bounds__(256)
shadow_update_rows_kernel(const __grid_constant__ ShadowSet ts, int dim, const float* __restrict__ inv_scale,
                          const unsigned char* __restrict__ gathered, size_t slot_bytes, size_t rows_off,
                          const int* __restrict__ uniq_rows, const int* __restrict__ uniq_count, long long capacity) {
  const int t = blockIdx.z, r = blockIdx.y;
  const int* cnt;
  const int* rows;
  if (gathered) {
    const unsigned char* slot = gathered + (size_t)r * slot_bytes;
    cnt = reinterpret_cast<const int*>(slot);
    rows = reinterpret_cast<const int*>(slot + rows_off) + (long long)t * capacity;
  } else {
    cnt = uniq_count;
    rows = uniq_rows + (long long)t * capacity;
  }
  const int U = cnt[t];
  const int per_row = dim / 4;                                             // one thread per float4 -> 2 bytes
  const float inv = inv_scale[t];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)U * per_row;
       e += (long long)gridDim.x * blockDim.x) {
    const long long row = rows[e / per_row];
    const int col = (int)(e % per_row);
    if (row < 0 || row >= ts.rows[t]) continue;
    const float4 v = *(reinterpret_cast<const float4*>(ts.w[t]) + row * per_row + col);
    reinterpret_cast<unsigned short*>(ts.p[t])[row * per_row + col] = (unsigned short)pack4(v, inv);
  }
}

constexpr int kShThreads = 256;
constexpr int kStageBytes = 32 * 1024;                                      // packed tables up to this size are staged

__device__ __forceinline__ int nib(unsigned h, int i) { return ((int)(h << (28 - 4 * i))) >> 28; }   // sign-extended

// One CTA column per table (blockIdx.y = table): lanes of dim/4 own a bag, like embbag_fwd_kernel.
template <typename CodeT>
__global__ void __launch_bounds__(kShThreads)
embbag_fwd_shadow_kernel(const __grid_constant__ ShadowSet ts, int dim4, int group,
                         const long long* __restrict__ indices, const long long* __restrict__ offsets, long long bags,
                         const float* __restrict__ scale, const float* __restrict__ inv_scale,
                         const float* __restrict__ shadow_scale, float* __restrict__ out, long long out_ts,
                         long long out_bs, CodeT* __restrict__ codes, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char s_tab[];
  const int t = blockIdx.y;
  const int lane = threadIdx.x % group;
  const long long L = ts.idx_begin[t + 1] - ts.idx_begin[t];
  const long long* idx = indices + ts.idx_begin[t];
  const long long* off = offsets + (long long)t * bags;
  const long long nrows = ts.rows[t];
  const float s = scale[t], inv = inv_scale[t];
  const bool fresh = __float_as_uint(s) == __float_as_uint(shadow_scale[t]);
  const long long packed_bytes = nrows * dim4 * 2;
  const bool staged = fresh && packed_bytes <= kStageBytes;                 // CTA-uniform
  if (staged) {                                                             // whole packed table -> shared memory
    const uint4* src = reinterpret_cast<const uint4*>(ts.p[t]);
    for (long long i = threadIdx.x; i < (packed_bytes + 15) / 16; i += kShThreads)
      reinterpret_cast<uint4*>(s_tab)[i] = __ldg(src + i);                  // (the arena pads every table to 16 bytes)
    __syncthreads();
  }
  const unsigned short* P = staged ? reinterpret_cast<const unsigned short*>(s_tab)
                                   : reinterpret_cast<const unsigned short*>(ts.p[t]);
  const float4* W = reinterpret_cast<const float4*>(ts.w[t]);
  const long long gpb = kShThreads / group;
  int bad = 0;
  for (long long b = blockIdx.x * gpb + threadIdx.x / group; b < bags; b += (long long)gridDim.x * gpb) {
    long long start = off[b], end = (b + 1 < bags) ? off[b + 1] : L;
    if (start < 0 || end > L || start > end) {
      bad |= DQRM_STATUS_OFFSET_ORDER;
      start = start < 0 ? 0 : (start > L ? L : start);
      end = end > L ? L : (end < start ? start : end);
    }
    for (int col = lane; col < dim4; col += group) {
      float q0, q1, q2, q3;
      if (fresh && end - start == 1) {                                      // the shadow fast path
        long long r = idx[start];
        if (r < 0 || r >= nrows) { bad |= DQRM_STATUS_INDEX_RANGE; r = r < 0 ? 0 : nrows - 1; }
        const unsigned h = P[r * dim4 + col];
        q0 = (float)nib(h, 0); q1 = (float)nib(h, 1); q2 = (float)nib(h, 2); q3 = (float)nib(h, 3);
      } else {                                                              // fp32 rows: pool in index order, then quantise
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long l = start; l < end; ++l) {
          long long r = idx[l];
          if (r < 0 || r >= nrows) { bad |= DQRM_STATUS_INDEX_RANGE; r = r < 0 ? 0 : nrows - 1; }
          const float4 v = __ldg(W + r * dim4 + col);
          if (l == start) acc = v;
          else { acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w); }
        }
        q0 = quant_code(acc.x, inv, -8.f, 7.f); q1 = quant_code(acc.y, inv, -8.f, 7.f);
        q2 = quant_code(acc.z, inv, -8.f, 7.f); q3 = quant_code(acc.w, inv, -8.f, 7.f);
      }
      if (codes) {
        CodeT* cp = codes + (((long long)t * bags + b) * dim4 + col) * 4;
        cp[0] = (CodeT)q0; cp[1] = (CodeT)q1; cp[2] = (CodeT)q2; cp[3] = (CodeT)q3;
      }
      *reinterpret_cast<float4*>(out + (long long)t * out_ts + b * out_bs + col * 4) =
          make_float4(__fmul_rn(q0, s), __fmul_rn(q1, s), __fmul_rn(q2, s), __fmul_rn(q3, s));
    }
  }
  if (bad) atomicOr(status, bad);
}

static int fill_shadow(ShadowSet& ts, int num_tables, const float* const* weight, uint8_t* const* shadow,
                       const int64_t* rows, const int64_t* idx_begin, const char* who) {
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "%s: num_tables=%d", who, num_tables);
  DQRM_REQUIRE(weight && shadow && rows, -EINVAL, "%s: null argument", who);
  ts.num_tables = num_tables;
  for (int k = 0; k < num_tables; ++k) {
    DQRM_REQUIRE(weight[k] && shadow[k] && rows[k] >= 1 && rows[k] < (1ll << 31), -EINVAL, "%s: table %d malformed", who, k);
    DQRM_REQUIRE(((reinterpret_cast<uintptr_t>(weight[k]) | reinterpret_cast<uintptr_t>(shadow[k])) & 15u) == 0, -EINVAL,
                 "%s: table %d: fp32 rows and shadow must be 16-byte aligned", who, k);
    ts.w[k] = weight[k]; ts.p[k] = shadow[k]; ts.rows[k] = rows[k];
    ts.idx_begin[k] = idx_begin ? idx_begin[k] : 0;
  }
  ts.idx_begin[num_tables] = idx_begin ? idx_begin[num_tables] : 0;
  return 0;
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_shadow_refresh(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                   const float* scale, const float* inv_scale, uint8_t* const* shadow,
                                   float* shadow_scale, int32_t* flags, void* stream) {
  DQRM_REQUIRE(scale && inv_scale && shadow_scale && flags, -EINVAL, "shadow_refresh: null argument");
  DQRM_REQUIRE(dim >= 16 && dim % 16 == 0 && dim <= 512, -EINVAL, "shadow_refresh: dim=%d must be a multiple of 16", dim);
  ShadowSet ts;
  if (int rc = fill_shadow(ts, num_tables, weight, shadow, rows, nullptr, "shadow_refresh")) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  shadow_flags_kernel<<<1, DQRM_MAX_TABLES, 0, st>>>(num_tables, scale, shadow_scale, flags);
  shadow_refresh_kernel<<<4 * kSMs, 256, 0, st>>>(ts, dim, inv_scale, flags);
  DQRM_LAUNCH_CHECK("shadow_refresh_kernel");
  return 0;
}

extern "C" int dqrm_shadow_update_rows(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                       const float* inv_scale, uint8_t* const* shadow, const void* gathered, int world,
                                       int64_t capacity, int bits, const int32_t* uniq_rows, const int32_t* uniq_count,
                                       void* stream) {
  DQRM_REQUIRE(inv_scale, -EINVAL, "shadow_update_rows: null argument");
  DQRM_REQUIRE(dim >= 16 && dim % 16 == 0 && dim <= 512, -EINVAL, "shadow_update_rows: dim=%d must be a multiple of 16", dim);
  DQRM_REQUIRE((gathered != nullptr) != (uniq_rows != nullptr), -EINVAL,
               "shadow_update_rows: pass either the gathered slots or a local row list");
  DQRM_REQUIRE(capacity >= 1, -EINVAL, "shadow_update_rows: capacity=%lld", (long long)capacity);
  ShadowSet ts;
  if (int rc = fill_shadow(ts, num_tables, weight, shadow, rows, nullptr, "shadow_update_rows")) return rc;
  size_t slot_bytes = 0, rows_off = 0, codes_off = 0;
  int ranks = 1;
  if (gathered) {
    DQRM_REQUIRE(world >= 1 && world <= 65535 && ((bits >= 2 && bits <= 16) || bits == 32), -EINVAL, "shadow_update_rows: world/bits");
    slot_bytes = dqrm_slot_bytes(num_tables, capacity, dim, bits);
    dqrm_slot_layout(num_tables, capacity, dim, bits, &rows_off, &codes_off);
    ranks = world;
  } else {
    DQRM_REQUIRE(uniq_count, -EINVAL, "shadow_update_rows: uniq_count missing");
  }
  long long gx = ceil_div(capacity * (dim / 4), 256);
  if (gx > 32) gx = 32;
  dim3 grid((unsigned)gx, ranks, num_tables);
  shadow_update_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ts, dim, inv_scale, static_cast<const unsigned char*>(gathered), slot_bytes, rows_off, uniq_rows, uniq_count, capacity);
  DQRM_LAUNCH_CHECK("shadow_update_rows_kernel");
  return 0;
}

extern "C" int dqrm_embbag_fwd_shadow(int num_tables, const float* const* weight, uint8_t* const* shadow,
                                      const int64_t* rows, int dim, const int64_t* indices, const int64_t* offsets,
                                      const int64_t* idx_begin, int64_t bags, const float* scale, const float* inv_scale,
                                      const float* shadow_scale, float* out, int64_t out_table_stride,
                                      int64_t out_bag_stride, void* codes, int32_t* status, void* stream) {
  DQRM_REQUIRE(indices && offsets && idx_begin && scale && inv_scale && shadow_scale && out && status, -EINVAL,
               "embbag_fwd_shadow: null argument");
  DQRM_REQUIRE(dim >= 16 && dim % 16 == 0 && dim <= 512, -EINVAL, "embbag_fwd_shadow: dim=%d must be a multiple of 16", dim);
  DQRM_REQUIRE(out_bag_stride % 4 == 0 && out_table_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
               -EINVAL, "embbag_fwd_shadow: out must be 16-byte aligned with strides multiple of 4");
  ShadowSet ts;
  if (int rc = fill_shadow(ts, num_tables, weight, shadow, rows, idx_begin, "embbag_fwd_shadow")) return rc;
  for (int k = 0; k < num_tables; ++k)
    DQRM_REQUIRE(idx_begin[k + 1] >= idx_begin[k], -EINVAL, "embbag_fwd_shadow: idx_begin not monotone at %d", k);
  if (bags <= 0) return 0;
  const RowLanes rl = row_lanes(dim);
  long long gx = ceil_div(bags, kShThreads / rl.group);
  const long long cap = ceil_div(8ll * kSMs, num_tables);                   // ~8 CTAs per SM over all tables
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, num_tables);
  auto kern = embbag_fwd_shadow_kernel<int8_t>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes);
    DQRM_REQUIRE(e == cudaSuccess, -EIO, "embbag_fwd_shadow: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  kern<<<grid, kShThreads, kStageBytes, static_cast<cudaStream_t>(stream)>>>(
      ts, dim / 4, rl.group, reinterpret_cast<const long long*>(indices), reinterpret_cast<const long long*>(offsets), bags,
      scale, inv_scale, shadow_scale, out, out_table_stride, out_bag_stride, static_cast<int8_t*>(codes), status);
  DQRM_LAUNCH_CHECK("embbag_fwd_shadow_kernel");
  return 0;
}
