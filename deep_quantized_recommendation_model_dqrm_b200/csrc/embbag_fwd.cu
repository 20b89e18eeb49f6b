// (a3) fused gather + sum-pool + fake-quantise + dequantise, all tables in one launch.
// Reference: QuantEmbeddingBagTwo.forward, quantization_supp/quant_modules_not_quantize_grad.py:367,378,393
//            SymmetricQuantFunction / linear_quantize, quantization_supp/quant_utils.py:75-101,322-346.
//
// Latency/HBM-gather bound.  A group of dim/4 lanes owns one bag (D=16: 4 lanes,
// 8 bags per warp; D=64: 16 lanes; D=128: a full warp); each lane moves one
// 128-bit column of the row per lookup.  Lookups of a bag are folded strictly in
// index order (the order ATen uses, so pooled sums and therefore codes are
// bit-identical), but the row loads of up to kUnroll consecutive lookups are
// issued before the first add so a long bag keeps several gathers in flight.
// The pooled vector never leaves registers: it is rounded to the signed
// `bits`-bit code, optionally stored as int8/int16, and written back dequantised
// -- one launch instead of the reference's ~12 per table.
#include "common.cuh"

namespace dqrm {

constexpr int kFwdThreads = 256;
constexpr int kFwdUnroll = 4;

template <int COLS, typename CodeT>
__global__ void __launch_bounds__(kFwdThreads)
embbag_fwd_kernel(const __grid_constant__ TableSet ts, int dim4, int group,
                  const long long* __restrict__ indices, const long long* __restrict__ offsets, long long bags,
                  const float* __restrict__ scale, const float* __restrict__ inv_scale, int bits,
                  float* __restrict__ out, long long out_ts, long long out_bs,
                  CodeT* __restrict__ codes, int* __restrict__ status) {
  const int lane = threadIdx.x % group;
  const long long groups_per_block = kFwdThreads / group;
  const long long total = (long long)ts.num_tables * bags;
  const float hi = qmax_of(bits), lo = -hi - 1.0f;
  int bad = 0;
  for (long long gb = blockIdx.x * groups_per_block + threadIdx.x / group; gb < total;
       gb += (long long)gridDim.x * groups_per_block) {
    const int t = (int)(gb / bags);
    const long long b = gb - (long long)t * bags;
    const long long L = ts.idx_begin[t + 1] - ts.idx_begin[t];
    const long long* idx = indices + ts.idx_begin[t];
    const long long* off = offsets + (long long)t * bags;
    long long start = off[b];
    long long end = (b + 1 < bags) ? off[b + 1] : L;
    if (start < 0 || end > L || start > end) {   // malformed offsets: clamp and flag
      bad |= DQRM_STATUS_OFFSET_ORDER;
      start = start < 0 ? 0 : (start > L ? L : start);
      end = end > L ? L : (end < start ? start : end);
    }
    const long long nrows = ts.rows[t];
    const float4* W = reinterpret_cast<const float4*>(ts.w[t]);

    float4 acc[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool first = true;
    for (long long l = start; l < end; l += kFwdUnroll) {
      long long r[kFwdUnroll];
      float4 v[kFwdUnroll][COLS];
#pragma unroll
      for (int u = 0; u < kFwdUnroll; ++u) {
        r[u] = (l + u < end) ? idx[l + u] : -1;
        if (l + u < end && (r[u] < 0 || r[u] >= nrows)) {
          bad |= DQRM_STATUS_INDEX_RANGE;
          r[u] = r[u] < 0 ? 0 : nrows - 1;
        }
      }
#pragma unroll
      for (int u = 0; u < kFwdUnroll; ++u)
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          const int col = lane + c * group;
          v[u][c] = (r[u] >= 0 && col < dim4) ? __ldg(W + r[u] * dim4 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < kFwdUnroll; ++u) {
        if (r[u] < 0) break;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          if (first) {
            acc[c] = v[u][c];
          } else {
            acc[c].x = __fadd_rn(acc[c].x, v[u][c].x);
            acc[c].y = __fadd_rn(acc[c].y, v[u][c].y);
            acc[c].z = __fadd_rn(acc[c].z, v[u][c].z);
            acc[c].w = __fadd_rn(acc[c].w, v[u][c].w);
          }
        }
        first = false;
      }
    }

    const bool quant = scale != nullptr;
    const float s = quant ? scale[t] : 1.0f;
    const float inv = quant ? inv_scale[t] : 1.0f;
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int col = lane + c * group;
      if (col >= dim4) continue;
      float4 o = acc[c];
      if (quant) {
        const float q0 = quant_code(o.x, inv, lo, hi), q1 = quant_code(o.y, inv, lo, hi);
        const float q2 = quant_code(o.z, inv, lo, hi), q3 = quant_code(o.w, inv, lo, hi);
        if (codes) {
          CodeT* cp = codes + (gb * dim4 + col) * 4;
          cp[0] = (CodeT)q0; cp[1] = (CodeT)q1; cp[2] = (CodeT)q2; cp[3] = (CodeT)q3;
        }
        o = make_float4(__fmul_rn(q0, s), __fmul_rn(q1, s), __fmul_rn(q2, s), __fmul_rn(q3, s));
      }
      float* dst = out + (long long)t * out_ts + b * out_bs + col * 4;
      *reinterpret_cast<float4*>(dst) = o;
    }
  }
  if (bad) atomicOr(status, bad);
}

template <int COLS, typename CodeT>
static int launch_fwd(const TableSet& ts, int dim, RowLanes rl, const int64_t* indices, const int64_t* offsets,
                      int64_t bags, const float* scale, const float* inv_scale, int bits, float* out,
                      int64_t out_ts, int64_t out_bs, void* codes, int32_t* status, cudaStream_t st) {
  const long long total = (long long)ts.num_tables * bags;
  const long long per_block = kFwdThreads / rl.group;
  long long grid = ceil_div(total, per_block);
  const long long cap = (long long)kSMs * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  embbag_fwd_kernel<COLS, CodeT><<<(unsigned)grid, kFwdThreads, 0, st>>>(
      ts, dim / 4, rl.group, reinterpret_cast<const long long*>(indices),
      reinterpret_cast<const long long*>(offsets), bags, scale, inv_scale, bits, out, out_ts, out_bs,
      static_cast<CodeT*>(codes), status);
  DQRM_LAUNCH_CHECK("embbag_fwd_kernel");
  return 0;
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int dqrm_embbag_fwd(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                               const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin, int64_t bags,
                               const float* scale, const float* inv_scale, int bits,
                               float* out, int64_t out_table_stride, int64_t out_bag_stride,
                               void* codes, int32_t* status, void* stream) {
  DQRM_REQUIRE(weight && rows && indices && offsets && idx_begin && out && status, -EINVAL, "embbag_fwd: null argument");
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "embbag_fwd: dim=%d must be a multiple of 4 in [4,512]", dim);
  DQRM_REQUIRE(bags >= 0, -EINVAL, "embbag_fwd: bags<0");
  DQRM_REQUIRE((scale == nullptr) == (inv_scale == nullptr), -EINVAL, "embbag_fwd: scale/inv_scale must both be set or both NULL");
  DQRM_REQUIRE(!scale || (bits >= 2 && bits <= 16), -EINVAL, "embbag_fwd: bits=%d outside [2,16]", bits);
  DQRM_REQUIRE(!codes || scale, -EINVAL, "embbag_fwd: codes requested on the full-precision path");
  DQRM_REQUIRE(out_bag_stride % 4 == 0 && out_table_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
               -EINVAL, "embbag_fwd: out must be 16-byte aligned with strides multiple of 4");
  TableSet ts;
  if (int rc = fill_tables(ts, num_tables, weight, rows, idx_begin)) return rc;
  if (bags == 0) return 0;
  const RowLanes rl = row_lanes(dim);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool wide = bits > 8;
#define DQRM_FWD(COLS)                                                                                          \
  return wide ? launch_fwd<COLS, int16_t>(ts, dim, rl, indices, offsets, bags, scale, inv_scale, bits, out,      \
                                          out_table_stride, out_bag_stride, codes, status, st)                  \
              : launch_fwd<COLS, int8_t>(ts, dim, rl, indices, offsets, bags, scale, inv_scale, bits, out,       \
                                         out_table_stride, out_bag_stride, codes, status, st)
  if (rl.cols == 1) { DQRM_FWD(1); }
  if (rl.cols == 2) { DQRM_FWD(2); }
  DQRM_FWD(4);
#undef DQRM_FWD
}
