// (a7 steps 3-5, a8, a9) quantised, sparsified embedding-gradient exchange.
// Reference: quantize_emb_grad, sgd_quantized_gradients_parallel_comm.py:850-890 (scale mean :865-866,
//            quantise :869, sparse all-reduce :878, 1/N :885) and weight_update_parallel_comm :601-628.
//
// The reference ships, per table, a Gloo sparse all-reduce of (int64 row, fp32 "int8" values): 52
// host-staged collectives per step.  Here every rank packs ALL tables into one fixed-capacity slot
//   int32 count[T] | int32 rows[T][cap] | int8 codes[T][cap][D]
// so the whole exchange is ONE all-gather (5 B/element-row instead of 12 B), and the merge kernel
// reproduces the sparse all-reduce + SGD update on every rank from the same bytes:
// each (rank, row) entry binary-searches the other ranks' sorted row lists; the lowest rank holding a
// row owns it, sums the integer codes of all holders exactly, and applies
//   W[row] += (-lr) * ((float(sum q) * (1/N)) * s_bar)
// once, in the reference's association.  No sort, no atomics on W, replicas stay bit-identical.
#include "common.cuh"

namespace dqrm {

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

struct SlotLayout { size_t rows_off, codes_off, bytes; int code_bytes; };
inline SlotLayout slot_layout(int num_tables, int64_t capacity, int dim, int bits) {
  SlotLayout l;
  l.code_bytes = bits == 32 ? 4 : (bits > 8 ? 2 : 1);   // 32 = un-quantised fp32 payload
  l.rows_off = align16((size_t)num_tables * 4);
  l.codes_off = l.rows_off + align16((size_t)num_tables * capacity * 4);
  l.bytes = l.codes_off + align16((size_t)num_tables * capacity * dim * l.code_bytes);
  return l;
}

template <typename CodeT> struct Code4;
template <> struct Code4<int8_t> { using type = char4; };
template <> struct Code4<int16_t> { using type = short4; };
template <> struct Code4<float> { using type = float4; };        // un-quantised exchange (emb_grad_quantized=False)
template <typename CodeT> struct AccOf { using type = int; };
template <> struct AccOf<float> { using type = float; };

// s_bar = (sum_r s_r) * (1/N) in rank order
__device__ __forceinline__ float mean_scale(const float* __restrict__ gathered, int world, long long stride, int t, float inv_world) {
  float acc = gathered[t];
  for (int r = 1; r < world; ++r) acc = __fadd_rn(acc, gathered[(long long)r * stride + t]);
  return __fmul_rn(acc, inv_world);
}

template <int COLS, typename CodeT>
__global__ void __launch_bounds__(256)
grad_pack_kernel(int T, int dim4, int group, const float* __restrict__ grad_sums, const int* __restrict__ uniq_rows,
                 const int* __restrict__ uniq_count, long long capacity, const float* __restrict__ gathered_scales,
                 long long scale_stride, int world, float inv_world, int bits, unsigned char* __restrict__ slot, SlotLayout lay,
                 float* __restrict__ scale_mean) {
  using C4 = typename Code4<CodeT>::type;
  const int t = blockIdx.y;
  const int U = uniq_count[t];
  const float s_bar = mean_scale(gathered_scales, world, scale_stride, t, inv_world);
  const float inv = __fdiv_rn(1.0f, s_bar);
  const float hi = qmax_of(bits), lo = -hi - 1.0f;
  int* cnt = reinterpret_cast<int*>(slot);
  int* rows = reinterpret_cast<int*>(slot + lay.rows_off);
  C4* codes = reinterpret_cast<C4*>(slot + lay.codes_off);
  if (blockIdx.x == 0 && threadIdx.x == 0) { cnt[t] = U; scale_mean[t] = s_bar; }
  const int lane = threadIdx.x % group, gpb = blockDim.x / group;
  for (int j = blockIdx.x * gpb + threadIdx.x / group; j < U; j += gridDim.x * gpb) {
    const long long e = (long long)t * capacity + j;
    if (lane == 0) rows[e] = uniq_rows[e];
    const float4* g4 = reinterpret_cast<const float4*>(grad_sums + e * dim4 * 4);
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int col = lane + c * group;
      if (col >= dim4) continue;
      const float4 g = g4[col];
      C4 q;
      if constexpr (sizeof(CodeT) == 4) {
        q.x = g.x; q.y = g.y; q.z = g.z; q.w = g.w;                      // fp32 sums travel as they are
      } else {
        q.x = (CodeT)quant_code(g.x, inv, lo, hi); q.y = (CodeT)quant_code(g.y, inv, lo, hi);
        q.z = (CodeT)quant_code(g.z, inv, lo, hi); q.w = (CodeT)quant_code(g.w, inv, lo, hi);
      }
      codes[e * dim4 + col] = q;
    }
  }
}

// position of x in the ascending list a[0..n), or -1
__device__ __forceinline__ int find_row(const int* __restrict__ a, int n, int x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int v = __ldg(a + mid);
    if (v < x) lo = mid + 1; else hi = mid;
  }
  return (lo < n && __ldg(a + lo) == x) ? lo : -1;
}

constexpr int kMergeRanks = 8;          // row lists searched in lock-step (world > 8: in chunks of 8)

// MULTI = false is the one-rank instantiation (world == 1: nothing to merge): without the lock-step searches it needs
// half the registers (64 -> full occupancy), which is what a latency-bound random-row read-modify-write lives on
// (1 M rows x 64 floats: 192 -> 103 us = 5.1 TB/s of row traffic; the lock-step multi-rank searches had doubled the
// registers of the one-rank case too).
template <int COLS, typename CodeT, bool MULTI>
__global__ void __launch_bounds__(256, MULTI ? 4 : 8)
grad_merge_apply_kernel(const __grid_constant__ TableSet ts, int dim4, int group,
                        const unsigned char* __restrict__ gathered, SlotLayout lay, int world, long long capacity,
                        const float* __restrict__ scale_mean, float neg_lr_arg, const float* __restrict__ lr_dev,
                        float inv_world,
                        int* __restrict__ updated_rows, int* __restrict__ updated_count, float* __restrict__ qbar,
                        int* __restrict__ status, int search_iters) {
  using C4 = typename Code4<CodeT>::type;
  if (*status & DQRM_STATUS_P2P_TIMEOUT) return;                         // an exchange timed out: never apply stale slots
  const float neg_lr = lr_dev ? -(*lr_dev) : neg_lr_arg;                 // lr from device memory: graph replays follow a schedule
  const int t = blockIdx.z, r = blockIdx.y;
  const unsigned char* my = gathered + (size_t)r * lay.bytes;
  const int U = reinterpret_cast<const int*>(my)[t];
  const int* my_rows = reinterpret_cast<const int*>(my + lay.rows_off) + (long long)t * capacity;
  const float s_bar = scale_mean[t];
  const long long nrows = ts.rows[t];
  const int lane = threadIdx.x % group, gpb = blockDim.x / group;
  int bad = 0;
  for (int jb = blockIdx.x * gpb; jb < U; jb += gridDim.x * gpb) {      // block-uniform trip count
    const int j = jb + threadIdx.x / group;
    bool live = j < U;
    int x = live ? my_rows[j] : 0;
    if (live && (x < 0 || x >= nrows)) { bad |= DQRM_STATUS_INDEX_RANGE; live = false; }
    // Look x up in the sorted row lists of ALL other ranks at once: the binary searches advance in lock-step, so the
    // dependent chain is log2(capacity) L2 round trips instead of (world-1) x log2(capacity) (22 us at eight ranks).
    // Ownership: the lowest rank that lists x applies the update; it sums the codes of the holders in rank order.
    typename AccOf<CodeT>::type q[COLS][4];
    if (live) {
      const C4* mc = reinterpret_cast<const C4*>(my + lay.codes_off) + ((long long)t * capacity + j) * dim4;
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const int col = lane + c * group;
        const C4 v = col < dim4 ? mc[col] : C4{0, 0, 0, 0};
        q[c][0] = v.x; q[c][1] = v.y; q[c][2] = v.z; q[c][3] = v.w;
      }
    }
    if constexpr (MULTI)
    for (int base = 0; world > 1 && base < world; base += kMergeRanks) {  // (block-uniform trip counts throughout)
      const int* lst[kMergeRanks];
      int n2[kMergeRanks], lo[kMergeRanks], hi[kMergeRanks];
#pragma unroll
      for (int i = 0; i < kMergeRanks; ++i) {
        const int r2 = base + i;
        const bool use = r2 < world && r2 != r;
        const unsigned char* o = gathered + (size_t)(use ? r2 : r) * lay.bytes;
        lst[i] = reinterpret_cast<const int*>(o + lay.rows_off) + (long long)t * capacity;
        n2[i] = 0;
        if (use) n2[i] = min(__ldg(reinterpret_cast<const int*>(o) + t), (int)capacity);
      }
#pragma unroll
      for (int i = 0; i < kMergeRanks; ++i) { lo[i] = 0; hi[i] = n2[i]; }
      for (int it = 0; it < search_iters; ++it) {
        int v[kMergeRanks];
#pragma unroll
        for (int i = 0; i < kMergeRanks; ++i) {                           // (mid <= capacity - 1; closed searches load nothing)
          v[i] = 0;
          if (lo[i] < hi[i]) v[i] = __ldg(lst[i] + ((lo[i] + hi[i]) >> 1));
        }
#pragma unroll
        for (int i = 0; i < kMergeRanks; ++i) {
          const int mid = (lo[i] + hi[i]) >> 1;
          const bool open = lo[i] < hi[i];
          if (open && v[i] < x) lo[i] = mid + 1; else if (open) hi[i] = mid;
        }
      }
      int at[kMergeRanks];
#pragma unroll
      for (int i = 0; i < kMergeRanks; ++i) {
        at[i] = -1;
        if (lo[i] < n2[i]) at[i] = __ldg(lst[i] + lo[i]);
      }
      bool found[kMergeRanks];
#pragma unroll
      for (int i = 0; i < kMergeRanks; ++i) {
        found[i] = live && lo[i] < n2[i] && at[i] == x;
        if (found[i] && base + i < r) live = false;                       // a lower rank owns the row
      }
#pragma unroll
      for (int i = 0; i < kMergeRanks; ++i) {
        if (!(found[i] && live && base + i > r)) continue;
        const unsigned char* o = gathered + (size_t)(base + i) * lay.bytes;
        const C4* oc = reinterpret_cast<const C4*>(o + lay.codes_off) + ((long long)t * capacity + lo[i]) * dim4;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          const int col = lane + c * group;
          if (col >= dim4) continue;
          const C4 v = oc[col];
          q[c][0] += v.x; q[c][1] += v.y; q[c][2] += v.z; q[c][3] += v.w;
        }
      }
    }
    int slot_pos = 0;
    if (updated_rows) {          // warp-uniform branch; every lane takes part in the shuffle
      if (live && lane == 0) {
        slot_pos = atomicAdd(&updated_count[t], 1);
        updated_rows[(long long)t * world * capacity + slot_pos] = x;
      }
      slot_pos = __shfl_sync(0xffffffffu, slot_pos, (threadIdx.x & 31) - lane);
    }
    if (!live) continue;
    float4* w4 = reinterpret_cast<float4*>(ts.w[t] + (long long)x * dim4 * 4);
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int col = lane + c * group;
      if (col >= dim4) continue;
      float qb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qb[i] = __fmul_rn((float)q[c][i], inv_world);       // (sum q) * (1/N)   :885 / :322
      if (qbar) reinterpret_cast<float4*>(qbar + ((long long)t * world * capacity + slot_pos) * dim4 * 4)[col] =
                    make_float4(qb[0], qb[1], qb[2], qb[3]);
      float4 w = w4[col];
      if constexpr (sizeof(CodeT) == 4) {                                             // W += -lr * grad   :626
        w.x = __fadd_rn(w.x, __fmul_rn(neg_lr, qb[0])); w.y = __fadd_rn(w.y, __fmul_rn(neg_lr, qb[1]));
        w.z = __fadd_rn(w.z, __fmul_rn(neg_lr, qb[2])); w.w = __fadd_rn(w.w, __fmul_rn(neg_lr, qb[3]));
      } else {                                                                         // :618, :622
        w.x = __fadd_rn(w.x, __fmul_rn(neg_lr, __fmul_rn(qb[0], s_bar)));
        w.y = __fadd_rn(w.y, __fmul_rn(neg_lr, __fmul_rn(qb[1], s_bar)));
        w.z = __fadd_rn(w.z, __fmul_rn(neg_lr, __fmul_rn(qb[2], s_bar)));
        w.w = __fadd_rn(w.w, __fmul_rn(neg_lr, __fmul_rn(qb[3], s_bar)));
      }
      w4[col] = w;
    }
  }
  if (bad) atomicOr(status, bad);
}

}  // namespace dqrm

using namespace dqrm;

extern "C" size_t dqrm_slot_bytes(int num_tables, int64_t capacity, int dim, int bits) {
  return slot_layout(num_tables, capacity, dim, bits).bytes;
}

extern "C" int dqrm_slot_layout(int num_tables, int64_t capacity, int dim, int bits, size_t* rows_offset,
                                size_t* codes_offset) {
  DQRM_REQUIRE(rows_offset && codes_offset, -EINVAL, "slot_layout: null argument");
  const SlotLayout l = slot_layout(num_tables, capacity, dim, bits);
  *rows_offset = l.rows_off;
  *codes_offset = l.codes_off;
  return 0;
}

extern "C" int dqrm_grad_pack(int num_tables, int dim, const float* grad_sums, const int32_t* uniq_rows,
                              const int32_t* uniq_count, int64_t capacity,
                              const float* gathered_scales, int64_t scale_stride_elems, int world, int bits,
                              void* slot, float* scale_mean, void* stream) {
  DQRM_REQUIRE(scale_stride_elems >= num_tables, -EINVAL, "grad_pack: scale_stride_elems=%lld < num_tables", (long long)scale_stride_elems);
  DQRM_REQUIRE(grad_sums && uniq_rows && uniq_count && gathered_scales && slot && scale_mean, -EINVAL, "grad_pack: null argument");
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "grad_pack: num_tables=%d", num_tables);
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "grad_pack: dim=%d", dim);
  DQRM_REQUIRE((bits >= 2 && bits <= 16) || bits == 32, -EINVAL, "grad_pack: bits=%d outside [2,16] and != 32", bits);
  DQRM_REQUIRE(world >= 1 && capacity >= 1, -EINVAL, "grad_pack: world=%d capacity=%lld", world, (long long)capacity);
  DQRM_REQUIRE((reinterpret_cast<uintptr_t>(slot) & 15u) == 0, -EINVAL, "grad_pack: slot not 16-byte aligned");
  const SlotLayout lay = slot_layout(num_tables, capacity, dim, bits);
  const RowLanes rl = row_lanes(dim);
  long long blocks = ceil_div(capacity, 256 / rl.group);
  const long long cap_blocks = ceil_div(16ll * kSMs, num_tables);        // ~16 CTAs per SM over all tables: the rows of one
  if (blocks > cap_blocks) blocks = cap_blocks;                          // big table need many loads in flight
  dim3 grid((unsigned)blocks, num_tables);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float inv_world = (float)(1.0 / world);
#define DQRM_PACK(COLS, CT)                                                                                       \
  grad_pack_kernel<COLS, CT><<<grid, 256, 0, st>>>(num_tables, dim / 4, rl.group, grad_sums, uniq_rows, uniq_count, \
                                                   capacity, gathered_scales, scale_stride_elems, world, inv_world, bits, \
                                                   static_cast<unsigned char*>(slot), lay, scale_mean)
  if (bits == 32) { if (rl.cols == 1) DQRM_PACK(1, float); else if (rl.cols == 2) DQRM_PACK(2, float); else DQRM_PACK(4, float); }
  else if (bits <= 8) { if (rl.cols == 1) DQRM_PACK(1, int8_t); else if (rl.cols == 2) DQRM_PACK(2, int8_t); else DQRM_PACK(4, int8_t); }
  else           { if (rl.cols == 1) DQRM_PACK(1, int16_t); else if (rl.cols == 2) DQRM_PACK(2, int16_t); else DQRM_PACK(4, int16_t); }
#undef DQRM_PACK
  DQRM_LAUNCH_CHECK("grad_pack_kernel");
  return 0;
}

extern "C" int dqrm_grad_merge_apply(int num_tables, float* const* weight, const int64_t* rows, int dim,
                                     const void* gathered, int world, int64_t capacity, int bits,
                                     const float* scale_mean, float lr, const float* lr_dev,
                                     int32_t* updated_rows, int32_t* updated_count, float* qbar,
                                     int32_t* status, void* stream) {
  DQRM_REQUIRE(weight && rows && gathered && scale_mean && status, -EINVAL, "grad_merge_apply: null argument");
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "grad_merge_apply: dim=%d", dim);
  DQRM_REQUIRE((bits >= 2 && bits <= 16) || bits == 32, -EINVAL, "grad_merge_apply: bits=%d outside [2,16] and != 32", bits);
  DQRM_REQUIRE(world >= 1 && world <= 65535 && capacity >= 1, -EINVAL, "grad_merge_apply: world=%d capacity=%lld", world, (long long)capacity);
  DQRM_REQUIRE((updated_rows == nullptr) == (updated_count == nullptr), -EINVAL, "grad_merge_apply: updated_rows/updated_count must come together");
  DQRM_REQUIRE(!qbar || updated_rows, -EINVAL, "grad_merge_apply: qbar needs updated_rows");
  TableSet ts;
  if (int rc = fill_tables(ts, num_tables, weight, rows, nullptr)) return rc;
  const SlotLayout lay = slot_layout(num_tables, capacity, dim, bits);
  const RowLanes rl = row_lanes(dim);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (updated_count) {
    cudaError_t e = cudaMemsetAsync(updated_count, 0, sizeof(int32_t) * num_tables, st);
    DQRM_REQUIRE(e == cudaSuccess, -EIO, "grad_merge_apply: memset failed: %s", cudaGetErrorString(e));
  }
  long long blocks = ceil_div(capacity, 256 / rl.group);
  const long long cap_blocks = ceil_div(16ll * kSMs, (long long)num_tables * world);   // random-row RMW: latency-bound,
  if (blocks > cap_blocks) blocks = cap_blocks;                                        // so fill the SMs
  dim3 grid((unsigned)blocks, world, num_tables);
  const float inv_world = (float)(1.0 / world);
  const float neg_lr = -lr;
  int search_iters = 0;                                       // a binary search over <= capacity entries ends in this many halvings
  while ((1ll << search_iters) <= capacity) ++search_iters;
#define DQRM_MERGE(COLS, CT)                                                                                         \
  do {                                                                                                               \
    if (world > 1)                                                                                                   \
      grad_merge_apply_kernel<COLS, CT, true><<<grid, 256, 0, st>>>(ts, dim / 4, rl.group, static_cast<const unsigned char*>(gathered), \
                                                                    lay, world, capacity, scale_mean, neg_lr, lr_dev, inv_world, \
                                                                    updated_rows, updated_count, qbar, status, search_iters); \
    else                                                                                                             \
      grad_merge_apply_kernel<COLS, CT, false><<<grid, 256, 0, st>>>(ts, dim / 4, rl.group, static_cast<const unsigned char*>(gathered), \
                                                                     lay, world, capacity, scale_mean, neg_lr, lr_dev, inv_world, \
                                                                     updated_rows, updated_count, qbar, status, search_iters); \
  } while (0)
  if (bits == 32) { if (rl.cols == 1) DQRM_MERGE(1, float); else if (rl.cols == 2) DQRM_MERGE(2, float); else DQRM_MERGE(4, float); }
  else if (bits <= 8) { if (rl.cols == 1) DQRM_MERGE(1, int8_t); else if (rl.cols == 2) DQRM_MERGE(2, int8_t); else DQRM_MERGE(4, int8_t); }
  else           { if (rl.cols == 1) DQRM_MERGE(1, int16_t); else if (rl.cols == 2) DQRM_MERGE(2, int16_t); else DQRM_MERGE(4, int16_t); }
#undef DQRM_MERGE
  DQRM_LAUNCH_CHECK("grad_merge_apply_kernel");
  return 0;
}
