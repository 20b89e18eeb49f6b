// Exact incremental tracker for the per-table max-abs ("next").
//
// The reference rescans every table on every forward (quant_utils.py:177-178; the period bookkeeping that
// would amortise it is commented out, quant_modules_not_quantize_grad.py:354-363).  A step only changes the
// rows it looked up, so max|W| can be maintained exactly: keep the max of every block of `block_rows` rows,
// after each update recompute only the blocks that contain an updated row (from the tables themselves, so
// decreases are handled), and reduce the small block-max array.  max is exact and order-free, therefore the
// resulting scale is BIT-IDENTICAL to a full rescan while the bytes moved per step drop from N*D*4 to
// (touched blocks)*block_rows*D*4 + (N/block_rows)*4  (Kaggle, batch 128: 2.16 GB -> <= 13.6 MB + 2.1 MB).
// The block maxima are stored as the bit pattern of |x| (a non-negative float), so the final reduction is
// the ordinary table_absmax_kernel run over the block-max arrays.
#include <stdlib.h>

#include "common.cuh"

namespace dqrm {

struct TrackArgs {
  const float* w[DQRM_MAX_TABLES];
  float* bm[DQRM_MAX_TABLES];          // block maxima of table k
  long long rows[DQRM_MAX_TABLES];
  long long blk_begin[DQRM_MAX_TABLES + 1];   // prefix of block counts (build mode)
  int blk_lo[DQRM_MAX_TABLES];         // this rank's block shard [blk_lo, blk_hi) of table k (whole table when unsharded)
  int blk_hi[DQRM_MAX_TABLES];
  int num_tables;
};

__device__ __forceinline__ void block_max_of(const float* __restrict__ w, long long rows, int dim, int block_rows,
                                             long long blk, float* __restrict__ bm, int lane) {
  const long long r0 = blk * block_rows;
  const long long r1 = min(rows, r0 + block_rows);
  const long long n = (r1 - r0) * dim;
  const float* p = w + r0 * dim;
  unsigned m = 0u;
  if ((dim & 3) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (long long i = lane; i < (n >> 2); i += 32) m = max(m, abs_bits4(__ldg(p4 + i)));
  } else {
    for (long long i = lane; i < n; i += 32) m = max(m, abs_bits(__ldg(p + i)));
  }
  m = warp_max_u32(m);
  if (lane == 0) bm[blk] = __uint_as_float(m);
}

// build: one warp per block over all tables
__global__ void __launch_bounds__(256)
blockmax_build_kernel(const __grid_constant__ TrackArgs a, int dim, int block_rows) {
  const int lane = threadIdx.x & 31;
  const long long total = a.blk_begin[a.num_tables];
  for (long long gb = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); gb < total;
       gb += (long long)gridDim.x * (blockDim.x >> 5)) {
    int t = 0;
    while (gb >= a.blk_begin[t + 1]) ++t;
    block_max_of(a.w[t], a.rows[t], dim, block_rows, gb - a.blk_begin[t], a.bm[t], lane);
  }
}

// update: one warp per (rank, table, entry) of the gathered exchange slots (or of a local row list when
// world == 0): recompute the block that holds the row.  Several entries may hit one block: they write the
// same value.
__global__ void __launch_bounds__(256)
blockmax_update_kernel(const __grid_constant__ TrackArgs a, int dim, int block_rows,
                       const unsigned char* __restrict__ gathered, size_t slot_bytes, size_t rows_off, int world,
                       const int* __restrict__ uniq_rows, const int* __restrict__ uniq_count, long long capacity) {
  const int lane = threadIdx.x & 31;
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
  for (int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < U; j += gridDim.x * (blockDim.x >> 5)) {
    const long long row = rows[j];
    if (row < 0 || row >= a.rows[t]) continue;
    const long long blk = row / block_rows;
    if (blk < a.blk_lo[t] || blk >= a.blk_hi[t]) continue;      // another rank's shard of the pipelined scan
    block_max_of(a.w[t], a.rows[t], dim, block_rows, blk, a.bm[t], lane);
  }
}


// ---- pipelined full rescan ---------------------------------------------------------------------
// The reference's period-1 rescan (quant_utils.py:177-178) reads every table byte on every step; serialised in
// front of the forward it is 64 % of a Kaggle-shape step.  The rescan does not have to sit on the critical path:
// the table is read-only from the end of one update to the start of the next, so the pass that produces the scale
// of step i+1 runs DURING step i (low-priority stream) and writes one max per block of `block_rows` rows; after
// update i the blocks that hold an updated row are recomputed (blockmax_update_kernel) and the block maxima are
// reduced (blockmax_reduce_kernel).  max is exact and order-free, so the scale is bit-identical to the serial
// rescan, every byte is still read once per step, and no state is carried from one step to the next.
//
// Occupancy is capped on purpose (persistent grid of kSMs * ctas_per_sm CTAs): HBM saturates with ~40-64 KiB in
// flight per SM, and every byte queued beyond that only adds queueing delay (Little's law: 256 KiB/SM outstanding
// = 38 MB = 5.6 us at 6.8 TB/s) to each dependent load of the step's latency-bound kernels that run concurrently.
// The capped grid also leaves threads, registers and all shared memory of every SM to those kernels.
// Loads are 256-bit (LDG.E.256, sm_100+) with L1 no-allocate and L2 evict-first, so the 2-48 GB stream does not
// displace the step's small working set from the 126 MB L2.
constexpr int kPipeThreads = 256;
constexpr int kPipeWarps = kPipeThreads / 32;

struct PipeArgs {
  const float* w[DQRM_MAX_TABLES];
  float* bm[DQRM_MAX_TABLES];
  long long rows[DQRM_MAX_TABLES];
  int blk_lo[DQRM_MAX_TABLES];              // first block of this rank's shard (table-local index)
  int gblk_begin[DQRM_MAX_TABLES + 1];      // prefix over tables of the shard's block counts
  int num_tables;
};

struct alignas(32) Vec8 { unsigned v[8]; };

__device__ __forceinline__ Vec8 ld_stream_256(const Vec8* p) {
  Vec8 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned abs_bits8(const Vec8& x) {
  unsigned m = 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i) m = max(m, x.v[i] & 0x7fffffffu);
  return m;
}

// kInFlight loads of one lane, issued back to back from ONE asm statement so that ptxas cannot interleave
// them with the max-folds to save registers (it otherwise keeps only ~2 in flight); the lane's elements are
// 32 vector elements (one warp-wide row of loads) apart.
template <bool WIDE> struct PipeVec;
template <> struct PipeVec<true> {
  using T = Vec8;
  static constexpr int kInFlight = 4;                        // 4 x 32 B per lane = one 64-row x dim-16 block per warp pass
  static __device__ __forceinline__ T ld(const T* p) { return ld_stream_256(p); }
  static __device__ __forceinline__ unsigned amax(const T& v) { return abs_bits8(v); }
  static __device__ __forceinline__ void ld_batch(const T* p, T (&r)[kInFlight]) {
#define DQRM_V8(e) "=r"(e.v[0]), "=r"(e.v[1]), "=r"(e.v[2]), "=r"(e.v[3]), "=r"(e.v[4]), "=r"(e.v[5]), "=r"(e.v[6]), "=r"(e.v[7])
    asm volatile(
        "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%32];\n\t"
        "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%32+1024];\n\t"
        "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%32+2048];\n\t"
        "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%24,%25,%26,%27,%28,%29,%30,%31}, [%32+3072];"
        : DQRM_V8(r[0]), DQRM_V8(r[1]), DQRM_V8(r[2]), DQRM_V8(r[3])
        : "l"(p));
#undef DQRM_V8
  }
};
template <> struct PipeVec<false> {
  using T = float4;
  static constexpr int kInFlight = 8;
  static __device__ __forceinline__ T ld(const T* p) { return ld_stream_f4(p); }
  static __device__ __forceinline__ unsigned amax(const T& v) { return abs_bits4(v); }
  static __device__ __forceinline__ void ld_batch(const T* p, T (&r)[kInFlight]) {
#define DQRM_V4(e) "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w)
    asm volatile(
        "ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%32];\n\t"
        "ld.global.nc.L1::no_allocate.v4.f32 {%4,%5,%6,%7}, [%32+512];\n\t"
        "ld.global.nc.L1::no_allocate.v4.f32 {%8,%9,%10,%11}, [%32+1024];\n\t"
        "ld.global.nc.L1::no_allocate.v4.f32 {%12,%13,%14,%15}, [%32+1536];\n\t"
        "ld.global.nc.L1::no_allocate.v4.f32 {%16,%17,%18,%19}, [%32+2048];\n\t"
        "ld.global.nc.L1::no_allocate.v4.f32 {%20,%21,%22,%23}, [%32+2560];\n\t"
        "ld.global.nc.L1::no_allocate.v4.f32 {%24,%25,%26,%27}, [%32+3072];\n\t"
        "ld.global.nc.L1::no_allocate.v4.f32 {%28,%29,%30,%31}, [%32+3584];"
        : DQRM_V4(r[0]), DQRM_V4(r[1]), DQRM_V4(r[2]), DQRM_V4(r[3]), DQRM_V4(r[4]), DQRM_V4(r[5]), DQRM_V4(r[6]),
          DQRM_V4(r[7])
        : "l"(p));
#undef DQRM_V4
  }
};

// dimv = vector elements per row (dim/8 when WIDE, dim/4 otherwise); a "unit" is blocks_per_unit consecutive
// blocks of the shard's concatenated block list, one warp per unit, consecutive warps on consecutive units.
template <bool WIDE>
__global__ void __launch_bounds__(kPipeThreads, 4)
blockmax_scan_kernel(const __grid_constant__ PipeArgs a, int dimv, int block_rows, int blocks_per_unit, int total_units) {
  using PV = PipeVec<WIDE>;
  using VT = typename PV::T;
  constexpr int V = PV::kInFlight;
  const int lane = threadIdx.x & 31;
  const int total = a.gblk_begin[a.num_tables];
  const int fullv = block_rows * dimv;       // vector elements of a complete block
  for (int unit = blockIdx.x * kPipeWarps + (threadIdx.x >> 5); unit < total_units; unit += gridDim.x * kPipeWarps) {
    int g = unit * blocks_per_unit;
    const int g_end = min(total, g + blocks_per_unit);
    int t;                                   // table of block g: last k with gblk_begin[k] <= g (warp-uniform)
    {
      int lo = 0, hi = a.num_tables - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.gblk_begin[mid] <= g) lo = mid; else hi = mid - 1;
      }
      t = lo;
    }
    for (; g < g_end; ++g) {
      while (g >= a.gblk_begin[t + 1]) ++t;  // skips empty shards too
      const long long blk = (long long)(g - a.gblk_begin[t]) + a.blk_lo[t];
      const long long r0 = blk * block_rows;
      const long long left = a.rows[t] - r0;
      const int nv = left >= block_rows ? fullv : (int)left * dimv;
      const VT* p = reinterpret_cast<const VT*>(a.w[t]) + r0 * dimv + lane;
      unsigned m = 0u;
      int i = 0;
      for (; i + 32 * V <= nv; i += 32 * V) {
        VT v[V];
        PV::ld_batch(p + i, v);
#pragma unroll
        for (int j = 0; j < V; ++j) m = max(m, PV::amax(v[j]));
      }
      for (i += lane; i < nv; i += 32) m = max(m, PV::amax(PV::ld(p + i - lane)));
      m = warp_max_u32(m);
      if (lane == 0) a.bm[t][blk] = __uint_as_float(m);
    }
  }
}

struct ReduceArgs {
  const float* bm[DQRM_MAX_TABLES];          // already offset to the shard's first block
  int n[DQRM_MAX_TABLES];
  int num_tables;
};

// block maxima -> per-table absmax (and scale, 1/scale): grid (chunks, tables); the last CTA finalises and
// re-zeros the workspace exactly like table_absmax_kernel (scan.cu).
__global__ void __launch_bounds__(256)
blockmax_reduce_kernel(const __grid_constant__ ReduceArgs a, unsigned* __restrict__ acc, unsigned* __restrict__ counter,
                       float* __restrict__ absmax_out, float* __restrict__ scale_out, float* __restrict__ inv_out,
                       int bits) {
  __shared__ unsigned s_max;
  __shared__ bool s_last;
  const int t = blockIdx.y;
  const float* __restrict__ p = a.bm[t];
  const int n = a.n[t];
  unsigned m = 0u;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) m = max(m, abs_bits(p[i]));
  const unsigned bmx = block_max_u32(m, &s_max);
  if (threadIdx.x == 0 && bmx) atomicMax(&acc[t], bmx);
  __threadfence();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (s_last) {
    for (int k = threadIdx.x; k < a.num_tables; k += 256) {
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

// balanced contiguous split of nb blocks over `world` ranks (same rule as get_my_slice,
// dlrm_s_pytorch_comm_grad.py:993-997, applied to blocks)
static inline void block_shard(long long nb, int rank, int world, int* lo, int* hi) {
  const long long q = nb / world, r = nb % world;
  const long long l = rank * q + (rank < r ? rank : r);
  *lo = (int)l;
  *hi = (int)(l + q + (rank < r ? 1 : 0));
}


static int fill_track(TrackArgs& a, int num_tables, const float* const* weight, const int64_t* rows, int dim,
                      int block_rows, float* const* blockmax, int shard_rank = 0, int shard_world = 1) {
  DQRM_REQUIRE(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world, -EINVAL, "blockmax: shard %d/%d",
               shard_rank, shard_world);
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "blockmax: num_tables=%d", num_tables);
  DQRM_REQUIRE(weight && rows && blockmax && dim >= 1 && block_rows >= 1, -EINVAL, "blockmax: bad argument");
  a.num_tables = num_tables;
  long long nb = 0;
  for (int k = 0; k < num_tables; ++k) {
    DQRM_REQUIRE(weight[k] && blockmax[k] && rows[k] >= 0, -EINVAL, "blockmax: table %d malformed", k);
    DQRM_REQUIRE((dim & 3) != 0 || (reinterpret_cast<uintptr_t>(weight[k]) & 15u) == 0, -EINVAL,
                 "blockmax: table %d not 16-byte aligned", k);
    a.w[k] = weight[k]; a.bm[k] = blockmax[k]; a.rows[k] = rows[k];
    a.blk_begin[k] = nb;
    const long long nbk = ceil_div(rows[k], block_rows);
    DQRM_REQUIRE(nbk < (1ll << 31), -E2BIG, "blockmax: table %d has too many blocks", k);
    block_shard(nbk, shard_rank, shard_world, &a.blk_lo[k], &a.blk_hi[k]);
    nb += nbk;
  }
  a.blk_begin[num_tables] = nb;
  return 0;
}

}  // namespace dqrm

using namespace dqrm;

extern "C" int64_t dqrm_blockmax_entries(int64_t rows, int block_rows) {
  if (rows < 0 || block_rows < 1) return -1;
  return (rows + block_rows - 1) / block_rows;
}

extern "C" int dqrm_blockmax_build(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                   int block_rows, float* const* blockmax, void* stream) {
  TrackArgs a;
  if (int rc = fill_track(a, num_tables, weight, rows, dim, block_rows, blockmax)) return rc;
  long long warps = a.blk_begin[num_tables];
  long long grid = ceil_div(warps, 8);
  if (grid > 16ll * kSMs) grid = 16ll * kSMs;
  if (grid < 1) return 0;
  blockmax_build_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, dim, block_rows);
  DQRM_LAUNCH_CHECK("blockmax_build_kernel");
  return 0;
}

extern "C" int dqrm_blockmax_update_shard(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                          int block_rows, float* const* blockmax,
                                          const void* gathered, int world, int64_t capacity, int bits,
                                          const int32_t* uniq_rows, const int32_t* uniq_count,
                                          int shard_rank, int shard_world, void* stream) {
  TrackArgs a;
  if (int rc = fill_track(a, num_tables, weight, rows, dim, block_rows, blockmax, shard_rank, shard_world)) return rc;
  DQRM_REQUIRE((gathered != nullptr) != (uniq_rows != nullptr), -EINVAL,
               "blockmax_update: pass either the gathered slots or a local row list");
  DQRM_REQUIRE(capacity >= 1, -EINVAL, "blockmax_update: capacity=%lld", (long long)capacity);
  size_t slot_bytes = 0, rows_off = 0, codes_off = 0;
  int ranks = 1;
  if (gathered) {
    DQRM_REQUIRE(world >= 1 && world <= 65535 && ((bits >= 2 && bits <= 16) || bits == 32), -EINVAL, "blockmax_update: world/bits");
    slot_bytes = dqrm_slot_bytes(num_tables, capacity, dim, bits);
    dqrm_slot_layout(num_tables, capacity, dim, bits, &rows_off, &codes_off);
    ranks = world;
  } else {
    DQRM_REQUIRE(uniq_count, -EINVAL, "blockmax_update: uniq_count missing");
  }
  long long gx = ceil_div(capacity, 8);
  if (gx > 64) gx = 64;
  dim3 grid((unsigned)gx, ranks, num_tables);
  blockmax_update_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, dim, block_rows, static_cast<const unsigned char*>(gathered), slot_bytes, rows_off, world, uniq_rows, uniq_count,
      capacity);
  DQRM_LAUNCH_CHECK("blockmax_update_kernel");
  return 0;
}

extern "C" int dqrm_blockmax_update(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                    int block_rows, float* const* blockmax,
                                    const void* gathered, int world, int64_t capacity, int bits,
                                    const int32_t* uniq_rows, const int32_t* uniq_count, void* stream) {
  return dqrm_blockmax_update_shard(num_tables, weight, rows, dim, block_rows, blockmax, gathered, world, capacity, bits,
                                    uniq_rows, uniq_count, 0, 1, stream);
}

// Resident CTAs per SM of the pipelined pass (see blockmax_scan_kernel).  Default from the B200 sweep in
// profiles/; DQRM_PIPE_CTAS_PER_SM overrides it for tuning (0 = uncapped, one visit per CTA).
static int pipe_ctas_per_sm() {
  static const int v = [] {
    const char* e = getenv("DQRM_PIPE_CTAS_PER_SM");
    int x = e ? atoi(e) : 2;
    return x < 0 ? 0 : (x > 8 ? 8 : x);
  }();
  return v;
}

extern "C" int dqrm_blockmax_scan(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                  int block_rows, float* const* blockmax, int shard_rank, int shard_world,
                                  void* stream) {
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "blockmax_scan: num_tables=%d", num_tables);
  DQRM_REQUIRE(weight && rows && blockmax && block_rows >= 1, -EINVAL, "blockmax_scan: bad argument");
  DQRM_REQUIRE(dim >= 4 && (dim & 3) == 0, -EINVAL, "blockmax_scan: dim=%d must be a multiple of 4", dim);
  DQRM_REQUIRE(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world, -EINVAL, "blockmax_scan: shard %d/%d",
               shard_rank, shard_world);
  PipeArgs a;
  a.num_tables = num_tables;
  long long total = 0;
  for (int k = 0; k < num_tables; ++k) {
    DQRM_REQUIRE(weight[k] && blockmax[k] && rows[k] >= 0, -EINVAL, "blockmax_scan: table %d malformed", k);
    DQRM_REQUIRE((reinterpret_cast<uintptr_t>(weight[k]) & 15u) == 0, -EINVAL, "blockmax_scan: table %d not 16-byte aligned", k);
    const long long nbk = ceil_div(rows[k], block_rows);
    DQRM_REQUIRE(nbk < (1ll << 31), -E2BIG, "blockmax_scan: table %d has too many blocks", k);
    int lo, hi;
    block_shard(nbk, shard_rank, shard_world, &lo, &hi);
    a.w[k] = weight[k]; a.bm[k] = blockmax[k]; a.rows[k] = rows[k];
    a.blk_lo[k] = lo;
    a.gblk_begin[k] = (int)total;
    total += hi - lo;
    DQRM_REQUIRE(total < (1ll << 31), -E2BIG, "blockmax_scan: too many blocks");
  }
  a.gblk_begin[num_tables] = (int)total;
  if (total == 0) return 0;
  bool wide = (dim & 7) == 0;                                // 256-bit loads need 32-byte aligned rows and bases
  for (int k = 0; k < num_tables && wide; ++k) wide = (reinterpret_cast<uintptr_t>(weight[k]) & 31u) == 0;
  const long long block_bytes = (long long)block_rows * dim * 4;
  long long bpu = 8192 / block_bytes;                        // ~8 KiB per warp visit, 64 KiB contiguous per CTA visit
  if (bpu < 1) bpu = 1;
  const long long units = ceil_div(total, bpu);
  long long grid = ceil_div(units, kPipeWarps);
  const int cap = pipe_ctas_per_sm();
  if (cap > 0 && grid > (long long)cap * kSMs) grid = (long long)cap * kSMs;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // The pass uses no shared memory; ask for the largest shared-memory carve-out anyway so that the SMs it occupies
  // stay configured for kernels of the step that do need shared memory (the pass streams with L1 no-allocate).
  static const bool carveout_set = [] {
    cudaFuncSetAttribute(blockmax_scan_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(blockmax_scan_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    return true;
  }();
  (void)carveout_set;
  if (wide)
    blockmax_scan_kernel<true><<<(unsigned)grid, kPipeThreads, 0, st>>>(a, dim / 8, block_rows, (int)bpu, (int)units);
  else
    blockmax_scan_kernel<false><<<(unsigned)grid, kPipeThreads, 0, st>>>(a, dim / 4, block_rows, (int)bpu, (int)units);
  DQRM_LAUNCH_CHECK("blockmax_scan_kernel");
  return 0;
}

extern "C" int dqrm_blockmax_reduce(int num_tables, const int64_t* rows, int block_rows, const float* const* blockmax,
                                    int shard_rank, int shard_world, int bits, float* absmax, float* scale,
                                    float* inv_scale, void* workspace, void* stream) {
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "blockmax_reduce: num_tables=%d", num_tables);
  DQRM_REQUIRE(rows && blockmax && absmax && workspace && block_rows >= 1, -EINVAL, "blockmax_reduce: bad argument");
  DQRM_REQUIRE((scale == nullptr) == (inv_scale == nullptr), -EINVAL, "blockmax_reduce: scale/inv_scale must both be set or both NULL");
  DQRM_REQUIRE(!scale || (bits >= 2 && bits <= 16), -EINVAL, "blockmax_reduce: bits=%d outside [2,16]", bits);
  DQRM_REQUIRE(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world, -EINVAL, "blockmax_reduce: shard %d/%d",
               shard_rank, shard_world);
  ReduceArgs a;
  a.num_tables = num_tables;
  for (int k = 0; k < num_tables; ++k) {
    DQRM_REQUIRE(blockmax[k] && rows[k] >= 0, -EINVAL, "blockmax_reduce: table %d malformed", k);
    const long long nbk = ceil_div(rows[k], block_rows);
    DQRM_REQUIRE(nbk < (1ll << 31), -E2BIG, "blockmax_reduce: table %d has too many blocks", k);
    int lo, hi;
    block_shard(nbk, shard_rank, shard_world, &lo, &hi);
    a.bm[k] = blockmax[k] + lo;
    a.n[k] = hi - lo;
  }
  unsigned* acc = static_cast<unsigned*>(workspace);          // dqrm_scan_workspace_bytes(num_tables), zeroed once
  dim3 grid(32, num_tables);
  blockmax_reduce_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, acc, acc + num_tables, absmax, scale,
                                                                              inv_scale, bits);
  DQRM_LAUNCH_CHECK("blockmax_reduce_kernel");
  return 0;
}
