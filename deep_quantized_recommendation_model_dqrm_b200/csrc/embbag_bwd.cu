// (a4 bwd, a5, a7 steps 1-2) sparse row gradient of the QAT EmbeddingBag, de-duplicated.
// Reference: autograd of quantization_supp/quant_modules_not_quantize_grad.py:393 and
//            SymmetricQuantFunction.backward (quantization_supp/quant_utils.py:348-363): dy = (g*s)/s;
//            ATen sparse EmbeddingBag backward (dlrm_s_pytorch_comm_grad.py:1938): one value row per lookup;
//            Tensor.coalesce + gradient scale (sgd_quantized_gradients_parallel_comm.py:859-861).
//
// One CTA per table.  The table's lookups become 64-bit keys (row << 32 | bag) in
// shared memory; because a key carries its bag, equal rows are ordered by bag =
// original lookup order, which makes the (unstable) bitonic network a stable
// sort by row.  Head flags + a block scan give the ascending unique rows; a
// dim/4-lane group then folds each row's duplicates left-to-right, reading
// dy = (g*s)/s straight from dOut (the per-lookup [L,D] value matrix the
// reference materialises never exists), and the CTA finishes with the table's
// max|sum| -> 8-bit gradient scale.  Rows with more than DQRM_FOLD_BLOCK (64) duplicates -- the 3- and 4-row tables
// of the Criteo sets collect thousands at batch 8192 -- are folded in blocks of 64 consecutive lookups by different
// lane groups in parallel and the block sums are folded left to right: still one fixed summation order (the
// oracle's coalesce_spec defines exactly this shape), but a row's latency is one block, not the whole chain.  Sort, unique, segmented sum and scale are
// ONE launch for all tables (the reference: index_select, thrust sort,
// coalesceValuesKernel, 4 reductions and a host sync per table).
// Optionally (DQRM_BWD_CLUSTER=2|4|8, off by default -- see the launch code) a thread-block CLUSTER per table: the
// sort stays in CTA 0, whose keys and segment starts the other CTAs of the cluster read through distributed shared
// memory while they fold their share of the unique rows.  Same fold order per row, so the same bits.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace dqrm {

constexpr unsigned long long kPadKey = ~0ull;
constexpr int kBwdPrefetchRows = 8;   // float4 row loads in flight per lane group (divided by COLS)
constexpr int kFoldBlock = DQRM_FOLD_BLOCK;
constexpr int kMaxLongRows = DQRM_BWD_CTA_MAX_LOOKUPS / (kFoldBlock + 1) + 4;   // rows with > kFoldBlock duplicates

struct BwdArgs {
  long long rows[DQRM_MAX_TABLES];
  long long idx_begin[DQRM_MAX_TABLES + 1];
};

__device__ __forceinline__ float ste_dy(float g, float s, bool quant) {
  return quant ? __fdiv_rn(__fmul_rn(g, s), s) : g;
}

template <int COLS>
__global__ void __launch_bounds__(1024)
embbag_bwd_cta_kernel(const __grid_constant__ BwdArgs a, int dim4, int group,
                      const long long* __restrict__ indices, const long long* __restrict__ offsets, long long bags,
                      const float* __restrict__ dout, long long dts, long long dbs,
                      const float* __restrict__ fwd_scale, long long capacity,
                      int* __restrict__ uniq_rows, int* __restrict__ uniq_count, float* __restrict__ grad_sums,
                      int grad_bits, float* __restrict__ grad_scale_local, int* __restrict__ status,
                      float* __restrict__ partials, long long partial_items, int csize, int radix) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_warp_tot[32];
  __shared__ int s_nvalid, s_unique, s_nlong;
  __shared__ unsigned s_max;
  __shared__ int s_long_j[kMaxLongRows], s_long_start[kMaxLongRows + 1];

  const int t = blockIdx.x / csize, cr = blockIdx.x % csize;              // table, rank of this CTA in the table's cluster
  const int tid = threadIdx.x, nthr = blockDim.x;
  const long long L = a.idx_begin[t + 1] - a.idx_begin[t];
  int n = 2;
  while (n < L) n <<= 1;
  // shared memory: keys [n] | (radix: second key buffer [n]) | seg_start [n + 4] | (radix: per-warp digit counters, digit totals, digit bases)
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  int* seg_start = reinterpret_cast<int*>(keys + (radix ? 2 * n : n));
  const long long* idx = indices + a.idx_begin[t];
  const long long* off = offsets + (long long)t * bags;
  const long long nrows = a.rows[t];
  int bad = 0;
  if (tid == 0) { s_max = 0u; s_nlong = 0; }
  int U = 0;
  if (cr == 0) {                                                         // sort + segments: the cluster's first CTA only

  // 1. keys (lookups no bag covers keep the pad: it sorts last -- the radix form gives it row id `nrows`)
  const unsigned long long pad = radix ? (((unsigned long long)nrows << 32) | 0xffffffffull) : kPadKey;
  for (int i = tid; i < n; i += nthr) keys[i] = pad;
  __syncthreads();
  for (long long b = tid; b < bags; b += nthr) {
    long long start = off[b];
    long long end = (b + 1 < bags) ? off[b + 1] : L;
    if (start < 0 || end > L || start > end) {
      bad |= DQRM_STATUS_OFFSET_ORDER;
      start = start < 0 ? 0 : (start > L ? L : start);
      end = end > L ? L : (end < start ? start : end);
    }
    for (long long l = start; l < end; ++l) {
      long long r = idx[l];
      if (r < 0 || r >= nrows) { bad |= DQRM_STATUS_INDEX_RANGE; r = r < 0 ? 0 : nrows - 1; }
      keys[l] = ((unsigned long long)r << 32) | (unsigned long long)(unsigned)b;
    }
  }
  __syncthreads();

  // 2. sort ascending by (row, lookup order).  From 1024 keys: stable LSD radix sort by row in shared memory, 8 bits per
  // pass over the bits of `nrows` (3-4 passes of five block barriers each; the bitonic network needs 55-105 stages) --
  // the scheme of embbag_bwd_large.cu inside one CTA: a warp walks its contiguous keys 32 at a time, rank = running
  // per-warp digit counter + match_any rank.  Below that, and from 8192 keys (no room for a second buffer): bitonic.
  if (radix) {
    unsigned long long* kin = keys;
    unsigned long long* kout = keys + n;
    unsigned* wcnt = reinterpret_cast<unsigned*>(seg_start + n + 4);       // [nwarps][256]
    const int nwarps = nthr >> 5, warp = tid >> 5, ln = tid & 31;
    unsigned* hist = wcnt + nwarps * 256;                                  // [256] digit totals
    unsigned* base = hist + 256;                                           // [256] exclusive digit bases
    const int Li = (int)L, ipt = (Li + nthr - 1) / nthr, wbeg = warp * 32 * ipt;   // ipt <= 8
    const unsigned ltm = (1u << ln) - 1u;
    int key_bits = 1;
    while ((1ll << key_bits) <= nrows) ++key_bits;                         // covers the pad's row id
    for (int shift = 0; shift < key_bits; shift += 8) {
      for (int j = tid; j < nwarps * 256; j += nthr) wcnt[j] = 0u;
      __syncthreads();
      unsigned loc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k >= ipt) break;                                               // (block-uniform)
        const int i = wbeg + k * 32 + ln;
        const bool live = i < Li;
        const unsigned d = live ? (((unsigned)(kin[i] >> 32) >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rank = __popc(peers & ltm);
        const unsigned prev = live ? wcnt[warp * 256 + d] : 0u;
        loc[k] = prev + rank;
        __syncwarp();
        if (live && rank == 0) wcnt[warp * 256 + d] = prev + __popc(peers);
        __syncwarp();
      }
      __syncthreads();
      for (int d = tid; d < 256; d += nthr) {                              // exclusive prefix over the warps, per digit
        unsigned run = 0;
        for (int w = 0; w < nwarps; ++w) { const unsigned c = wcnt[w * 256 + d]; wcnt[w * 256 + d] = run; run += c; }
        hist[d] = run;
      }
      __syncthreads();
      if (tid < 32) {                                                      // exclusive scan of the 256 digit totals: 8 per lane
        unsigned v[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = hist[tid * 8 + i]; sum += v[i]; }
        unsigned incl = sum;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, sh); if (tid >= sh) incl += u; }
        unsigned ex = incl - sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) { base[tid * 8 + i] = ex; ex += v[i]; }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k >= ipt) break;
        const int i = wbeg + k * 32 + ln;
        if (i < Li) {
          const unsigned long long x = kin[i];
          const unsigned d = ((unsigned)(x >> 32) >> shift) & 255u;
          kout[base[d] + wcnt[warp * 256 + d] + loc[k]] = x;
        }
      }
      __syncthreads();
      unsigned long long* sw = kin; kin = kout; kout = sw;
    }
    keys = kin;
  } else
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < (n >> 1); i += nthr) {
        const int lo = 2 * i - (i & (j - 1));
        const int hi = lo + j;
        const unsigned long long x = keys[lo], y = keys[hi];
        const bool up = (lo & k) == 0;
        if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
      }
      __syncthreads();
    }
  }

  // 3. number of real keys (pads sort last; the radix form sorted the first L entries only)
  if (tid == 0) {
    int lo = 0, hi = radix ? (int)L : n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] == pad) hi = mid; else lo = mid + 1; }
    s_nvalid = lo;
  }
  __syncthreads();
  const int nvalid = s_nvalid;

  // 4. head flags -> exclusive scan -> seg_start[]
  const int chunk = (nvalid + nthr - 1) / nthr;
  const int c0 = min(tid * chunk, nvalid), c1 = min(c0 + chunk, nvalid);
  int heads = 0;
  for (int i = c0; i < c1; ++i)
    heads += (i == 0) || ((unsigned)(keys[i] >> 32) != (unsigned)(keys[i - 1] >> 32));
  int incl = heads;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += v;
  }
  if ((tid & 31) == 31) s_warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    const int nw = (nthr + 31) >> 5;
    int v = tid < nw ? s_warp_tot[tid] : 0;
    int w = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, w, d);
      if (tid >= d) w += u;
    }
    s_warp_tot[tid] = w - v;                   // exclusive prefix of warp totals
    if (tid == 31) s_unique = w;
  }
  __syncthreads();
  int pos = s_warp_tot[tid >> 5] + incl - heads;
  for (int i = c0; i < c1; ++i)
    if ((i == 0) || ((unsigned)(keys[i] >> 32) != (unsigned)(keys[i - 1] >> 32))) seg_start[pos++] = i;
  U = s_unique;
  if (tid == 0) {
    seg_start[U] = nvalid;
    if (U > capacity) bad |= DQRM_STATUS_CAPACITY;
    uniq_count[t] = (int)(U > capacity ? capacity : U);
  }
  }                                                                      // (cr == 0)
  // the other CTAs of the cluster read CTA 0's sorted keys / segment starts / counters through distributed shared memory
  int* q_nlong = &s_nlong;
  int* q_long_j = s_long_j;
  unsigned* q_max = &s_max;
  if (csize > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (cr != 0 && radix) {                                                // which buffer CTA 0's sort ended in
      int key_bits = 1;
      while ((1ll << key_bits) <= nrows) ++key_bits;
      if (((key_bits + 7) / 8) & 1) keys += n;
    }
    keys = cluster.map_shared_rank(keys, 0);
    seg_start = cluster.map_shared_rank(seg_start, 0);
    q_nlong = cluster.map_shared_rank(&s_nlong, 0);
    q_long_j = cluster.map_shared_rank(s_long_j, 0);
    q_max = cluster.map_shared_rank(&s_max, 0);
    U = *cluster.map_shared_rank(&s_unique, 0);
  } else {
    __syncthreads();
  }
  if (U > capacity) U = (int)capacity;

  // 5. segmented fold of dy over each unique row
  const bool quant = fwd_scale != nullptr;
  const float s = quant ? fwd_scale[t] : 1.0f;
  const int lane = tid % group;
  const float* dbase = dout + (long long)t * dts;
  unsigned m = 0u;
  constexpr int kBwdPrefetch = kBwdPrefetchRows / COLS;
  // left fold of the lookups keys[p0..p1) into acc (p1 > p0)
  auto fold = [&](int p0, int p1, float4 (&acc)[COLS]) {
    for (int p = p0; p < p1; p += kBwdPrefetch) {
      float4 v[kBwdPrefetch][COLS];
#pragma unroll
      for (int u = 0; u < kBwdPrefetch; ++u) {
        const bool live = p + u < p1;
        const long long bag = live ? (long long)(unsigned)keys[p + u] : 0;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          const int col = lane + c * group;
          v[u][c] = (live && col < dim4)
                        ? __ldg(reinterpret_cast<const float4*>(dbase + bag * dbs) + col)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < kBwdPrefetch; ++u) {
        if (p + u >= p1) break;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          const float4 d = make_float4(ste_dy(v[u][c].x, s, quant), ste_dy(v[u][c].y, s, quant),
                                       ste_dy(v[u][c].z, s, quant), ste_dy(v[u][c].w, s, quant));
          if (p + u == p0) {
            acc[c] = d;
          } else {
            acc[c].x = __fadd_rn(acc[c].x, d.x); acc[c].y = __fadd_rn(acc[c].y, d.y);
            acc[c].z = __fadd_rn(acc[c].z, d.z); acc[c].w = __fadd_rn(acc[c].w, d.w);
          }
        }
      }
    }
  };
  auto emit = [&](int j, int p0, const float4 (&acc)[COLS]) {
    if (lane == 0) uniq_rows[(long long)t * capacity + j] = (int)(keys[p0] >> 32);
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int col = lane + c * group;
      if (col >= dim4) continue;
      reinterpret_cast<float4*>(grad_sums + ((long long)t * capacity + j) * dim4 * 4)[col] = acc[c];
      m = max(m, abs_bits4(acc[c]));
    }
  };
  // 5a. rows with <= kFoldBlock duplicates: left fold each; longer rows are queued.  A lane group folds
  // kRowsInFlight rows at once (step s of all of them together): several independent gathers in flight per group
  // instead of one dependent load per unique row.
  constexpr int kRowsInFlight = 4, kShortRow = 16;   // (16: measured on the sort kernel, 2.5x on rows of 5..16 duplicates)
  const int gstride = (nthr / group) * csize;                           // lane groups of the whole cluster
  const int gfirst = cr * (nthr / group) + tid / group;
  for (int jb = gfirst; jb < U; jb += gstride * kRowsInFlight) {
    int p0[kRowsInFlight], len[kRowsInFlight], maxlen = 0;
#pragma unroll
    for (int r = 0; r < kRowsInFlight; ++r) {
      const int j = jb + r * gstride;
      p0[r] = 0; len[r] = 0;
      if (j < U) {
        p0[r] = seg_start[j];
        len[r] = seg_start[j + 1] - p0[r];
        if (len[r] > kFoldBlock && partials != nullptr) {
          if (lane == 0) q_long_j[atomicAdd(q_nlong, 1)] = j;              // queue order is irrelevant to the results
          len[r] = 0;
        }
      }
      if (len[r] <= kShortRow) maxlen = max(maxlen, len[r]);
    }
    float4 acc[kRowsInFlight][COLS];
    for (int st = 0; st < maxlen; ++st) {
      float4 v[kRowsInFlight][1][COLS];   // (64 registers per thread at 1024 threads)
#pragma unroll
      for (int r = 0; r < kRowsInFlight; ++r)
#pragma unroll
        for (int u = 0; u < 1; ++u) {
          const bool live = st + u < len[r] && len[r] <= kShortRow;
          const long long bag = live ? (long long)(unsigned)keys[p0[r] + st + u] : 0;
#pragma unroll
          for (int c = 0; c < COLS; ++c) {
            const int col = lane + c * group;
            v[r][u][c] = (live && col < dim4) ? __ldg(reinterpret_cast<const float4*>(dbase + bag * dbs) + col)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
      for (int r = 0; r < kRowsInFlight; ++r)
#pragma unroll
        for (int u = 0; u < 1; ++u) {
          if (st + u >= len[r] || len[r] > kShortRow) continue;
#pragma unroll
          for (int c = 0; c < COLS; ++c) {
            const float4 d = make_float4(ste_dy(v[r][u][c].x, s, quant), ste_dy(v[r][u][c].y, s, quant),
                                         ste_dy(v[r][u][c].z, s, quant), ste_dy(v[r][u][c].w, s, quant));
            if (st + u == 0) acc[r][c] = d;
            else {
              acc[r][c].x = __fadd_rn(acc[r][c].x, d.x); acc[r][c].y = __fadd_rn(acc[r][c].y, d.y);
              acc[r][c].z = __fadd_rn(acc[r][c].z, d.z); acc[r][c].w = __fadd_rn(acc[r][c].w, d.w);
            }
          }
        }
    }
#pragma unroll
    for (int r = 0; r < kRowsInFlight; ++r)
      if (len[r] > 0 && len[r] <= kShortRow) emit(jb + r * gstride, p0[r], acc[r]);
  }
  // rows with 5..kFoldBlock duplicates: one row at a time, 8 gathers in flight (a separate loop keeps the two
  // register working sets apart: 64 registers per thread at 1024 threads)
  for (int j = gfirst; j < U; j += gstride) {
    const int p0 = seg_start[j], p1 = seg_start[j + 1];
    if (p1 - p0 <= kShortRow || (p1 - p0 > kFoldBlock && partials != nullptr)) continue;
    float4 acc[COLS];
    fold(p0, p1, acc);
    emit(j, p0, acc);
  }
  if (csize > 1) cg::this_cluster().sync(); else __syncthreads();       // the queue of long rows is complete
  const int nlong = cr == 0 ? s_nlong : 0;                               // (long rows: CTA 0, from its own shared memory)
  if (nlong > 0) {                                                       // CTA-uniform
    // 5b. work items = blocks of kFoldBlock consecutive lookups of the queued rows
    if (tid == 0) {
      int it = 0;
      for (int i = 0; i < nlong; ++i) {
        s_long_start[i] = it;
        const int j = s_long_j[i];
        it += (seg_start[j + 1] - seg_start[j] + kFoldBlock - 1) / kFoldBlock;
      }
      s_long_start[nlong] = it;
    }
    __syncthreads();
    const int items = s_long_start[nlong];
    float* part_t = partials + (long long)t * partial_items * dim4 * 4;
    for (int it = tid / group; it < items; it += nthr / group) {
      int lo = 0, hi = nlong;                                            // last i with s_long_start[i] <= it
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_long_start[mid] <= it) lo = mid; else hi = mid; }
      const int j = s_long_j[lo], blk = it - s_long_start[lo];
      const int p0 = seg_start[j] + blk * kFoldBlock, p1 = min(seg_start[j + 1], p0 + kFoldBlock);
      float4 acc[COLS];
      fold(p0, p1, acc);
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const int col = lane + c * group;
        if (col < dim4) reinterpret_cast<float4*>(part_t + (long long)it * dim4 * 4)[col] = acc[c];
      }
    }
    __threadfence_block();
    __syncthreads();
    // 5c. block sums of a row folded left to right
    for (int i = tid / group; i < nlong; i += nthr / group) {
      const int j = s_long_j[i], it0 = s_long_start[i], it1 = s_long_start[i + 1];
      float4 acc[COLS];
      for (int it = it0; it < it1; ++it) {
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          const int col = lane + c * group;
          if (col >= dim4) continue;
          const float4 d = reinterpret_cast<const float4*>(part_t + (long long)it * dim4 * 4)[col];
          if (it == it0) {
            acc[c] = d;
          } else {
            acc[c].x = __fadd_rn(acc[c].x, d.x); acc[c].y = __fadd_rn(acc[c].y, d.y);
            acc[c].z = __fadd_rn(acc[c].z, d.z); acc[c].w = __fadd_rn(acc[c].w, d.w);
          }
        }
      }
      emit(j, seg_start[j], acc);
    }
  }

  // 6. per-table gradient scale: warp maxima into CTA 0's slot (CTA 0 stays resident until every reader is done)
  m = warp_max_u32(m);
  if ((tid & 31) == 0 && m) atomicMax(q_max, m);
  if (csize > 1) cg::this_cluster().sync(); else __syncthreads();
  if (cr == 0 && tid == 0 && grad_scale_local) grad_scale_local[t] = scale_of(__uint_as_float(s_max), grad_bits);
  if (bad) atomicOr(status, bad);
}

__global__ void grad_absmax_scale_kernel(int dim, const float* __restrict__ grad_sums, const int* __restrict__ uniq_count,
                                         long long capacity, int bits, float* __restrict__ scale_local) {
  __shared__ unsigned s_max;
  const int t = blockIdx.x;
  const long long n = (long long)uniq_count[t] * dim;
  const float* g = grad_sums + (long long)t * capacity * dim;
  unsigned m = 0u;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) m = max(m, abs_bits(g[i]));
  const unsigned bm = block_max_u32(m, &s_max);
  if (threadIdx.x == 0) scale_local[t] = scale_of(__uint_as_float(bm), bits);
}

struct MomPtrs { float* p[DQRM_MAX_TABLES]; };

// W[row] += (-lr) * (sum * inv_world)                       sgd_quantized_gradients_parallel_comm.py:626
// or row-wise sparse Adagrad (USE_MOM):                      optim/rwsadagrad.py:97-113
//   m[row] += mean_d(g^2);  W[row] += (-lr) * (g / (sqrt(m[row]) + eps))
template <int COLS, bool USE_MOM>
__global__ void __launch_bounds__(256)
sgd_rows_kernel(const __grid_constant__ TableSet ts, const __grid_constant__ MomPtrs mom, int dim4, int group,
                const int* __restrict__ uniq_rows, const int* __restrict__ uniq_count,
                const float* __restrict__ grad_sums, long long capacity, float neg_lr_arg, const float* __restrict__ lr_dev,
                float inv_world, float eps) {
  const float neg_lr = lr_dev ? -(*lr_dev) : neg_lr_arg;
  const int t = blockIdx.y;
  const int U = uniq_count[t];
  const int lane = threadIdx.x % group;
  const int gpb = blockDim.x / group;
  for (int jb = blockIdx.x * gpb; jb < U; jb += gridDim.x * gpb) {     // block-uniform trip count (shuffles below)
    const int j = min(jb + (int)threadIdx.x / group, U - 1);
    const bool live = jb + (int)threadIdx.x / group < U;
    const long long row = uniq_rows[(long long)t * capacity + j];
    const float4* g4 = reinterpret_cast<const float4*>(grad_sums + ((long long)t * capacity + j) * dim4 * 4);
    float4* w4 = reinterpret_cast<float4*>(ts.w[t] + row * dim4 * 4);
    float4 g[COLS];
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int col = lane + c * group;
      g[c] = col < dim4 ? g4[col] : make_float4(0.f, 0.f, 0.f, 0.f);
      g[c].x = __fmul_rn(g[c].x, inv_world); g[c].y = __fmul_rn(g[c].y, inv_world);
      g[c].z = __fmul_rn(g[c].z, inv_world); g[c].w = __fmul_rn(g[c].w, inv_world);
      if (USE_MOM) sq += g[c].x * g[c].x + g[c].y * g[c].y + g[c].z * g[c].z + g[c].w * g[c].w;
    }
    float std = 1.0f;
    if (USE_MOM) {
      for (int d = group >> 1; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
      float* mrow = mom.p[t] + row;
      float mval = 0.f;
      if (lane == 0 && live) { mval = *mrow + sq / (float)(dim4 * 4); *mrow = mval; }
      mval = __shfl_sync(0xffffffffu, mval, (threadIdx.x & 31) - lane);
      std = sqrtf(mval) + eps;
    }
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int col = lane + c * group;
      if (col >= dim4 || !live) continue;
      float4 w = w4[col];
      float4 u = g[c];
      if (USE_MOM) { u.x = __fdiv_rn(u.x, std); u.y = __fdiv_rn(u.y, std); u.z = __fdiv_rn(u.z, std); u.w = __fdiv_rn(u.w, std); }
      w.x = __fadd_rn(w.x, __fmul_rn(neg_lr, u.x)); w.y = __fadd_rn(w.y, __fmul_rn(neg_lr, u.y));
      w.z = __fadd_rn(w.z, __fmul_rn(neg_lr, u.z)); w.w = __fadd_rn(w.w, __fmul_rn(neg_lr, u.w));
      w4[col] = w;
    }
  }
}

}  // namespace dqrm

using namespace dqrm;

namespace dqrm { size_t bwd_large_workspace_bytes(int64_t lookups, int dim); }   // embbag_bwd_large.cu

// blocks of a table's long rows: every one has > DQRM_FOLD_BLOCK lookups, so at most L/64 full + L/65 partial blocks
static long long partial_items_per_table(int64_t lookups) { return lookups / DQRM_FOLD_BLOCK + lookups / (DQRM_FOLD_BLOCK + 1) + 2; }

extern "C" size_t dqrm_bwd_workspace_bytes(int num_tables, int64_t max_lookups_per_table, int dim) {
  // either path may be chosen per call (pick_sort_path below): room for both
  const size_t sort_bytes = dqrm::bwd_large_workspace_bytes(max_lookups_per_table, dim);
  if (max_lookups_per_table > DQRM_BWD_CTA_MAX_LOOKUPS) return sort_bytes;
  const size_t cta_bytes =                                        // block sums of the long rows, [T][items][dim] fp32
      (size_t)num_tables * (size_t)partial_items_per_table(max_lookups_per_table) * (size_t)dim * sizeof(float);
  return cta_bytes > sort_bytes ? cta_bytes : sort_bytes;
}

// One CTA per table (all tables in ONE launch, time ~ the longest table) against one cooperative whole-chip sort
// launch per table (time ~ a fixed ~60 us of grid barriers each, then ~0.15 us per 1k lookups).  Measured on B200
// (profiles/r02_sweep.jsonl): the CTA path costs ~(12 + dim/4) us per 1k lookups of the longest table.  26 tables of
// 8192 lookups stay on the CTA path (0.2 ms vs 26 x 60 us); ONE table of >= 4k lookups goes to the sort kernel.
// DQRM_BWD_PATH=cta|sort pins the choice (tests).  Both paths produce the same bits (blocked left fold, same order).
static bool pick_sort_path(int num_tables, const long long* lookups, long long lmax, int dim) {
  if (lmax > DQRM_BWD_CTA_MAX_LOOKUPS) return true;
  const char* e = getenv("DQRM_BWD_PATH");                       // read per call: tests switch it
  if (e && !strcmp(e, "cta")) return false;
  if (e && !strcmp(e, "sort")) return true;
  const double cta_us = (double)lmax * (12.0 + 0.25 * dim) / 1024.0;
  double sort_us = 0.0;
  for (int k = 0; k < num_tables; ++k) sort_us += lookups[k] ? 60.0 + (double)lookups[k] * (0.08 + 0.001 * dim) / 1024.0 : 2.0;
  return sort_us < cta_us;
}

namespace dqrm {
int embbag_bwd_large(int t, long long rows, long long idx_begin, long long idx_end, int dim,
                     const int64_t* indices, const int64_t* offsets, int64_t bags,
                     const float* dout, int64_t dts, int64_t dbs, const float* fwd_scale,
                     int64_t capacity, int32_t* uniq_rows, int32_t* uniq_count, float* grad_sums,
                     int grad_bits, float* grad_scale_local, int32_t* status,
                     void* workspace, size_t workspace_bytes, cudaStream_t st, const RowUpdate* upd);   // embbag_bwd_large.cu
}

struct FusedUpdate {                   // dqrm_embbag_bwd_sgd: the row update that follows the de-duplication
  float* const* weight;
  float* const* momentum;
  float lr;
  const float* lr_dev;
  float inv_world;
  float eps;
};

static int bwd_impl(int num_tables, const int64_t* rows, int dim,
                    const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin, int64_t bags,
                    const float* dout, int64_t dout_table_stride, int64_t dout_bag_stride,
                    const float* fwd_scale,
                    int64_t capacity, int32_t* uniq_rows, int32_t* uniq_count, float* grad_sums,
                    int grad_bits, float* grad_scale_local,
                    int32_t* status, void* workspace, size_t workspace_bytes, void* stream, const FusedUpdate* fu) {
  DQRM_REQUIRE(rows && indices && offsets && idx_begin && dout && uniq_rows && uniq_count && grad_sums && status,
               -EINVAL, "embbag_bwd: null argument");
  DQRM_REQUIRE(num_tables >= 1 && num_tables <= DQRM_MAX_TABLES, -E2BIG, "embbag_bwd: num_tables=%d", num_tables);
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "embbag_bwd: dim=%d must be a multiple of 4 in [4,512]", dim);
  DQRM_REQUIRE(bags >= 1 && bags < (1ll << 31), -EINVAL, "embbag_bwd: bags=%lld", (long long)bags);
  DQRM_REQUIRE(!grad_scale_local || (grad_bits >= 2 && grad_bits <= 16), -EINVAL, "embbag_bwd: grad_bits=%d", grad_bits);
  DQRM_REQUIRE(dout_bag_stride % 4 == 0 && dout_table_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(dout) & 15u) == 0,
               -EINVAL, "embbag_bwd: dout must be 16-byte aligned with strides multiple of 4");
  DQRM_REQUIRE((reinterpret_cast<uintptr_t>(grad_sums) & 15u) == 0, -EINVAL, "embbag_bwd: grad_sums not 16-byte aligned");
  BwdArgs a;
  long long lmax = 0, lookups[DQRM_MAX_TABLES];
  for (int k = 0; k < num_tables; ++k) {
    DQRM_REQUIRE(rows[k] >= 1 && rows[k] < (1ll << 31), -EINVAL, "embbag_bwd: rows[%d]=%lld", k, (long long)rows[k]);
    DQRM_REQUIRE(idx_begin[k + 1] >= idx_begin[k], -EINVAL, "embbag_bwd: idx_begin not monotone at %d", k);
    a.rows[k] = rows[k];
    a.idx_begin[k] = idx_begin[k];
    const long long L = idx_begin[k + 1] - idx_begin[k];
    DQRM_REQUIRE(L <= capacity, -EINVAL, "embbag_bwd: table %d has %lld lookups > capacity %lld", k, L, (long long)capacity);
    if (L > lmax) lmax = L;
    lookups[k] = L;
  }
  a.idx_begin[num_tables] = idx_begin[num_tables];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const RowLanes rl = row_lanes(dim);

  if (pick_sort_path(num_tables, lookups, lmax, dim) &&
      (lmax > DQRM_BWD_CTA_MAX_LOOKUPS || workspace_bytes >= dqrm::bwd_large_workspace_bytes(lmax, dim))) {
    // one persistent whole-chip radix-sort + fold launch per table
    for (int k = 0; k < num_tables; ++k) {
      RowUpdate upd{};
      if (fu) upd = RowUpdate{fu->weight[k], fu->momentum ? fu->momentum[k] : nullptr, -fu->lr, fu->lr_dev, fu->inv_world, fu->eps};
      int rc = embbag_bwd_large(k, rows[k], idx_begin[k], idx_begin[k + 1], dim, indices, offsets, bags, dout,
                                dout_table_stride, dout_bag_stride, fwd_scale, capacity, uniq_rows, uniq_count,
                                grad_sums, grad_bits, grad_scale_local, status, workspace, workspace_bytes, st,
                                fu ? &upd : nullptr);
      if (rc) return rc;
    }
    return 0;
  }

  // long-row block sums live in the caller's workspace; without one (NULL / too small) every row is a plain left
  // fold, which is the same arithmetic for rows of <= DQRM_FOLD_BLOCK lookups
  const long long items = partial_items_per_table(capacity);
  float* partials = (workspace && workspace_bytes >= (size_t)num_tables * items * dim * sizeof(float) &&
                     (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0) ? static_cast<float*>(workspace) : nullptr;
  DQRM_REQUIRE(partials || capacity <= DQRM_FOLD_BLOCK, -ENOMEM,
               "embbag_bwd: workspace of %zu B needed for the blocked fold (dqrm_bwd_workspace_bytes), got %zu B",
               (size_t)num_tables * items * dim * sizeof(float), workspace_bytes);
  int n = 2;
  while (n < lmax) n <<= 1;
  int threads = n / 2;
  if (threads < 128) threads = 128;
  if (threads > 1024) threads = 1024;
  // stable radix sort in shared memory from 1024 keys while a second key buffer fits (<= 8192 keys); DQRM_BWD_CTA_SORT=bitonic pins the network
  int radix = (n >= 1024 && n <= 8192) ? 1 : 0;
  if (const char* e = getenv("DQRM_BWD_CTA_SORT")) { if (!strcmp(e, "bitonic")) radix = 0; }
  const size_t smem = radix ? (size_t)n * 2 * sizeof(unsigned long long) + ((size_t)n + 4) * sizeof(int) +
                                  ((size_t)(threads / 32) * 256 + 512) * sizeof(unsigned)
                            : (size_t)n * sizeof(unsigned long long) + ((size_t)n + 4) * sizeof(int);
  // CTAs per table.  A thread-block cluster per table (DQRM_BWD_CLUSTER=2|4|8) folds 3x faster in isolation, but inside
  // a training step this kernel runs on a side stream BESIDE the bottom-MLP backward: measured at Terabyte shape, batch
  // 8192, 26 clusters of 4 x 1024 threads slow that critical chain more than they save (step 8.12 -> 8.24 ms), so the
  // default stays one CTA per table.
  int csize = 1;
  if (const char* e = getenv("DQRM_BWD_CLUSTER")) { const int v = atoi(e); if (v >= 1 && v <= 8) csize = v; }
#define DQRM_BWD(COLS)                                                                                         \
  do {                                                                                                         \
    auto kern = embbag_bwd_cta_kernel<COLS>;                                                                   \
    if (smem > 48 * 1024) {                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
      DQRM_REQUIRE(e == cudaSuccess, -EIO, "embbag_bwd: cannot opt in to %zu B shared memory: %s", smem,       \
                   cudaGetErrorString(e));                                                                     \
    }                                                                                                          \
    cudaLaunchConfig_t cfg = {};                                                                               \
    cfg.gridDim = dim3((unsigned)(num_tables * csize));                                                        \
    cfg.blockDim = dim3((unsigned)threads);                                                                    \
    cfg.dynamicSmemBytes = smem;                                                                               \
    cfg.stream = st;                                                                                           \
    cudaLaunchAttribute attr[1];                                                                               \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                          \
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;    \
    cfg.attrs = attr;                                                                                          \
    cfg.numAttrs = csize > 1 ? 1 : 0;                                                                          \
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a, dim / 4, rl.group, reinterpret_cast<const long long*>(indices), \
                                        reinterpret_cast<const long long*>(offsets), (long long)bags, dout,    \
                                        (long long)dout_table_stride, (long long)dout_bag_stride, fwd_scale,   \
                                        (long long)capacity, uniq_rows, uniq_count, grad_sums, grad_bits,      \
                                        grad_scale_local, status, partials, (long long)items, csize, radix);   \
    DQRM_REQUIRE(le == cudaSuccess, -EIO, "embbag_bwd_cta_kernel: %s", cudaGetErrorString(le));                \
  } while (0)
  if (rl.cols == 1) DQRM_BWD(1);
  else if (rl.cols == 2) DQRM_BWD(2);
  else DQRM_BWD(4);
#undef DQRM_BWD
  DQRM_LAUNCH_CHECK("embbag_bwd_cta_kernel");
  if (fu)                                // few lookups per table: the single-CTA de-duplication, then the row update
    return dqrm_sgd_rows(num_tables, fu->weight, rows, dim, uniq_rows, uniq_count, grad_sums, capacity, fu->lr, fu->lr_dev,
                         fu->inv_world, fu->momentum, fu->eps, stream);
  return 0;
}

extern "C" int dqrm_embbag_bwd(int num_tables, const int64_t* rows, int dim,
                               const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin, int64_t bags,
                               const float* dout, int64_t dout_table_stride, int64_t dout_bag_stride,
                               const float* fwd_scale,
                               int64_t capacity, int32_t* uniq_rows, int32_t* uniq_count, float* grad_sums,
                               int grad_bits, float* grad_scale_local,
                               int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_impl(num_tables, rows, dim, indices, offsets, idx_begin, bags, dout, dout_table_stride, dout_bag_stride,
                  fwd_scale, capacity, uniq_rows, uniq_count, grad_sums, grad_bits, grad_scale_local, status, workspace,
                  workspace_bytes, stream, nullptr);
}

extern "C" int dqrm_embbag_bwd_sgd(int num_tables, float* const* weight, const int64_t* rows, int dim,
                                   const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin, int64_t bags,
                                   const float* dout, int64_t dout_table_stride, int64_t dout_bag_stride,
                                   const float* fwd_scale,
                                   int64_t capacity, int32_t* uniq_rows, int32_t* uniq_count, float* grad_sums,
                                   float lr, const float* lr_dev, float inv_world, float* const* momentum, float eps,
                                   int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  DQRM_REQUIRE(weight, -EINVAL, "embbag_bwd_sgd: null argument");
  for (int k = 0; k < num_tables && k < DQRM_MAX_TABLES; ++k) {
    DQRM_REQUIRE(weight[k] && (reinterpret_cast<uintptr_t>(weight[k]) & 15u) == 0, -EINVAL,
                 "embbag_bwd_sgd: table %d base pointer is null or not 16-byte aligned", k);
    DQRM_REQUIRE(!momentum || momentum[k], -EINVAL, "embbag_bwd_sgd: momentum[%d] is null", k);
  }
  const FusedUpdate fu{weight, momentum, lr, lr_dev, inv_world, eps};
  return bwd_impl(num_tables, rows, dim, indices, offsets, idx_begin, bags, dout, dout_table_stride, dout_bag_stride,
                  fwd_scale, capacity, uniq_rows, uniq_count, grad_sums, 8, nullptr, status, workspace,
                  workspace_bytes, stream, &fu);
}

extern "C" int dqrm_grad_absmax_scale(int num_tables, int dim, const float* grad_sums, const int32_t* uniq_count,
                                      int64_t capacity, int bits, float* scale_local, void* stream) {
  DQRM_REQUIRE(grad_sums && uniq_count && scale_local, -EINVAL, "grad_absmax_scale: null argument");
  DQRM_REQUIRE(num_tables >= 1 && bits >= 2 && bits <= 16, -EINVAL, "grad_absmax_scale: bad argument");
  grad_absmax_scale_kernel<<<num_tables, 256, 0, static_cast<cudaStream_t>(stream)>>>(dim, grad_sums, uniq_count,
                                                                                      capacity, bits, scale_local);
  DQRM_LAUNCH_CHECK("grad_absmax_scale_kernel");
  return 0;
}

extern "C" int dqrm_sgd_rows(int num_tables, float* const* weight, const int64_t* rows, int dim,
                             const int32_t* uniq_rows, const int32_t* uniq_count, const float* grad_sums,
                             int64_t capacity, float lr, const float* lr_dev, float inv_world, float* const* momentum, float eps,
                             void* stream) {
  DQRM_REQUIRE(weight && rows && uniq_rows && uniq_count && grad_sums, -EINVAL, "sgd_rows: null argument");
  DQRM_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 512, -EINVAL, "sgd_rows: dim=%d", dim);
  TableSet ts;
  if (int rc = fill_tables(ts, num_tables, weight, rows, nullptr)) return rc;
  const RowLanes rl = row_lanes(dim);
  MomPtrs mom;
  for (int k = 0; k < DQRM_MAX_TABLES; ++k) mom.p[k] = (momentum && k < num_tables) ? momentum[k] : nullptr;
  long long blocks = ceil_div(capacity, 256 / rl.group);
  if (blocks > 4 * kSMs) blocks = 4 * kSMs;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, num_tables);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float neg_lr = -lr;
#define DQRM_SGD(COLS)                                                                                              \
  do {                                                                                                              \
    if (momentum) sgd_rows_kernel<COLS, true><<<grid, 256, 0, st>>>(ts, mom, dim / 4, rl.group, uniq_rows, uniq_count, \
                                                                    grad_sums, capacity, neg_lr, lr_dev, inv_world, eps); \
    else sgd_rows_kernel<COLS, false><<<grid, 256, 0, st>>>(ts, mom, dim / 4, rl.group, uniq_rows, uniq_count,      \
                                                            grad_sums, capacity, neg_lr, lr_dev, inv_world, eps);   \
  } while (0)
  if (rl.cols == 1) DQRM_SGD(1);
  else if (rl.cols == 2) DQRM_SGD(2);
  else DQRM_SGD(4);
#undef DQRM_SGD
  DQRM_LAUNCH_CHECK("sgd_rows_kernel");
  return 0;
}
