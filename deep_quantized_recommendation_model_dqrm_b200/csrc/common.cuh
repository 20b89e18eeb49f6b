// Shared helpers for libdqrm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dqrm_b200.h"

#ifndef __CUDA_ARCH_LIST__
#endif

namespace dqrm {

constexpr int kSMs = 148;                 // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);     // api.cu

#define DQRM_REQUIRE(cond, code, ...)                       \
  do {                                                      \
    if (!(cond)) { ::dqrm::set_error(__VA_ARGS__); return (code); } \
  } while (0)

#define DQRM_LAUNCH_CHECK(name)                                                  \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      ::dqrm::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return -EIO;                                                               \
    }                                                                            \
  } while (0)

// ---- bit-exact fp32 primitives (never contracted into FMA) -------------------
__device__ __forceinline__ float qmax_of(int bits) { return (float)((1 << (bits - 1)) - 1); }

// s = max(absmax, 1e-8) / n          quant_utils.py:191-192
__device__ __forceinline__ float scale_of(float absmax, int bits) {
  return __fdiv_rn(fmaxf(absmax, 1e-8f), qmax_of(bits));
}
// clamp(rint(inv * x + 0), -n-1, n)  quant_utils.py:101,343   (integer-valued float)
__device__ __forceinline__ float quant_code(float x, float inv, float lo, float hi) {
  float q = rintf(__fadd_rn(__fmul_rn(inv, x), 0.0f));
  return fminf(fmaxf(q, lo), hi);
}
__device__ __forceinline__ unsigned abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__device__ __forceinline__ unsigned abs_bits4(float4 v) {
  return max(max(abs_bits(v.x), abs_bits(v.y)), max(abs_bits(v.z), abs_bits(v.w)));
}

// streaming 128-bit load: read-only path, do not allocate in L1
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ unsigned warp_max_u32(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }

// block-wide max of a u32; result valid in thread 0 (and in smem slot). THREADS multiple of 32.
__device__ __forceinline__ unsigned block_max_u32(unsigned v, unsigned* s_slot) {
  if (threadIdx.x == 0) *s_slot = 0u;
  __syncthreads();
  v = warp_max_u32(v);
  if ((threadIdx.x & 31) == 0 && v) atomicMax(s_slot, v);
  __syncthreads();
  return *s_slot;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Per-table descriptor passed BY VALUE as a kernel parameter (so calls need no
// device-side metadata and are graph-capturable).
struct TableSet {
  float* w[DQRM_MAX_TABLES];
  long long rows[DQRM_MAX_TABLES];
  long long idx_begin[DQRM_MAX_TABLES + 1];
  int num_tables;
};

inline int fill_tables(TableSet& ts, int num_tables, const float* const* weight, const int64_t* rows,
                       const int64_t* idx_begin) {
  if (num_tables < 1 || num_tables > DQRM_MAX_TABLES) {
    set_error("num_tables=%d outside [1,%d]", num_tables, DQRM_MAX_TABLES);
    return -E2BIG;
  }
  ts.num_tables = num_tables;
  for (int k = 0; k < num_tables; ++k) {
    ts.w[k] = weight ? const_cast<float*>(weight[k]) : nullptr;
    ts.rows[k] = rows ? rows[k] : 0;
    if (weight && (reinterpret_cast<uintptr_t>(weight[k]) & 15u)) {
      set_error("table %d base pointer is not 16-byte aligned", k);
      return -EINVAL;
    }
    if (rows && (rows[k] < 0 || rows[k] >= (1ll << 31))) {
      set_error("table %d has %lld rows (must be < 2^31)", k, (long long)rows[k]);
      return -EINVAL;
    }
  }
  if (idx_begin) {
    for (int k = 0; k <= num_tables; ++k) ts.idx_begin[k] = idx_begin[k];
    for (int k = 0; k < num_tables; ++k)
      if (idx_begin[k + 1] < idx_begin[k]) { set_error("idx_begin not monotone at %d", k); return -EINVAL; }
  }
  return 0;
}

// Fused in-place row update of the de-duplicating backward (dqrm_embbag_bwd_sgd): plain SGD, or row-wise sparse
// Adagrad when `mom` is set; the arithmetic of sgd_rows_kernel (embbag_bwd.cu).
struct RowUpdate {
  float* W;
  float* mom;
  float neg_lr;
  const float* lr_dev;
  float inv_world;
  float eps;
};

// lanes cooperating on one row: dim/4 float4 columns, rounded up to a power of two (<= 32);
// wider rows give each lane several columns.
struct RowLanes { int group; int cols; };
inline RowLanes row_lanes(int dim) {
  int d4 = dim / 4, g = 1;
  while (g < d4 && g < 32) g <<= 1;
  return RowLanes{g, (d4 + g - 1) / g};
}

}  // namespace dqrm
