// Shared device helpers for libtiger_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tiger_b200.h"

#define TIGER_FULL_MASK 0xffffffffu

static inline int tiger_launch_status() {
  return cudaGetLastError() == cudaSuccess ? TIGER_OK : TIGER_ECUDA;
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id_in_block() { return threadIdx.x >> 5; }

// cos(t*w + b) with the product and the sum rounded separately, exactly like the
// reference's CPU TimeEncode (time_encoding.py:26: `ts * basis_freq + phase`, no FMA).
// cosf (not __cosf): arguments reach 1e6 rad, the fast intrinsic is useless there.
__device__ __forceinline__ float time_enc(float t, float w, float b) {
  return cosf(__fadd_rn(__fmul_rn(t, w), b));
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// order-preserving map float -> unsigned (for atomicMax on timestamps); 0 is never produced
// for finite inputs, so 0 can mean "empty slot".  -0.0 is canonicalised to +0.0 first.
__device__ __forceinline__ uint64_t orderable_f32(float x) {
  uint32_t u = __float_as_uint(x + 0.0f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (uint64_t)u;
}
__device__ __forceinline__ uint64_t orderable_f64(double x) {
  uint64_t u = (uint64_t)__double_as_longlong(x + 0.0);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TIGER_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(TIGER_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(TIGER_FULL_MASK, v, o));
  return v;
}

// Warp-cooperative copy of one row of `width` floats (global -> global), 16-byte vectors when
// both rows are 16-byte aligned, coalesced scalars otherwise.
__device__ __forceinline__ void warp_copy_row(float* __restrict__ dst, const float* __restrict__ src,
                                              int width, int lane) {
  if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0 && (width & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < (width >> 2); i += 32) d4[i] = s4[i];
  } else {
    for (int i = lane; i < width; i += 32) dst[i] = src[i];
  }
}

// dst = a + b (b may be NULL => copy), same vector/scalar policy.
__device__ __forceinline__ void warp_add_row(float* __restrict__ dst, const float* __restrict__ a,
                                             const float* __restrict__ b, int width, int lane) {
  if (b == nullptr) {
    warp_copy_row(dst, a, width, lane);
    return;
  }
  if ((((uintptr_t)dst | (uintptr_t)a | (uintptr_t)b) & 15) == 0 && (width & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < (width >> 2); i += 32) {
      float4 x = a4[i], y = b4[i];
      d4[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
  } else {
    for (int i = lane; i < width; i += 32) dst[i] = a[i] + b[i];
  }
}

// Warp-cooperative 32-ary lower bound on a sorted float64 segment: first index in [lo,hi) whose
// timestamp is >= t (np.searchsorted side='left', reference graph.py:51).  All 32 lanes call it
// with the same arguments and receive the same result.
__device__ __forceinline__ int64_t warp_lower_bound(const double* __restrict__ ts, int64_t lo, int64_t hi,
                                                    double t, int lane) {
  // invariant: ts[i] < t for i < lo ; ts[i] >= t for i >= hi
  while (hi - lo > 32) {
    const int64_t step = (hi - lo + 31) >> 5;
    const int64_t p = lo + (int64_t)(lane + 1) * step - 1;
    const bool valid = p < hi;
    const bool less = valid && (__ldg(ts + p) < t);
    const int c = __popc(__ballot_sync(TIGER_FULL_MASK, less));
    // lanes 0..c-1 are 'less' (timestamps are sorted); lane c, if in range, is '>= t'
    const int64_t pc = lo + (int64_t)(c + 1) * step - 1;
    const int64_t new_lo = lo + (int64_t)c * step;
    if (c < 32 && pc < hi) hi = pc;
    lo = new_lo < hi ? new_lo : hi;
  }
  const int64_t p = lo + lane;
  const bool less = (p < hi) && (__ldg(ts + p) < t);
  return lo + __popc(__ballot_sync(TIGER_FULL_MASK, less));
}

