// Shared device helpers for libtiger_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <utility>

#include "tiger_b200.h"

#define TIGER_FULL_MASK 0xffffffffu

static inline int tiger_launch_status() {
  return cudaGetLastError() == cudaSuccess ? TIGER_OK : TIGER_ECUDA;
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch.  The kernels of the per-batch chain run back to back on one stream, each a
// few microseconds long: with the launch attribute below a kernel's CTAs may be scheduled while its
// predecessor still runs (as soon as every CTA of the predecessor has called pdl_trigger or exited), set up
// what does not depend on it (barriers, tensor-memory allocation, weight tiles) and block in pdl_wait until
// the predecessor has completed and its writes are visible.  pdl_wait is a no-op for a normal launch.
// TIGER_NO_PDL=1 in the environment turns the attribute off (plain stream order).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline bool tiger_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TIGER_NO_PDL");
    on = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

template <typename... KP, typename... Args>
static inline int tiger_launch_chain(void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     dim3 cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n_attr = 0;
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = cluster.x;
    attr[n_attr].val.clusterDim.y = cluster.y;
    attr[n_attr].val.clusterDim.z = cluster.z;
    ++n_attr;
  }
  // The relaxed (programmatic) ordering is requested only while the stream is being captured into a CUDA graph: there
  // every dependency of the launch - including event joins from other branches - is an explicit edge.  Eager
  // launches keep plain stream order, so library calls compose with arbitrary caller code.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool capturing = cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive;
  static const bool eager_pdl = getenv("TIGER_EAGER_PDL") != nullptr;   // debug aid
  if ((capturing || eager_pdl) && tiger_pdl_enabled()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  if (cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...) != cudaSuccess) return TIGER_ECUDA;
  return tiger_launch_status();
}

// Dropout seeds under CUDA-graph replay.  The training step draws its masks from (seed, site, element) with a seed
// that advances by 101 per step on the host (train.py).  A captured step would replay the seed of the capture step, so
// the kernels add a device-side step counter (incremented inside the graph) when one is registered:
// tiger_train_seed_step(ptr).  Eager launches (no counter registered) use the host seed as given; the k-th replay
// draws exactly the masks the k-th eager step after the capture point would have drawn.
static __device__ const int32_t* tiger_seed_step_ptr = nullptr;        // one copy per translation unit
__device__ __forceinline__ uint32_t tiger_step_seed(uint32_t seed) {
  const int32_t* p = tiger_seed_step_ptr;
  return p == nullptr ? seed : ((seed + 101u * (uint32_t)(*p)) & 0x7fffffffu);
}
static inline int tiger_seed_step_set_here(const int32_t* p) {
  return cudaMemcpyToSymbol(tiger_seed_step_ptr, &p, sizeof(p)) == cudaSuccess ? TIGER_OK : TIGER_ECUDA;
}
int tiger_seed_step_set_train_seq(const int32_t* p);       // csrc/train_seq.cu
int tiger_seed_step_set_restart_seq(const int32_t* p);     // csrc/restart_seq.cu

// Counter-based dropout mask of the seq restarter's training step: element `idx` of mask stream `stream` (3: attention
// probabilities, 4: merger hidden layer) under `seed` is kept with probability 1 - p.
__device__ __forceinline__ uint32_t seq_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool seq_keep(uint32_t seed, uint32_t stream, uint32_t idx, float p) {
  if (p <= 0.f) return true;
  const uint32_t h = seq_mix32(idx ^ seq_mix32(seed + 0x9E3779B9u * (stream + 1u)));
  return (float)(h >> 8) * (1.0f / 16777216.0f) >= p;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id_in_block() { return threadIdx.x >> 5; }

// cos(t*w + b) with the product and the sum rounded separately, exactly like the reference's CPU
// TimeEncode (time_encoding.py:26: `ts * basis_freq + phase`, no FMA).  Arguments reach 1e6 rad (seconds
// since the last update times a basis frequency of 1): __cosf is useless there and cosf takes its
// Payne-Hanek slow path (local memory, hundreds of instructions, divergent within a warp because only
// the highest frequencies need it).  Instead the fp32 argument is reduced to [-pi, pi] in double
// (two-constant Cody-Waite, error < 1e-9 rad for |x| < 1e8) and cosf runs on its fast path.
__device__ __forceinline__ float cos_reduced(float x) {
  const double xd = (double)x;
  const double n = rint(xd * 0.15915494309189535);               // x / (2 pi)
  double r = fma(-n, 6.283185307179586, xd);                     // 2 pi (head)
  r = fma(-n, 2.4492935982947064e-16, r);                        // 2 pi (tail)
  return cosf((float)r);
}
__device__ __forceinline__ float time_enc(float t, float w, float b) {
  return cos_reduced(__fadd_rn(__fmul_rn(t, w), b));
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

// Warp-cooperative copy of NR rows at once (dst[r] == NULL skips row r): all loads of the NR rows are
// issued before the first store, so a warp keeps up to 2 * NR 16-byte loads per lane in flight - the
// memory-level parallelism a random-row gather / scatter needs to approach the HBM rate (a 172-float row
// is 43 vectors: one warp-wide load plus an 11-lane tail).
#define ROWS_PER_WARP 4
template <int NR>
__device__ __forceinline__ void warp_copy_rows(float* const (&dst)[NR], const float* const (&src)[NR], int width,
                                               int lane) {
  bool vec = (width & 3) == 0;
#pragma unroll
  for (int r = 0; r < NR; ++r)
    if (dst[r] != nullptr) vec = vec && ((((uintptr_t)dst[r] | (uintptr_t)src[r]) & 15) == 0);
  if (vec) {
    const int w4 = width >> 2;
    for (int base = 0; base < w4; base += 64) {
      const int c0 = base + lane, c1 = base + 32 + lane;
      float4 a[NR], b[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (dst[r] != nullptr && c0 < w4) a[r] = reinterpret_cast<const float4*>(src[r])[c0];
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (dst[r] != nullptr && c1 < w4) b[r] = reinterpret_cast<const float4*>(src[r])[c1];
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (dst[r] != nullptr && c0 < w4) reinterpret_cast<float4*>(dst[r])[c0] = a[r];
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (dst[r] != nullptr && c1 < w4) reinterpret_cast<float4*>(dst[r])[c1] = b[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r)
      if (dst[r] != nullptr)
        for (int i = lane; i < width; i += 32) dst[r][i] = src[r][i];
  }
}

// Lane l of the warp holds the (dst, src) pair of row l (l < rows; dst NULL = skip): the rows are copied four at
// a time.  A kernel resolves the index chain of up to 32 rows once, in parallel across the lanes, and then
// streams rows back to back - with 4 rows per warp the chain's latency was most of a warp's life and the
// kernels idled at half the bandwidth their loads in flight could have carried.
__device__ __forceinline__ void warp_copy_lane_rows(float* my_dst, const float* my_src, int rows, int width, int lane) {
  for (int g = 0; g < rows; g += ROWS_PER_WARP) {
    float* dst[ROWS_PER_WARP];
    const float* src[ROWS_PER_WARP];
#pragma unroll
    for (int i = 0; i < ROWS_PER_WARP; ++i) {
      dst[i] = reinterpret_cast<float*>(__shfl_sync(TIGER_FULL_MASK, reinterpret_cast<unsigned long long>(my_dst), g + i));
      src[i] = reinterpret_cast<const float*>(
          __shfl_sync(TIGER_FULL_MASK, reinterpret_cast<unsigned long long>(my_src), g + i));
    }
    warp_copy_rows<ROWS_PER_WARP>(dst, src, width, lane);
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

// Same search with G lanes per query (G = 8 or 16: 32 / G independent queries per warp, which is what keeps
// enough dependent-load chains in flight when segments are short).  Every lane of the warp must call it
// (inactive groups pass lo == hi); lanes of one group pass the same arguments and get the same result.
template <int G>
__device__ __forceinline__ int64_t group_lower_bound(const double* __restrict__ ts, int64_t lo, int64_t hi, double t,
                                                     int lane) {
  const int l = lane & (G - 1), shift = lane & ~(G - 1);
  const uint32_t gmask = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
  while (__any_sync(TIGER_FULL_MASK, hi - lo > G)) {
    const bool big = hi - lo > G;
    const int64_t step = (hi - lo + G - 1) / G;
    const int64_t p = lo + (int64_t)(l + 1) * step - 1;
    const bool less = big && p < hi && (__ldg(ts + p) < t);
    const int c = __popc((__ballot_sync(TIGER_FULL_MASK, less) >> shift) & gmask);
    if (big) {
      const int64_t pc = lo + (int64_t)(c + 1) * step - 1;
      const int64_t new_lo = lo + (int64_t)c * step;
      if (c < G && pc < hi) hi = pc;
      lo = new_lo < hi ? new_lo : hi;
    }
  }
  const int64_t p = lo + l;
  const bool less = (p < hi) && (__ldg(ts + p) < t);
  return lo + __popc((__ballot_sync(TIGER_FULL_MASK, less) >> shift) & gmask);
}
