// a1  Device construction of the per-node time-sorted CSR (reference: Graph.__init__ +
// data2adjlist, tiger/data/graph.py:11-36,226-241, Python loops: 3 s per 157 k events).
//
// The stream must be ordered by time (the host wrapper stable-sorts it first if it is not).
// Entry 2e is (owner=src[e], other=dst[e], flag 0), entry 2e+1 is (owner=dst[e], other=src[e],
// flag 1).  The reference's per-node stable sort by time is then exactly a STABLE sort of the
// entry indices by owner id: an LSD radix sort, 8 bits per pass, each pass = per-warp-tile digit
// histogram -> exclusive scan (digit-major) -> order-preserving scatter (warp match_any ranks).
// Deterministic: no result depends on atomics ordering.
#include "common.cuh"

#define CSR_TILE 2048       // entries per warp tile
#define CSR_WARPS 8

__device__ __forceinline__ uint32_t owner_of(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                             uint32_t entry) {
  return (uint32_t)((entry & 1u) ? dst[entry >> 1] : src[entry >> 1]);
}

__global__ void csr_iota_degree_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                       int64_t n_entries, uint32_t* __restrict__ perm, uint32_t* __restrict__ deg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  perm[i] = (uint32_t)i;
  atomicAdd(deg + owner_of(src, dst, (uint32_t)i), 1u);   // counts only: order independent
}

__global__ void __launch_bounds__(CSR_WARPS * 32)
csr_hist_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, const uint32_t* __restrict__ perm,
                int64_t n_entries, int shift, int64_t n_tiles, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[CSR_WARPS][256];
  const int lane = lane_id(), warp = warp_id_in_block();
  const int64_t tile = (int64_t)blockIdx.x * CSR_WARPS + warp;
  for (int b = lane; b < 256; b += 32) sh[warp][b] = 0;
  __syncwarp();
  if (tile < n_tiles) {
    const int64_t base = tile * CSR_TILE;
    for (int it = 0; it < CSR_TILE / 32; ++it) {
      const int64_t i = base + it * 32 + lane;
      const bool valid = i < n_entries;
      const uint32_t key = valid ? ((owner_of(src, dst, perm[i]) >> shift) & 255u) : (256u + lane);
      const uint32_t peers = __match_any_sync(TIGER_FULL_MASK, key);
      if (valid && lane == __ffs(peers) - 1) sh[warp][key] += __popc(peers);
      __syncwarp();
    }
    for (int b = lane; b < 256; b += 32) hist[(int64_t)b * n_tiles + tile] = sh[warp][b];
  }
}

// Exclusive scan of `n` uint32 counters into OutT (uint32 in place, or int64), three launches:
// per-chunk sums (one CTA per 8192 counters) -> scan of the chunk sums (one CTA) -> per-chunk scan + offset.
// (The histogram of one radix pass has 256 x tiles counters - 2 M for 8 M events - and a single-CTA scan
// of it was 60 % of the build.)
#define SCAN_THREADS 1024
#define SCAN_PER_THREAD 8
#define SCAN_CHUNK (SCAN_THREADS * SCAN_PER_THREAD)

__device__ __forceinline__ unsigned long long block_exscan_u64(unsigned long long v, unsigned long long* total,
                                                               unsigned long long* warp_sums) {
  const int lane = lane_id(), warp = warp_id_in_block();
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(TIGER_FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const unsigned long long w = warp_sums[lane];
    unsigned long long s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(TIGER_FULL_MASK, s, o);
      if (lane >= o) s += t;
    }
    warp_sums[lane] = s - w;
    if (lane == 31 && total != nullptr) *total = s;
  }
  __syncthreads();
  return inc - v + warp_sums[warp];
}

__global__ void __launch_bounds__(SCAN_THREADS) csr_scan_sums_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                     unsigned long long* __restrict__ chunk_sums) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long total;
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_PER_THREAD;
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER_THREAD; ++i)
    if (base + i < n) s += in[base + i];
  block_exscan_u64(s, &total, warp_sums);
  if (threadIdx.x == 0) chunk_sums[blockIdx.x] = total;
}

// exclusive scan of the chunk sums in place (one CTA, any number of chunks)
__global__ void __launch_bounds__(SCAN_THREADS) csr_scan_chunks_kernel(unsigned long long* chunk_sums, int64_t n_chunks) {
  __shared__ unsigned long long warp_sums[32];
  const int64_t per = (n_chunks + SCAN_THREADS - 1) / SCAN_THREADS;
  const int64_t i0 = (int64_t)threadIdx.x * per, i1 = (i0 + per < n_chunks) ? i0 + per : n_chunks;
  unsigned long long s = 0;
  for (int64_t i = i0; i < i1; ++i) s += chunk_sums[i];
  unsigned long long run = block_exscan_u64(s, nullptr, warp_sums);
  for (int64_t i = i0; i < i1; ++i) {
    const unsigned long long v = chunk_sums[i];
    chunk_sums[i] = run;
    run += v;
  }
}

template <typename OutT>
__global__ void __launch_bounds__(SCAN_THREADS) csr_scan_apply_kernel(const uint32_t* in, int64_t n,
                                                                      const unsigned long long* __restrict__ chunk_offsets,
                                                                      OutT* out) {
  __shared__ unsigned long long warp_sums[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_PER_THREAD;
  uint32_t v[SCAN_PER_THREAD];
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER_THREAD; ++i) {
    v[i] = base + i < n ? in[base + i] : 0u;
    s += v[i];
  }
  unsigned long long run = block_exscan_u64(s, nullptr, warp_sums) + chunk_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_PER_THREAD; ++i) {
    if (base + i < n) out[base + i] = (OutT)run;   // in may alias out: every element was read above
    run += v[i];
  }
}

template <typename OutT>
static void csr_exscan(const uint32_t* in, int64_t n, OutT* out, unsigned long long* chunk_sums, cudaStream_t st) {
  const int64_t n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  csr_scan_sums_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, st>>>(in, n, chunk_sums);
  csr_scan_chunks_kernel<<<1, SCAN_THREADS, 0, st>>>(chunk_sums, n_chunks);
  csr_scan_apply_kernel<OutT><<<(unsigned)n_chunks, SCAN_THREADS, 0, st>>>(in, n, chunk_sums, out);
}

__global__ void __launch_bounds__(CSR_WARPS * 32)
csr_scatter_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                   const uint32_t* __restrict__ perm_in, uint32_t* __restrict__ perm_out, int64_t n_entries,
                   int shift, int64_t n_tiles, const uint32_t* __restrict__ offsets) {
  __shared__ uint32_t sh[CSR_WARPS][256];
  const int lane = lane_id(), warp = warp_id_in_block();
  const int64_t tile = (int64_t)blockIdx.x * CSR_WARPS + warp;
  if (tile >= n_tiles) return;
  for (int b = lane; b < 256; b += 32) sh[warp][b] = offsets[(int64_t)b * n_tiles + tile];
  __syncwarp();
  const int64_t base = tile * CSR_TILE;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int it = 0; it < CSR_TILE / 32; ++it) {
    const int64_t i = base + it * 32 + lane;
    const bool valid = i < n_entries;
    const uint32_t entry = valid ? perm_in[i] : 0u;
    const uint32_t key = valid ? ((owner_of(src, dst, entry) >> shift) & 255u) : (256u + lane);
    const uint32_t peers = __match_any_sync(TIGER_FULL_MASK, key);
    uint32_t pos = 0;
    if (valid) pos = sh[warp][key] + __popc(peers & lt_mask);
    __syncwarp();
    if (valid && lane == __ffs(peers) - 1) sh[warp][key] += __popc(peers);
    __syncwarp();
    if (valid) perm_out[pos] = entry;
  }
}

__global__ void csr_emit_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                const double* __restrict__ ts, const int64_t* __restrict__ eid,
                                const uint32_t* __restrict__ perm, int64_t n_entries, int32_t* __restrict__ adj_nbr,
                                int32_t* __restrict__ adj_eid, double* __restrict__ adj_ts,
                                uint8_t* __restrict__ adj_flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  const uint32_t entry = perm[i];
  const int64_t e = entry >> 1;
  const uint32_t side = entry & 1u;
  adj_nbr[i] = (int32_t)(side ? src[e] : dst[e]);
  adj_eid[i] = (int32_t)eid[e];
  adj_ts[i] = ts[e];
  adj_flag[i] = (uint8_t)side;
}

static inline int64_t csr_tiles(int64_t n_entries) { return (n_entries + CSR_TILE - 1) / CSR_TILE; }

extern "C" int64_t tiger_csr_build_work_bytes(int64_t n_events, int64_t n_nodes) {
  const int64_t n_entries = 2 * n_events;
  const int64_t words = 2 * n_entries + 256 * csr_tiles(n_entries) + (n_nodes + 1);
  const int64_t scan_n = 256 * csr_tiles(n_entries) > n_nodes + 1 ? 256 * csr_tiles(n_entries) : n_nodes + 1;
  const int64_t chunk_words = 2 * ((scan_n + SCAN_CHUNK - 1) / SCAN_CHUNK + 2);   // uint64 chunk sums
  return ((words + 1 + chunk_words) * 4 + 15) / 16 * 16;
}

extern "C" int tiger_csr_build(const int64_t* src, const int64_t* dst, const double* ts, const int64_t* eid,
                               int64_t n_events, int64_t n_nodes, int64_t* indptr, int32_t* adj_nbr,
                               int32_t* adj_eid, double* adj_ts, uint8_t* adj_flag, void* work, void* stream) {
  if (n_events < 0 || n_nodes <= 0 || 2 * n_events >= 0xffffffffll || n_nodes > 0x7fffffffll) return TIGER_EINVAL;
  cudaStream_t st = as_stream(stream);
  const int64_t n_entries = 2 * n_events;
  const int64_t n_tiles = csr_tiles(n_entries);
  uint32_t* perm_a = reinterpret_cast<uint32_t*>(work);
  uint32_t* perm_b = perm_a + n_entries;
  uint32_t* hist = perm_b + n_entries;
  uint32_t* deg = hist + 256 * n_tiles;
  unsigned long long* chunk_sums =
      reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(deg + n_nodes + 1) + 7) & ~(uintptr_t)7);
  cudaMemsetAsync(deg, 0, (size_t)(n_nodes + 1) * sizeof(uint32_t), st);
  if (n_entries > 0)
    csr_iota_degree_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, st>>>(src, dst, n_entries, perm_a, deg);
  csr_exscan<int64_t>(deg, n_nodes + 1, indptr, chunk_sums, st);
  if (n_entries == 0) return tiger_launch_status();
  int bits = 0;
  while (((int64_t)1 << bits) < n_nodes) ++bits;
  const unsigned tgrid = (unsigned)((n_tiles + CSR_WARPS - 1) / CSR_WARPS);
  for (int shift = 0; shift < bits || shift == 0; shift += 8) {
    csr_hist_kernel<<<tgrid, CSR_WARPS * 32, 0, st>>>(src, dst, perm_a, n_entries, shift, n_tiles, hist);
    csr_exscan<uint32_t>(hist, 256 * n_tiles, hist, chunk_sums, st);
    csr_scatter_kernel<<<tgrid, CSR_WARPS * 32, 0, st>>>(src, dst, perm_a, perm_b, n_entries, shift, n_tiles, hist);
    uint32_t* t = perm_a;
    perm_a = perm_b;
    perm_b = t;
  }
  csr_emit_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, st>>>(src, dst, ts, eid, perm_a, n_entries, adj_nbr,
                                                                     adj_eid, adj_ts, adj_flag);
  return tiger_launch_status();
}
