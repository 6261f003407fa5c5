// Last-message aggregator selection: per-node argmax-by-timestamp with torch_scatter's CPU tie
// rule (lowest position among the maxima), bit-exact with select_latest_nids
// (tiger/model/utils.py:10-16), and anonymized_reindex (tiger/model/utils.py:19-27).
#include "common.cuh"

#define SELECT_SMALL_N 2048

__device__ __forceinline__ uint64_t load_key(const void* ts, int is_f64, int64_t i) {
  return is_f64 ? orderable_f64(reinterpret_cast<const double*>(ts)[i])
                : orderable_f32(reinterpret_cast<const float*>(ts)[i]);
}

// ------------------------------------------------------------------------------------------
// small n: all-pairs in shared memory, one warp per position (lanes split the comparison partners).
// Winner flags only: ceil(n/32) CTAs side by side.  Ordered (unique_ids, index) output: one CTA, which
// then ranks the winners by id.  Deterministic, no scratch tables.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
select_latest_small_kernel(const int64_t* __restrict__ nids, const void* __restrict__ ts, int is_f64, int n,
                           int64_t ts_period, uint8_t* __restrict__ winner, int64_t* __restrict__ unique_ids,
                           int64_t* __restrict__ index, int32_t* __restrict__ count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int64_t* s_id = reinterpret_cast<int64_t*>(smem_raw);
  uint64_t* s_key = reinterpret_cast<uint64_t*>(s_id + n);
  uint8_t* s_win = reinterpret_cast<uint8_t*>(s_key + n);
  __shared__ int s_count;
  const int lane = lane_id(), warp = warp_id_in_block(), n_warps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_count = 0;
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    s_id[p] = nids[p];
    s_key[p] = load_key(ts, is_f64, p % ts_period);
  }
  __syncthreads();
  for (int p = blockIdx.x * n_warps + warp; p < n; p += gridDim.x * n_warps) {
    const int64_t id = s_id[p];
    const uint64_t key = s_key[p];
    bool beaten = false;
    for (int q = lane; q < n; q += 32) {
      if (s_id[q] == id) {
        const uint64_t kq = s_key[q];
        beaten |= (kq > key) || (kq == key && q < p);
      }
    }
    const bool win = !__any_sync(TIGER_FULL_MASK, beaten);
    if (lane == 0) {
      if (unique_ids != nullptr) s_win[p] = win;
      if (winner != nullptr) winner[p] = win;
      if (win && count != nullptr && gridDim.x == 1) atomicAdd(&s_count, 1);
    }
  }
  if (gridDim.x != 1) return;          // winner-only mode
  __syncthreads();
  if (count != nullptr && threadIdx.x == 0) *count = s_count;
  if (unique_ids == nullptr) return;
  for (int p = warp; p < n; p += n_warps) {
    if (!s_win[p]) continue;
    const int64_t id = s_id[p];
    int rank = 0;
    for (int q = lane; q < n; q += 32) rank += (s_win[q] && s_id[q] < id);
    rank = (int)warp_sum((float)rank);   // n <= 2048: exact in float
    if (lane == 0) {
      unique_ids[rank] = id;
      index[rank] = p;
    }
  }
}

// ------------------------------------------------------------------------------------------
// large n: per-node slots + atomics (max timestamp, then min position), flag, ordered compaction
// ------------------------------------------------------------------------------------------
__global__ void select_max_ts_kernel(const int64_t* __restrict__ nids, const void* __restrict__ ts, int is_f64,
                                     int64_t n, int64_t ts_period, uint64_t* __restrict__ slot_ts) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  atomicMax(reinterpret_cast<unsigned long long*>(slot_ts + nids[p]),
            (unsigned long long)load_key(ts, is_f64, p % ts_period));
}

__global__ void select_min_pos_kernel(const int64_t* __restrict__ nids, const void* __restrict__ ts, int is_f64,
                                      int64_t n, int64_t ts_period, const uint64_t* __restrict__ slot_ts,
                                      uint32_t* __restrict__ slot_pos) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int64_t id = nids[p];
  if (slot_ts[id] == load_key(ts, is_f64, p % ts_period)) atomicMax(slot_pos + id, 0xffffffffu - (uint32_t)p);
}

__global__ void select_flag_kernel(const int64_t* __restrict__ nids, int64_t n,
                                   const uint32_t* __restrict__ slot_pos, uint8_t* __restrict__ winner,
                                   uint32_t* __restrict__ bitmap, int32_t* __restrict__ count) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool win = false;
  if (p < n) {
    const int64_t id = nids[p];
    win = slot_pos[id] == 0xffffffffu - (uint32_t)p;
    if (winner != nullptr) winner[p] = win;
    if (win && bitmap != nullptr) atomicOr(bitmap + (id >> 5), 1u << (id & 31));
  }
  // flags-only mode: the number of winners (= distinct ids) is an integer sum, order-independent
  if (count != nullptr) {
    const unsigned votes = __ballot_sync(TIGER_FULL_MASK, win);
    if (lane_id() == 0 && votes) atomicAdd(count, __popc(votes));
  }
}

// scratch reset when no ordered output was requested (idempotent writes)
__global__ void select_reset_kernel(const int64_t* __restrict__ nids, int64_t n, uint64_t* __restrict__ slot_ts,
                                    uint32_t* __restrict__ slot_pos) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  slot_ts[nids[p]] = 0ull;
  slot_pos[nids[p]] = 0u;
}

__global__ void __launch_bounds__(1024)
select_compact_kernel(uint32_t* __restrict__ bitmap, int64_t n_words, uint64_t* __restrict__ slot_ts,
                      uint32_t* __restrict__ slot_pos, int64_t* __restrict__ unique_ids,
                      int64_t* __restrict__ index, int32_t* __restrict__ count) {
  __shared__ int warp_sums[32];
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id_in_block();
  const int64_t wpt = (n_words + blockDim.x - 1) / blockDim.x;
  const int64_t w0 = (int64_t)tid * wpt;
  const int64_t w1 = (w0 + wpt < n_words) ? (w0 + wpt) : n_words;
  int c = 0;
  for (int64_t w = w0; w < w1; ++w) c += __popc(bitmap[w]);
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(TIGER_FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int v = warp_sums[lane];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(TIGER_FULL_MASK, s, o);
      if (lane >= o) s += t;
    }
    warp_sums[lane] = s - v;
    if (lane == 31 && count != nullptr) *count = s;
  }
  __syncthreads();
  int o_out = inc - c + warp_sums[warp];
  for (int64_t w = w0; w < w1; ++w) {
    uint32_t bits = bitmap[w];
    if (bits) bitmap[w] = 0u;
    while (bits) {
      const int64_t u = w * 32 + (__ffs(bits) - 1);
      bits &= bits - 1;
      unique_ids[o_out] = u;
      index[o_out] = (int64_t)(0xffffffffu - slot_pos[u]);
      slot_pos[u] = 0u;
      slot_ts[u] = 0ull;
      ++o_out;
    }
  }
}

extern "C" int tiger_select_latest(const int64_t* nids, const void* ts, int ts_is_f64, int64_t n,
                                   int64_t ts_period, int64_t n_nodes, uint64_t* slot_ts, uint32_t* slot_pos, uint32_t* bitmap,
                                   uint8_t* winner, int64_t* unique_ids, int64_t* index, int32_t* count,
                                   void* stream) {
  if (n < 0 || ts_period < 0 || (unique_ids == nullptr) != (index == nullptr)) return TIGER_EINVAL;
  if (ts_period == 0) ts_period = n > 0 ? n : 1;
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    if (count != nullptr) cudaMemsetAsync(count, 0, sizeof(int32_t), st);
    return tiger_launch_status();
  }
  if (n <= SELECT_SMALL_N) {
    const size_t smem = (size_t)n * (sizeof(int64_t) + sizeof(uint64_t) + 1);
    // winner flags only (and no count requested): spread the positions over several CTAs
    const bool flags_only = unique_ids == nullptr && count == nullptr;
    const unsigned grid = flags_only ? (unsigned)((n + 31) / 32) : 1u;
    select_latest_small_kernel<<<grid, 1024, smem, st>>>(nids, ts, ts_is_f64, (int)n, ts_period, winner, unique_ids,
                                                        index, count);
    return tiger_launch_status();
  }
  if (slot_ts == nullptr || slot_pos == nullptr || n >= 0xffffffffll) return TIGER_EINVAL;
  if (unique_ids != nullptr && bitmap == nullptr) return TIGER_EINVAL;
  const unsigned grid = (unsigned)((n + 255) / 256);
  select_max_ts_kernel<<<grid, 256, 0, st>>>(nids, ts, ts_is_f64, n, ts_period, slot_ts);
  select_min_pos_kernel<<<grid, 256, 0, st>>>(nids, ts, ts_is_f64, n, ts_period, slot_ts, slot_pos);
  // without the ordered output nothing else writes *count (select_compact_kernel does when it runs)
  const bool count_in_flags = unique_ids == nullptr && count != nullptr;
  if (count_in_flags) cudaMemsetAsync(count, 0, sizeof(int32_t), st);
  select_flag_kernel<<<grid, 256, 0, st>>>(nids, n, slot_pos, winner, unique_ids != nullptr ? bitmap : nullptr,
                                           count_in_flags ? count : nullptr);
  if (unique_ids != nullptr) {
    select_compact_kernel<<<1, 1024, 0, st>>>(bitmap, (n_nodes + 31) / 32, slot_ts, slot_pos, unique_ids, index,
                                              count);
  } else {
    select_reset_kernel<<<grid, 256, 0, st>>>(nids, n, slot_ts, slot_pos);
  }
  return tiger_launch_status();
}

// ------------------------------------------------------------------------------------------
// anonymized_reindex: one CTA per history row.  rank(v) = 1 + #distinct values whose last
// occurrence lies to the right of v's last occurrence; padding (0) stays 0 but, like the
// reference's OrderedDict, still counts as a distinct value for the ids to its left.
// ------------------------------------------------------------------------------------------
__global__ void anonymized_reindex_kernel(const int64_t* __restrict__ hist, int len, const int32_t* __restrict__ count,
                                          int64_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int64_t* s_v = reinterpret_cast<int64_t*>(smem_raw);
  uint8_t* s_last = reinterpret_cast<uint8_t*>(s_v + len);
  const int64_t row = blockIdx.x;
  if (count != nullptr && row >= *count) return;
  for (int j = threadIdx.x; j < len; j += blockDim.x) s_v[j] = hist[row * len + j];
  __syncthreads();
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    bool last = true;
    for (int q = j + 1; q < len; ++q) last = last && (s_v[q] != s_v[j]);
    s_last[j] = last;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    const int64_t v = s_v[j];
    int rank = 0;
    if (v != 0) {
      int q = j;
      while (!(s_last[q] && s_v[q] == v)) ++q;  // v's last occurrence
      rank = 1;
      for (int r = q + 1; r < len; ++r) rank += s_last[r];
    }
    out[row * len + j] = rank;
  }
}

extern "C" int tiger_anonymized_reindex(const int64_t* hist_nids, int64_t n, int len, int64_t* out,
                                        const int32_t* count, void* stream) {
  if (n < 0 || len <= 0 || len > 4096) return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  const int threads = len >= 256 ? 256 : (len + 31) / 32 * 32;
  anonymized_reindex_kernel<<<(unsigned)n, threads, (size_t)len * 9, as_stream(stream)>>>(hist_nids, len, count, out);
  return tiger_launch_status();
}
