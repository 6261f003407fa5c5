// Seq restarter: SeqRestarter.forward (tiger/model/restarters.py:51-114) as a short launch sequence
//   tiger_min_time        restart() is called with ts.min() of the batch (train_self_supervised.py:161)
//   tiger_find_recent     get_history(nids, t, hist_len)                       (graph.cu)
//   tiger_anonymized_reindex                                                   (select.cu)
//   tiger_seq_tokens      event tokens [src | dst | anony | edge | time], key-padding mask, prev_ts
//   tiger_sgemm_nt        packed q,k in-projection of all tokens               (gemm.cu)
//   tiger_seq_attn_pool   per (node, head): L x L scores, masked softmax, position-mean of the
//                         attention weights, pooled tokens xbar_h = sum_i pbar_hi x_i
//   tiger_sgemm_nt x5     value projection of the pooled tokens (per head), out-proj (+ReLU), out_fn,
//                         merger fc1 (+ReLU, first d input columns only), merger fc2
// Folding (exact by linearity, because the reference takes the MEAN over positions of the MHA output
// before anything non-linear): mean_j out_j = W_o [concat_h W_v,h xbar_h + b_v,h] + b_o.  The value
// projection and the out-projection therefore run on n rows instead of n * L.
// last_event_feat is identically zero in the reference (a view that is zeroed in place one line
// later, restarters.py:103-104), so the merger only sees h_prev_left.
#include "common.cuh"

// ------------------------------------------------------------------------------------------
__global__ void min_time_kernel(const float* __restrict__ ts, int64_t n, double* __restrict__ out) {
  __shared__ float red[32];
  float m = INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fminf(m, ts[i]);
  m = warp_min(m);
  if (lane_id() == 0) red[warp_id_in_block()] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : INFINITY;
    m = warp_min(m);
    if (threadIdx.x == 0) out[0] = (double)m;
  }
}

extern "C" int tiger_min_time(const float* ts, int64_t n, double* out, void* stream) {
  if (n <= 0 || out == nullptr) return TIGER_EINVAL;
  min_time_kernel<<<1, 256, 0, as_stream(stream)>>>(ts, n, out);
  return tiger_launch_status();
}

// ------------------------------------------------------------------------------------------
// tokens: one warp per (node, position)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
seq_tokens_kernel(const int64_t* __restrict__ nids, const int32_t* __restrict__ count, int64_t n, int len,
                  const int64_t* __restrict__ hist_nids, const int64_t* __restrict__ hist_eids,
                  const float* __restrict__ hist_ts, const int64_t* __restrict__ hist_dirs,
                  const int64_t* __restrict__ anony_ids, const float* __restrict__ nfeats,
                  const float* __restrict__ efeats, int d, int de, const float* __restrict__ anony_emb,
                  const float* __restrict__ time_w, const float* __restrict__ time_b, float* __restrict__ x,
                  uint8_t* __restrict__ mask, float* __restrict__ prev_ts) {
  int64_t total = n;
  if (count != nullptr) {
    const int64_t c = *count;
    total = c < n ? c : n;
  }
  total *= len;
  const int lane = lane_id();
  const int64_t dm = 4 * (int64_t)d + de;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id_in_block(); r < total; r += n_warps) {
    const int64_t i = r / len;
    const int j = (int)(r % len);
    const int64_t owner = nids[i];
    const int64_t hn = hist_nids[r], dir = hist_dirs[r];
    const float t_last = hist_ts[i * len + len - 1];
    float* row = x + r * dm;
    const bool last = j == len - 1;
    if (lane == 0) {
      mask[r] = (hn == 0) && !last;                      // mask[:, -1] = False (restarters.py:87)
      if (last) prev_ts[i] = t_last;
    }
    if (last) {
      for (int c = lane; c < 3 * d + de; c += 32) row[c] = 0.f;   // restarters.py:104
    } else {
      // the reference's literal direction rule (restarters.py:93-94)
      const int64_t s_id = owner * dir + hn * (1 - dir);
      const int64_t d_id = owner * (1 - dir) + hn * dir;
      if (nfeats != nullptr) {
        warp_copy_row(row, nfeats + s_id * d, d, lane);
        warp_copy_row(row + d, nfeats + d_id * d, d, lane);
      } else {
        for (int c = lane; c < 2 * d; c += 32) row[c] = 0.f;
      }
      warp_copy_row(row + 2 * d, anony_emb + anony_ids[r] * d, d, lane);
      if (efeats != nullptr) {
        warp_copy_row(row + 3 * d, efeats + hist_eids[r] * de, de, lane);
      } else {
        for (int c = lane; c < de; c += 32) row[3 * d + c] = 0.f;
      }
    }
    const float dt = t_last - hist_ts[r];
    float* trow = row + 3 * d + de;
    for (int c = lane; c < d; c += 32) trow[c] = time_enc(dt, time_w[c], time_b[c]);
  }
}

extern "C" int tiger_seq_tokens(const int64_t* nids, const int32_t* count, int64_t n, int len,
                                const int64_t* hist_nids, const int64_t* hist_eids, const float* hist_ts,
                                const int64_t* hist_dirs, const int64_t* anony_ids, const float* nfeats,
                                const float* efeats, int d, int de, const float* anony_emb, const float* time_w,
                                const float* time_b, float* x, uint8_t* mask, float* prev_ts, void* stream) {
  if (n < 0 || len <= 0 || d <= 0 || de <= 0) return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  const int64_t rows = n * len;
  int64_t grid = (rows + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  seq_tokens_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(nids, count, n, len, hist_nids, hist_eids,
                                                                  hist_ts, hist_dirs, anony_ids, nfeats, efeats, d,
                                                                  de, anony_emb, time_w, time_b, x, mask, prev_ts);
  return tiger_launch_status();
}

// ------------------------------------------------------------------------------------------
// per (node, head): scores -> masked softmax -> mean over query positions -> pooled tokens
// ------------------------------------------------------------------------------------------
#define POOL_THREADS 256
#define POOL_CW 32       // head-dim chunk staged in shared memory
#define POOL_MAXL 64     // hist_len limit (default 40)
#define POOL_MAXP ((POOL_MAXL * POOL_MAXL + POOL_THREADS - 1) / POOL_THREADS)

__global__ void __launch_bounds__(POOL_THREADS)
seq_attn_pool_kernel(const float* __restrict__ qk, int64_t ld_qk, const float* __restrict__ x,
                     const uint8_t* __restrict__ mask, const int32_t* __restrict__ count, int64_t n, int len,
                     int dm, int n_head, float* __restrict__ xbar) {
  __shared__ float qs[POOL_MAXL][POOL_CW + 1];
  __shared__ float ks[POOL_MAXL][POOL_CW + 1];
  __shared__ float sc[POOL_MAXL][POOL_MAXL + 1];
  __shared__ float pbar[POOL_MAXL];
  int64_t total = n;
  if (count != nullptr) {
    const int64_t c = *count;
    total = c < n ? c : n;
  }
  total *= n_head;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id_in_block();
  const int hd = dm / n_head;
  const float scale = sqrtf(1.0f / (float)hd);
  const int n_pairs = len * len;
  for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const int64_t i = item / n_head;
    const int h = (int)(item % n_head);
    const float* qbase = qk + i * len * ld_qk + h * hd;
    const float* kbase = qbase + dm;
    float acc[POOL_MAXP];
#pragma unroll
    for (int r = 0; r < POOL_MAXP; ++r) acc[r] = 0.f;
    for (int c0 = 0; c0 < hd; c0 += POOL_CW) {
      const int cw = (hd - c0) < POOL_CW ? (hd - c0) : POOL_CW;
      __syncthreads();
      for (int e = tid; e < len * POOL_CW; e += POOL_THREADS) {
        const int j = e / POOL_CW, c = e % POOL_CW;
        float qv = 0.f, kv = 0.f;
        if (c < cw) {
          qv = qbase[(int64_t)j * ld_qk + c0 + c] * scale;   // torch scales q before q.k^T
          kv = kbase[(int64_t)j * ld_qk + c0 + c];
        }
        qs[j][c] = qv;
        ks[j][c] = kv;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < POOL_MAXP; ++r) {
        const int p = tid + r * POOL_THREADS;
        if (p < n_pairs) {
          const int j = p / len, i2 = p % len;
          float s = acc[r];
#pragma unroll 8
          for (int c = 0; c < POOL_CW; ++c) s = fmaf(qs[j][c], ks[i2][c], s);
          acc[r] = s;
        }
      }
    }
    const uint8_t* mrow = mask + i * len;
#pragma unroll
    for (int r = 0; r < POOL_MAXP; ++r) {
      const int p = tid + r * POOL_THREADS;
      if (p < n_pairs) {
        const int j = p / len, i2 = p % len;
        sc[j][i2] = mrow[i2] ? -INFINITY : acc[r];
      }
    }
    __syncthreads();
    for (int j = warp; j < len; j += POOL_THREADS / 32) {       // softmax over keys, one warp per query row
      float m = -INFINITY;
      for (int c = lane; c < len; c += 32) m = fmaxf(m, sc[j][c]);
      m = warp_max(m);
      float sum = 0.f;
      for (int c = lane; c < len; c += 32) {
        const float e = expf(sc[j][c] - m);
        sc[j][c] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      for (int c = lane; c < len; c += 32) sc[j][c] = sc[j][c] / sum;
    }
    __syncthreads();
    if (tid < len) {                                             // mean over ALL query positions (restarters.py:108)
      float s = 0.f;
      for (int j = 0; j < len; ++j) s += sc[j][tid];
      pbar[tid] = s / (float)len;
    }
    __syncthreads();
    const float* xrow = x + i * len * (int64_t)dm;
    float* orow = xbar + (i * n_head + h) * (int64_t)dm;
    for (int c = tid; c < dm; c += POOL_THREADS) {
      float s = 0.f;
      for (int j = 0; j < len; ++j) s = fmaf(pbar[j], xrow[(int64_t)j * dm + c], s);
      orow[c] = s;
    }
  }
}

extern "C" int tiger_seq_attn_pool(const float* qk, int64_t ld_qk, const float* x, const uint8_t* mask,
                                   const int32_t* count, int64_t n, int len, int d_model, int n_head, float* xbar,
                                   void* stream) {
  if (n < 0 || len <= 0 || len > POOL_MAXL || d_model <= 0 || n_head <= 0 || d_model % n_head != 0 ||
      ld_qk < 2 * (int64_t)d_model)
    return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  int64_t grid = n * n_head;
  if (grid > 148 * 4) grid = 148 * 4;
  seq_attn_pool_kernel<<<(unsigned)grid, POOL_THREADS, 0, as_stream(stream)>>>(qk, ld_qk, x, mask, count, n, len,
                                                                              d_model, n_head, xbar);
  return tiger_launch_status();
}

// ------------------------------------------------------------------------------------------
// Small-R tail of the seq restarter.  After the pooling kernel the path is five dependent layers on the n pooled
// rows - value projection per head, out-projection (+ReLU), out_fn, merger fc1 (+ReLU), fc2 - and in steady state
// n is ~20-50 (the nodes a batch touches for the first time).  As tensor-core products each layer is one CTA per
// tile walking 54 dependent k-steps (K = d_model = 860): 6 launches x 33 us for a few hundred MFLOP.  Here one CTA
// owns one row and runs all five layers as matrix-vector products out of shared memory, one warp per output
// channel, weights streamed from L2 with 16-byte loads (6.6 MB per row).  Above `max_rows` restarted rows (the
// first batches of a chunk) the kernel exits at once and the GEMM chain runs instead: tiger_seq_gate_count splits
// the device-side row count into (count if <= max_rows else 0, count if > max_rows else 0).
// ------------------------------------------------------------------------------------------
__global__ void seq_gate_count_kernel(const int32_t* __restrict__ count, int32_t max_rows, int32_t* __restrict__ small,
                                      int32_t* __restrict__ big) {
  const int32_t c = *count;
  *small = c <= max_rows ? c : 0;
  *big = c > max_rows ? c : 0;
}

extern "C" int tiger_seq_gate_count(const int32_t* count, int max_rows, int32_t* count_small, int32_t* count_big,
                                    void* stream) {
  if (count == nullptr || count_small == nullptr || count_big == nullptr) return TIGER_EINVAL;
  seq_gate_count_kernel<<<1, 1, 0, as_stream(stream)>>>(count, max_rows, count_small, count_big);
  return tiger_launch_status();
}

// y[c] = act(bias[c] + W[c, :k] . x) for c in [0, n_out): one warp per output channel
__device__ __forceinline__ void tail_layer(const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias,
                                           const float* x, int k, int n_out, bool relu, float* y, int warp, int n_warps,
                                           int lane) {
  const bool vec = ((((uintptr_t)W) & 15) == 0) && (ldw & 3) == 0 && (k & 3) == 0;
  for (int c = warp; c < n_out; c += n_warps) {
    const float* w = W + (int64_t)c * ldw;
    float acc = 0.f;
    if (vec) {
      const float4* w4 = reinterpret_cast<const float4*>(w);
      const float4* x4 = reinterpret_cast<const float4*>(x);
      for (int i = lane; i < (k >> 2); i += 32) {
        const float4 a = __ldg(w4 + i), b = x4[i];
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
    } else {
      for (int i = lane; i < k; i += 32) acc = fmaf(__ldg(w + i), x[i], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = acc + bias[c];
      y[c] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

#define TAIL_THREADS 512
__global__ void __launch_bounds__(TAIL_THREADS)
seq_tail_kernel(const float* __restrict__ xbar, const int32_t* __restrict__ count, int64_t n_cap, int dm, int n_head,
                int d, const float* __restrict__ wv, const float* __restrict__ bv, const float* __restrict__ wo,
                const float* __restrict__ bo, const float* __restrict__ wfn, const float* __restrict__ bfn,
                const float* __restrict__ wfc1, int64_t ld_fc1, const float* __restrict__ bfc1,
                const float* __restrict__ wfc2, const float* __restrict__ bfc2, float* __restrict__ h_left,
                float* __restrict__ h_right) {
  extern __shared__ __align__(16) float tail_smem[];
  int64_t n = n_cap;
  if (count != nullptr) {
    const int64_t c = *count;
    n = c < n ? c : n;
  }
  const int hd = dm / n_head;
  const int dmp = (dm + 3) & ~3, dp = (d + 3) & ~3;
  float* xb = tail_smem;                       // [n_head][dmp]
  float* att = xb + n_head * dmp;              // [dmp]
  float* o = att + dmp;                        // [dmp]
  float* hl = o + dmp;                         // [dp]
  float* hid = hl + dp;                        // [dp]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = TAIL_THREADS / 32;
  for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < n_head * dm; i += TAIL_THREADS) xb[(i / dm) * dmp + (i % dm)] = xbar[r * n_head * dm + i];
    __syncthreads();
    for (int h = 0; h < n_head; ++h)            // value projection of the head's pooled tokens
      tail_layer(wv + (int64_t)h * hd * dm, dm, bv + h * hd, xb + h * dmp, dm, hd, false, att + h * hd, warp, n_warps, lane);
    __syncthreads();
    tail_layer(wo, dm, bo, att, dm, dm, true, o, warp, n_warps, lane);
    __syncthreads();
    tail_layer(wfn, dm, bfn, o, dm, d, false, hl, warp, n_warps, lane);
    __syncthreads();
    for (int c = tid; c < d; c += TAIL_THREADS) h_left[r * d + c] = hl[c];
    tail_layer(wfc1, ld_fc1, bfc1, hl, d, d, true, hid, warp, n_warps, lane);
    __syncthreads();
    tail_layer(wfc2, d, bfc2, hid, d, d, false, hl, warp, n_warps, lane);
    __syncthreads();
    for (int c = tid; c < d; c += TAIL_THREADS) h_right[r * d + c] = hl[c];
  }
}

extern "C" int tiger_seq_tail(const float* xbar, const int32_t* count, int64_t n, int d_model, int n_head, int d,
                              const float* w_v, const float* b_v, const float* w_out, const float* b_out,
                              const float* w_fn, const float* b_fn, const float* w_fc1, int64_t ld_fc1, const float* b_fc1,
                              const float* w_fc2, const float* b_fc2, int max_rows, float* h_left, float* h_right,
                              void* stream) {
  if (xbar == nullptr || w_v == nullptr || w_out == nullptr || w_fn == nullptr || w_fc1 == nullptr || w_fc2 == nullptr ||
      h_left == nullptr || h_right == nullptr || n < 0 || d_model <= 0 || n_head <= 0 || d_model % n_head != 0 || d <= 0 ||
      max_rows <= 0 || ld_fc1 < d)
    return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  const int dmp = (d_model + 3) & ~3, dp = (d + 3) & ~3;
  const size_t smem = (size_t)((n_head + 2) * dmp + 2 * dp) * sizeof(float);
  if (smem > 48 * 1024) return TIGER_EINVAL;
  const int64_t grid = n < max_rows ? n : max_rows;
  seq_tail_kernel<<<(unsigned)grid, TAIL_THREADS, smem, as_stream(stream)>>>(
      xbar, count, n, d_model, n_head, d, w_v, b_v, w_out, b_out, w_fn, b_fn, w_fc1, ld_fc1, b_fc1, w_fc2, b_fc2, h_left,
      h_right);
  return tiger_launch_status();
}
