// Link scorer: step 7 of TIGE.contrast_learning (tiger/model/tiger.py:259-288), eval mode:
// hit flags ('bin': tiger.py:266-270, data_loader.py:61-75) -> hit embedding -> score_fn
// MergeLayer (basic_modules.py:16-19) on positive and negative pairs -> BCE-with-logits mean.
#include "common.cuh"

#define SCORE_GP 8  // pairs per CTA

__global__ void __launch_bounds__(512)
link_score_kernel(const float* __restrict__ h, int64_t batch, int d, const int64_t* __restrict__ src,
                  const int64_t* __restrict__ dst, const int64_t* __restrict__ neg,
                  const int64_t* __restrict__ neigh, int k, const float* __restrict__ hit_emb,
                  const float* __restrict__ fc1T, int ld, const float* __restrict__ fc1_b,
                  const float* __restrict__ fc2_w, const float* __restrict__ fc2_b, float* __restrict__ scores,
                  float* __restrict__ loss, uint32_t* __restrict__ done_counter) {
  extern __shared__ __align__(16) float sm[];
  float* xin = sm;                       // [2d][GP]
  float* red = xin + 2 * d * SCORE_GP;   // [warps][GP]
  __shared__ int s_flag[SCORE_GP][2];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  const int64_t p0 = (int64_t)blockIdx.x * SCORE_GP;
  const int64_t n_pairs = 2 * batch;
  // hit flags: one warp per (pair, side)
  for (int i = warp; i < SCORE_GP * 2; i += n_warps) {
    const int g = i >> 1, side = i & 1;
    const int64_t p = p0 + g;
    int flag = 0;
    if (p < n_pairs && hit_emb != nullptr) {
      const bool is_neg = p >= batch;
      const int64_t e = is_neg ? p - batch : p;
      int64_t center, row;
      if (!is_neg) {
        center = side == 0 ? src[e] : dst[e];          // src_hits: src in N(dst) ; dst_hits: dst in N(src)
        row = side == 0 ? batch + e : e;
      } else {
        center = side == 0 ? src[e] : neg[e];          // neg_src_hits: src in N(neg) ; neg_dst_hits: neg in N(src)
        row = side == 0 ? 2 * batch + e : e;
      }
      for (int j = lane; j < k; j += 32) flag |= (neigh[row * k + j] == center);
      flag = __any_sync(TIGER_FULL_MASK, flag);
    }
    if (lane == 0) s_flag[g][side] = flag;
  }
  __syncthreads();
  // inputs [x + he | y + he] in [k][GP] layout
  for (int i = tid; i < SCORE_GP * 2 * d; i += blockDim.x) {
    const int g = i / (2 * d), c = i % (2 * d);
    const int64_t p = p0 + g;
    float v = 0.f;
    if (p < n_pairs) {
      const bool is_neg = p >= batch;
      const int64_t e = is_neg ? p - batch : p;
      const int side = c >= d;
      const int cc = side ? c - d : c;
      const int64_t row = side == 0 ? e : (is_neg ? 2 * batch + e : batch + e);
      v = h[row * d + cc];
      if (hit_emb != nullptr) v += hit_emb[s_flag[g][side] * d + cc];
    }
    xin[c * SCORE_GP + g] = v;
  }
  __syncthreads();
  float part[SCORE_GP];
#pragma unroll
  for (int g = 0; g < SCORE_GP; ++g) part[g] = 0.f;
  for (int n = tid; n < d; n += blockDim.x) {
    float acc[SCORE_GP];
#pragma unroll
    for (int g = 0; g < SCORE_GP; ++g) acc[g] = 0.f;
#pragma unroll 4
    for (int c = 0; c < 2 * d; ++c) {
      const float w = __ldg(fc1T + (int64_t)c * ld + n);
      const float4 x0 = *reinterpret_cast<const float4*>(xin + c * SCORE_GP);
      const float4 x1 = *reinterpret_cast<const float4*>(xin + c * SCORE_GP + 4);
      acc[0] = fmaf(w, x0.x, acc[0]); acc[1] = fmaf(w, x0.y, acc[1]);
      acc[2] = fmaf(w, x0.z, acc[2]); acc[3] = fmaf(w, x0.w, acc[3]);
      acc[4] = fmaf(w, x1.x, acc[4]); acc[5] = fmaf(w, x1.y, acc[5]);
      acc[6] = fmaf(w, x1.z, acc[6]); acc[7] = fmaf(w, x1.w, acc[7]);
    }
    const float b = fc1_b[n], w2 = fc2_w[n];
#pragma unroll
    for (int g = 0; g < SCORE_GP; ++g) part[g] = fmaf(fmaxf(acc[g] + b, 0.f), w2, part[g]);
  }
#pragma unroll
  for (int g = 0; g < SCORE_GP; ++g) {
    const float t = warp_sum(part[g]);
    if (lane == 0) red[warp * SCORE_GP + g] = t;
  }
  __syncthreads();
  if (tid < SCORE_GP && p0 + tid < n_pairs) {
    float s = 0.f;
    for (int w = 0; w < n_warps; ++w) s += red[w * SCORE_GP + tid];
    scores[p0 + tid] = s + fc2_b[0];
  }
  if (loss == nullptr) return;
  // last CTA reduces the BCE-with-logits mean in a fixed order
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float acc = 0.f;
  for (int64_t p = tid; p < n_pairs; p += blockDim.x) {
    const float x = __ldcg(scores + p);
    const float y = p < batch ? 1.f : 0.f;
    acc += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < n_warps; ++w) s += red[w];
    *loss = s / (float)n_pairs;
    *done_counter = 0u;
  }
}

extern "C" int tiger_link_score(const float* h, int64_t batch, int d, const int64_t* src, const int64_t* dst,
                                const int64_t* neg, const int64_t* neigh_nids, int k, const float* hit_emb,
                                const float* fc1T, const float* fc1_b, const float* fc2_w, const float* fc2_b,
                                float* scores, float* loss, uint32_t* done_counter, void* stream) {
  if (batch < 0 || d <= 0 || (hit_emb != nullptr && (neigh_nids == nullptr || k <= 0))) return TIGER_EINVAL;
  if (loss != nullptr && done_counter == nullptr) return TIGER_EINVAL;
  if (batch == 0) return TIGER_OK;
  int threads = (d + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  if (threads > 512) threads = 512;
  const int ld = (d + 3) / 4 * 4;
  const size_t smem = ((size_t)2 * d * SCORE_GP + (size_t)(threads / 32) * SCORE_GP + 32) * sizeof(float);
  if (smem > 48 * 1024) return TIGER_EINVAL;
  const unsigned grid = (unsigned)((2 * batch + SCORE_GP - 1) / SCORE_GP);
  link_score_kernel<<<grid, threads, smem, as_stream(stream)>>>(h, batch, d, src, dst, neg, neigh_nids, k, hit_emb,
                                                               fc1T, ld, fc1_b, fc2_w, fc2_b, scores, loss,
                                                               done_counter);
  return tiger_launch_status();
}
