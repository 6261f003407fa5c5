// Training step of the seq restarter (SeqRestarter.forward, tiger/model/restarters.py:51-114, under autograd in
// TIGER.contrast_and_mutual_learning, tiger/model/tiger.py:574-590): the attention pieces between the tensor-core
// products, forward with dropout and backward.
//
// The forward keeps the folding of csrc/restart_seq.cu, which stays exact under dropout because everything between
// the attention probabilities and the first non-linearity is linear:
//   mean_i out_i = W_o [ concat_h ( W_v,h xbar_h + b_v,h psum_h ) ] + b_o,
//   xbar_h = sum_j pbar_hj x_j,   pbar_hj = mean_i P'_h,ij  (P' = dropout(softmax)),   psum_h = sum_j pbar_hj
// (psum_h = 1 without dropout).  Only the q / k in-projection runs on all n * L tokens.
#include "common.cuh"

__device__ __forceinline__ float seq_sin_reduced(float x) {
  const double xd = (double)x;
  const double n = rint(xd * 0.15915494309189535);
  double r = fma(-n, 6.283185307179586, xd);
  r = fma(-n, 2.4492935982947064e-16, r);
  return sinf((float)r);
}

__device__ __forceinline__ int64_t seq_rows(const int32_t* count, int64_t n, int64_t per = 1) {
  if (count == nullptr) return n;
  const int64_t c = (int64_t)(*count) * per;
  return c < n ? c : n;
}

#define SP_THREADS 256
#define SP_CW 32
#define SP_MAXL 64
#define SP_MAXP ((SP_MAXL * SP_MAXL + SP_THREADS - 1) / SP_THREADS)
#define SP_PRE ((SP_MAXL * SP_CW + SP_THREADS - 1) / SP_THREADS)

// ------------------------------------------------------------------------------------------
// forward: per (node, head) L x L scores, key-padding mask, softmax, dropout, mean over the query positions,
// pooled tokens.  Keeps P (softmax output) [n, H, L, L], pbar [n, H, L], psum [n, H].
// Scores: the head's q / k columns are staged in chunks of 32 columns, TRANSPOSED ([column][position]), and every
// thread owns a 2 x 4 block of (query, key) pairs: per column one 8-byte and one 16-byte shared-memory load feed
// 8 FMAs (the first version read two scalars per FMA and was bound by the shared-memory pipe: 207 us for 562
// items, profiles/r02_train_kernels.md).
// ------------------------------------------------------------------------------------------
#define SP_LP 68      // padded positions per staged column: multiple of 4 (vector loads), 68 % 32 = 4 spreads the stores
__global__ void __launch_bounds__(SP_THREADS)
train_seq_pool_kernel(const float* __restrict__ qk, int64_t ld_qk, const float* __restrict__ x,
                      const uint8_t* __restrict__ mask, const int32_t* __restrict__ count, int64_t n, int len, int dm,
                      int n_head, float p_drop, uint32_t seed, float* __restrict__ P, float* __restrict__ pbar_out,
                      float* __restrict__ psum_out, float* __restrict__ xbar) {
  seed = tiger_step_seed(seed);
  __shared__ __align__(16) float qsT[SP_CW][SP_LP];
  __shared__ __align__(16) float ksT[SP_CW][SP_LP];
  __shared__ float sc[SP_MAXL][SP_MAXL + 1];
  __shared__ float pbar[SP_MAXL];
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id_in_block();
  const int hd = dm / n_head;
  const float scale = sqrtf(1.0f / (float)hd);
  const float inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const int n_pairs = len * len;
  const int q_tiles = (len + 1) / 2, k_tiles = (len + 3) / 4, n_tiles = q_tiles * k_tiles;   // <= 32 * 16
  const int64_t total = seq_rows(count, n) * n_head;
  for (int e = tid; e < SP_CW * SP_LP; e += SP_THREADS) {       // the padding positions stay zero
    (&qsT[0][0])[e] = 0.f;
    (&ksT[0][0])[e] = 0.f;
  }
  for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const int64_t i = item / n_head;
    const int h = (int)(item % n_head);
    const float* qbase = qk + i * len * ld_qk + h * hd;
    const float* kbase = qbase + dm;
    float acc[2][8];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[t][r] = 0.f;
    // the next chunk's q / k elements are fetched into registers while the current chunk is multiplied (with ~23 items
    // per launch in the inference step the kernel is one CTA's latency chain: 14 chunks x (global load -> smem -> FMA))
    float qreg[SP_PRE], kreg[SP_PRE];
    auto fetch = [&](int c0) {
#pragma unroll
      for (int t = 0; t < SP_PRE; ++t) {
        const int e = tid + t * SP_THREADS;
        const int j = e / SP_CW, c = e % SP_CW;
        float qv = 0.f, kv = 0.f;
        if (j < len && c0 + c < hd) {
          qv = qbase[(int64_t)j * ld_qk + c0 + c];               // consumed (scaled) only when stored to shared memory:
          kv = kbase[(int64_t)j * ld_qk + c0 + c];               // multiplying here would wait for the load at once
        }
        qreg[t] = qv;
        kreg[t] = kv;
      }
    };
    fetch(0);
    for (int c0 = 0; c0 < hd; c0 += SP_CW) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < SP_PRE; ++t) {
        const int e = tid + t * SP_THREADS;
        const int j = e / SP_CW, c = e % SP_CW;
        if (j < len) {
          qsT[c][j] = qreg[t] * scale;                           // torch scales q before q.k^T
          ksT[c][j] = kreg[t];
        }
      }
      __syncthreads();
      if (c0 + SP_CW < hd) fetch(c0 + SP_CW);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int tile = tid + t * SP_THREADS;
        if (tile < n_tiles) {
          const int qt = tile / k_tiles, kt = tile % k_tiles;
#pragma unroll 8
          for (int c = 0; c < SP_CW; ++c) {
            const float2 q2 = *reinterpret_cast<const float2*>(&qsT[c][2 * qt]);
            const float4 k4 = *reinterpret_cast<const float4*>(&ksT[c][4 * kt]);
            acc[t][0] = fmaf(q2.x, k4.x, acc[t][0]);
            acc[t][1] = fmaf(q2.x, k4.y, acc[t][1]);
            acc[t][2] = fmaf(q2.x, k4.z, acc[t][2]);
            acc[t][3] = fmaf(q2.x, k4.w, acc[t][3]);
            acc[t][4] = fmaf(q2.y, k4.x, acc[t][4]);
            acc[t][5] = fmaf(q2.y, k4.y, acc[t][5]);
            acc[t][6] = fmaf(q2.y, k4.z, acc[t][6]);
            acc[t][7] = fmaf(q2.y, k4.w, acc[t][7]);
          }
        }
      }
    }
    const uint8_t* mrow = mask + i * len;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int tile = tid + t * SP_THREADS;
      if (tile < n_tiles) {
        const int qt = tile / k_tiles, kt = tile % k_tiles;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int q = 2 * qt + (r >> 2), k = 4 * kt + (r & 3);
          if (q < len && k < len) sc[q][k] = mrow[k] ? -INFINITY : acc[t][r];
        }
      }
    }
    __syncthreads();
    float* Pout = P + item * n_pairs;
    for (int j = warp; j < len; j += SP_THREADS / 32) {
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
      for (int c = lane; c < len; c += 32) {
        const float p = sc[j][c] / sum;
        Pout[j * len + c] = p;
        sc[j][c] = seq_keep(seed, 3u, (uint32_t)(item * n_pairs + j * len + c), p_drop) ? p * inv_keep : 0.f;
      }
    }
    __syncthreads();
    if (tid < len) {
      float s = 0.f;
      for (int j = 0; j < len; ++j) s += sc[j][tid];
      s = s / (float)len;
      pbar[tid] = s;
      pbar_out[item * len + tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int j = 0; j < len; ++j) s += pbar[j];
      psum_out[item] = s;
    }
    const float* xrow = x + i * len * (int64_t)dm;
    float* orow = xbar + (i * n_head + h) * (int64_t)dm;
    if ((dm & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(xbar) & 15) == 0) {
      // four columns per thread, 16-byte loads, eight token rows in flight (was: one column, 160 dependent rounds)
      for (int c4 = tid; c4 < (dm >> 2); c4 += SP_THREADS) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int j = 0; j < len; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(xrow + (int64_t)j * dm + 4 * c4);
          const float w = pbar[j];
          a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
        }
        *reinterpret_cast<float4*>(orow + 4 * c4) = a;
      }
    } else {
      for (int c = tid; c < dm; c += SP_THREADS) {
        float s = 0.f;
#pragma unroll 8
        for (int j = 0; j < len; ++j) s = fmaf(pbar[j], xrow[(int64_t)j * dm + c], s);
        orow[c] = s;
      }
    }
  }
}

extern "C" int tiger_train_seq_pool(const float* qk, int64_t ld_qk, const float* x, const uint8_t* mask,
                                    const int32_t* count, int64_t n, int len, int d_model, int n_head, float p_drop,
                                    int seed, float* P, float* pbar, float* psum, float* xbar, void* stream) {
  if (qk == nullptr || x == nullptr || mask == nullptr || P == nullptr || pbar == nullptr || psum == nullptr ||
      xbar == nullptr || n < 0 || len <= 0 || len > SP_MAXL || d_model <= 0 || n_head <= 0 || d_model % n_head != 0 ||
      ld_qk < 2 * (int64_t)d_model || p_drop < 0.f || p_drop >= 1.f)
    return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  int64_t grid = n * n_head;
  if (grid > 148 * 6) grid = 148 * 6;
  train_seq_pool_kernel<<<(unsigned)grid, SP_THREADS, 0, as_stream(stream)>>>(
      qk, ld_qk, x, mask, count, n, len, d_model, n_head, p_drop, (uint32_t)seed, P, pbar, psum, xbar);
  return tiger_launch_status();
}

// ------------------------------------------------------------------------------------------
// backward, two kernels.
//   value path (element-wise, one pass over the token gradients):   dX_j = sum_h pbar_hj dxbar_h
//   score path, one CTA per (node, head):
//     dpbar_j = dxbar_h . x_j + dpsum_h,  dP'_ij = dpbar_j / L,  dP_ij = dP'_ij keep_ij / (1 - p),
//     dS_ij = P_ij (dP_ij - sum_l P_il dP_il),  dQ_i = scale sum_j dS_ij K_j,  dK_j = scale sum_i dS_ij Q_i
//   the two L x L x head_dim products run out of shared memory in column chunks of SPB_CW (Q / K columns staged once,
//   dS broadcast), so nothing is indexed dynamically in registers (the first version kept 2 x 64 floats per thread
//   in local memory and ran at 12 % occupancy: 756 us for 281 nodes; profiles/r02_train_kernels.md)
// ------------------------------------------------------------------------------------------
__global__ void train_seq_pool_bwd_value_kernel(const float* __restrict__ dxbar, const float* __restrict__ pbar,
                                                const int32_t* __restrict__ count, int64_t n_cap, int len, int dm,
                                                int n_head, float* __restrict__ dX) {
  const int64_t n = seq_rows(count, n_cap);
  const int64_t total = n * len * dm;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % dm);
    const int64_t r = e / dm;
    const int64_t i = r / len;
    const int j = (int)(r % len);
    float s = 0.f;
    for (int h = 0; h < n_head; ++h)
      s = fmaf(pbar[(i * n_head + h) * len + j], dxbar[(i * n_head + h) * (int64_t)dm + c], s);
    dX[e] = s;
  }
}

#define SPB_CW 32
// shared memory (dynamic): dS [L][LP] and its transpose [L][LP] (LP = L rounded up to 4, + 4), q / k column chunks
// [L][SPB_CW].  Every thread owns a 4 x 2 block of (row, column) outputs of BOTH products: per reduction step two
// 16-byte loads of dS / dS^T and two 8-byte loads of k / q feed 16 FMAs.
__global__ void __launch_bounds__(SP_THREADS)
train_seq_pool_bwd_kernel(const float* __restrict__ dxbar, const float* __restrict__ dpsum, const float* __restrict__ x,
                          const float* __restrict__ qk, int64_t ld_qk, const float* __restrict__ P,
                          const int32_t* __restrict__ count, int64_t n_cap, int len, int dm, int n_head, float p_drop,
                          uint32_t seed, float* __restrict__ dqk) {
  seed = tiger_step_seed(seed);
  extern __shared__ __align__(16) float spb_smem[];
  const int LP = ((len + 3) & ~3) + 4;
  float* ds = spb_smem;                    // ds[i * LP + r]  = dS[i][r]
  float* dsT = ds + len * LP;              // dsT[j * LP + r] = dS[r][j]
  float* qs = dsT + len * LP;              // qs[j * SPB_CW + c]
  float* ks = qs + len * SPB_CW;
  float* dpb = ks + len * SPB_CW;          // [len]
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id_in_block();
  const int hd = dm / n_head;
  const float scale = sqrtf(1.0f / (float)hd);
  const float inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const int n_pairs = len * len;
  const int r_tiles = (len + 3) / 4, c_tiles = SPB_CW / 2;      // <= 16 x 16 thread tiles
  const int64_t total = seq_rows(count, n_cap) * n_head;
  for (int e = tid; e < 2 * len * LP; e += SP_THREADS) ds[e] = 0.f;      // padding columns stay zero
  for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const int64_t i = item / n_head;
    const int h = (int)(item % n_head);
    const float* xrow = x + i * len * (int64_t)dm;
    const float* gx = dxbar + item * (int64_t)dm;
    __syncthreads();
    for (int j = warp; j < len; j += SP_THREADS / 32) {
      float s = 0.f;
      for (int c = lane; c < dm; c += 32) s = fmaf(gx[c], xrow[(int64_t)j * dm + c], s);
      s = warp_sum(s);
      if (lane == 0) dpb[j] = (s + dpsum[item]) / (float)len;
    }
    __syncthreads();
    const float* Pin = P + item * n_pairs;
    for (int q = warp; q < len; q += SP_THREADS / 32) {          // one warp per query row
      float dot = 0.f;
      for (int c = lane; c < len; c += 32) {
        const bool keep = seq_keep(seed, 3u, (uint32_t)(item * n_pairs + q * len + c), p_drop);
        const float dp = keep ? dpb[c] * inv_keep : 0.f;
        ds[q * LP + c] = dp;
        dot = fmaf(Pin[q * len + c], dp, dot);
      }
      dot = warp_sum(dot);
      for (int c = lane; c < len; c += 32) {
        const float v = Pin[q * len + c] * (ds[q * LP + c] - dot) * scale;
        ds[q * LP + c] = v;
        dsT[c * LP + q] = v;
      }
    }
    const float* qbase = qk + i * len * ld_qk + h * hd;
    const float* kbase = qbase + dm;
    float* dqbase = dqk + i * len * ld_qk + h * hd;
    float* dkbase = dqbase + dm;
    for (int c0 = 0; c0 < hd; c0 += SPB_CW) {
      const int cw = (hd - c0) < SPB_CW ? (hd - c0) : SPB_CW;
      __syncthreads();                                   // dS complete / previous chunk consumed
      for (int e = tid; e < len * SPB_CW; e += SP_THREADS) {
        const int j = e / SPB_CW, c = e % SPB_CW;
        qs[e] = c < cw ? qbase[(int64_t)j * ld_qk + c0 + c] : 0.f;
        ks[e] = c < cw ? kbase[(int64_t)j * ld_qk + c0 + c] : 0.f;
      }
      __syncthreads();
      if (tid < r_tiles * c_tiles) {
        const int rt = tid / c_tiles, ct = tid % c_tiles;
        float aq[8], ak[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) aq[e] = ak[e] = 0.f;
        for (int j = 0; j < len; ++j) {
          const float4 sT = *reinterpret_cast<const float4*>(dsT + j * LP + 4 * rt);   // dS[r0..r0+3][j]
          const float4 sR = *reinterpret_cast<const float4*>(ds + j * LP + 4 * rt);    // dS[j][r0..r0+3]
          const float2 k2 = *reinterpret_cast<const float2*>(ks + j * SPB_CW + 2 * ct);
          const float2 q2 = *reinterpret_cast<const float2*>(qs + j * SPB_CW + 2 * ct);
          aq[0] = fmaf(sT.x, k2.x, aq[0]); aq[1] = fmaf(sT.x, k2.y, aq[1]);
          aq[2] = fmaf(sT.y, k2.x, aq[2]); aq[3] = fmaf(sT.y, k2.y, aq[3]);
          aq[4] = fmaf(sT.z, k2.x, aq[4]); aq[5] = fmaf(sT.z, k2.y, aq[5]);
          aq[6] = fmaf(sT.w, k2.x, aq[6]); aq[7] = fmaf(sT.w, k2.y, aq[7]);
          ak[0] = fmaf(sR.x, q2.x, ak[0]); ak[1] = fmaf(sR.x, q2.y, ak[1]);
          ak[2] = fmaf(sR.y, q2.x, ak[2]); ak[3] = fmaf(sR.y, q2.y, ak[3]);
          ak[4] = fmaf(sR.z, q2.x, ak[4]); ak[5] = fmaf(sR.z, q2.y, ak[5]);
          ak[6] = fmaf(sR.w, q2.x, ak[6]); ak[7] = fmaf(sR.w, q2.y, ak[7]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int r = 4 * rt + (e >> 1), c = 2 * ct + (e & 1);
          if (r < len && c < cw) {
            dqbase[(int64_t)r * ld_qk + c0 + c] = aq[e];       // dQ_r = sum_j dS_rj K_j
            dkbase[(int64_t)r * ld_qk + c0 + c] = ak[e];       // dK_r = sum_i dS_ir Q_i
          }
        }
      }
    }
  }
}

extern "C" int tiger_train_seq_pool_bwd(const float* dxbar, const float* dpsum, const float* x, const float* qk,
                                        int64_t ld_qk, const float* P, const float* pbar, const int32_t* count,
                                        int64_t n, int len, int d_model, int n_head, float p_drop, int seed, float* dX,
                                        float* dqk, void* stream) {
  if (dxbar == nullptr || dpsum == nullptr || x == nullptr || qk == nullptr || P == nullptr || pbar == nullptr ||
      dX == nullptr || dqk == nullptr || n < 0 || len <= 0 || len > SP_MAXL || d_model <= 0 || n_head <= 0 ||
      d_model % n_head != 0 || ld_qk < 2 * (int64_t)d_model)
    return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  cudaStream_t st = as_stream(stream);
  int64_t gv = (n * len * d_model + 255) / 256;
  if (gv > 148 * 16) gv = 148 * 16;
  train_seq_pool_bwd_value_kernel<<<(unsigned)gv, 256, 0, st>>>(dxbar, pbar, count, n, len, d_model, n_head, dX);
  int64_t grid = n * n_head;
  if (grid > 148 * 6) grid = 148 * 6;
  const int lp = ((len + 3) & ~3) + 4;
  const size_t smem = (size_t)(2 * len * lp + 2 * len * SPB_CW + len) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(train_seq_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024) != cudaSuccess)
      return TIGER_ECUDA;
    attr_set = true;
  }
  train_seq_pool_bwd_kernel<<<(unsigned)grid, SP_THREADS, smem, st>>>(dxbar, dpsum, x, qk, ld_qk, P, count, n, len,
                                                                      d_model, n_head, p_drop, (uint32_t)seed, dqk);
  return tiger_launch_status();
}

// ------------------------------------------------------------------------------------------
// value-projection bias under dropout: att[r, h*hd + c] += psum[r, h] * b_v[h*hd + c]; backward:
// g_bv[h*hd + c] += sum_r psum[r, h] datt[r, h*hd + c],  dpsum[r, h] = sum_c datt[r, h*hd + c] b_v[h*hd + c]
// ------------------------------------------------------------------------------------------
__global__ void train_seq_vbias_kernel(float* __restrict__ att, const float* __restrict__ psum,
                                       const float* __restrict__ bv, const int32_t* __restrict__ count, int64_t n,
                                       int dm, int n_head) {
  const int hd = dm / n_head;
  const int64_t total = seq_rows(count, n) * dm;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dm;
    const int c = (int)(i % dm);
    att[i] += psum[r * n_head + c / hd] * bv[c];
  }
}

extern "C" int tiger_train_seq_vbias(float* att, const float* psum, const float* bv, const int32_t* count, int64_t n,
                                     int d_model, int n_head, void* stream) {
  if (att == nullptr || psum == nullptr || bv == nullptr || n < 0 || d_model <= 0 || n_head <= 0) return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  int64_t grid = (n * d_model + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  train_seq_vbias_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(att, psum, bv, count, n, d_model, n_head);
  return tiger_launch_status();
}

__global__ void __launch_bounds__(256)
train_seq_vbias_bwd_kernel(const float* __restrict__ datt, const float* __restrict__ psum, const float* __restrict__ bv,
                           const int32_t* __restrict__ count, int64_t n_cap, int dm, int n_head,
                           float* __restrict__ g_bv, float* __restrict__ dpsum) {
  const int64_t n = seq_rows(count, n_cap);
  // one warp per (row, head) for dpsum; the bias gradient is accumulated by a second pass of column owners
  const int hd = dm / n_head;
  const int lane = lane_id();
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id_in_block(); item < n * n_head; item += n_warps) {
    const int64_t r = item / n_head;
    const int h = (int)(item % n_head);
    const float* g = datt + r * dm + h * hd;
    const float* b = bv + h * hd;
    float s = 0.f;
    for (int c = lane; c < hd; c += 32) s = fmaf(g[c], b[c], s);
    s = warp_sum(s);
    if (lane == 0) dpsum[item] = s;
  }
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < dm; c += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(c / hd);
    float s = 0.f;
    for (int64_t r = 0; r < n; ++r) s = fmaf(psum[r * n_head + h], datt[r * dm + c], s);
    atomicAdd(g_bv + c, s);
  }
}

extern "C" int tiger_train_seq_vbias_bwd(const float* datt, const float* psum, const float* bv, const int32_t* count,
                                         int64_t n, int d_model, int n_head, float* g_bv, float* dpsum, void* stream) {
  if (datt == nullptr || psum == nullptr || bv == nullptr || g_bv == nullptr || dpsum == nullptr || n < 0 ||
      d_model <= 0 || n_head <= 0)
    return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  int64_t grid = (n * n_head + 7) / 8;
  const int64_t need = (d_model + 255) / 256;
  if (grid < need) grid = need;
  if (grid > 148 * 4) grid = 148 * 4;
  train_seq_vbias_bwd_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(datt, psum, bv, count, n, d_model, n_head,
                                                                           g_bv, dpsum);
  return tiger_launch_status();
}

// ------------------------------------------------------------------------------------------
// token gradients (restarters.py:96-104): columns [2d, 3d) are rows of anony_emb, columns [3d + de, 4d + de) the
// time code cos((t_last - t_j) w + b); the last position keeps only its time code.  One warp per token.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
train_seq_tokens_bwd_kernel(const float* __restrict__ dX, const int32_t* __restrict__ count, int64_t n, int len,
                            const int64_t* __restrict__ anony_ids,
                            const float* __restrict__ hist_ts, int d, int de, const float* __restrict__ time_w,
                            const float* __restrict__ time_b, float* __restrict__ g_anony, float* __restrict__ g_w,
                            float* __restrict__ g_b) {
  extern __shared__ float acc_smem[];          // [2][d]
  float* s_w = acc_smem;
  float* s_b = acc_smem + d;
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) acc_smem[c] = 0.f;
  __syncthreads();
  const int lane = lane_id();
  const int64_t dm = 4 * (int64_t)d + de;
  const int64_t total = seq_rows(count, n) * len;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id_in_block(); r < total; r += n_warps) {
    const int64_t i = r / len;
    const int j = (int)(r % len);
    const float* g = dX + r * dm;
    if (j != len - 1) {
      float* ga = g_anony + anony_ids[r] * d;
      for (int c = lane; c < d; c += 32) {
        const float v = g[2 * d + c];
        if (v != 0.f) atomicAdd(ga + c, v);
      }
    }
    const float dt = hist_ts[i * len + len - 1] - hist_ts[r];
    for (int c = lane; c < d; c += 32) {
      const float v = -seq_sin_reduced(__fadd_rn(__fmul_rn(dt, time_w[c]), time_b[c])) * g[3 * d + de + c];
      atomicAdd(s_w + c, v * dt);
      atomicAdd(s_b + c, v);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    if (s_w[c] != 0.f) atomicAdd(g_w + c, s_w[c]);
    if (s_b[c] != 0.f) atomicAdd(g_b + c, s_b[c]);
  }
}

extern "C" int tiger_train_seq_tokens_bwd(const float* dX, const int32_t* count, int64_t n, int len,
                                          const int64_t* anony_ids,
                                          const float* hist_ts, int d, int de, const float* time_w, const float* time_b,
                                          float* g_anony_emb, float* g_time_w, float* g_time_b, void* stream) {
  if (dX == nullptr || anony_ids == nullptr || hist_ts == nullptr || time_w == nullptr || time_b == nullptr ||
      g_anony_emb == nullptr || g_time_w == nullptr || g_time_b == nullptr || n < 0 || len <= 0 || d <= 0 || de <= 0)
    return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  int64_t grid = (n * len + 7) / 8;
  if (grid > 148 * 2) grid = 148 * 2;
  train_seq_tokens_bwd_kernel<<<(unsigned)grid, 256, 2 * d * sizeof(float), as_stream(stream)>>>(
      dX, count, n, len, anony_ids, hist_ts, d, de, time_w, time_b, g_anony_emb, g_time_w, g_time_b);
  return tiger_launch_status();
}

// x[i] = keep(i) ? x[i] / (1 - p) : 0 in place (nn.Dropout in training mode; MergeLayer, basic_modules.py:16-19)
__global__ void train_dropout_kernel(float* __restrict__ x, const int32_t* __restrict__ count, int64_t per,
                                     int64_t n_cap, float p, uint32_t seed, uint32_t stream_id) {
  seed = tiger_step_seed(seed);
  const float inv_keep = 1.0f / (1.0f - p);
  const int64_t n = seq_rows(count, n_cap, per);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = seq_keep(seed, stream_id, (uint32_t)i, p) ? x[i] * inv_keep : 0.f;
}

extern "C" int tiger_train_dropout(float* x, const int32_t* count, int64_t per_count, int64_t n, float p_drop, int seed,
                                   int stream_id, void* stream) {
  if (x == nullptr || n < 0 || p_drop < 0.f || p_drop >= 1.f) return TIGER_EINVAL;
  if (n == 0 || p_drop == 0.f) return TIGER_OK;
  int64_t grid = (n + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  train_dropout_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(x, count, per_count > 0 ? per_count : 1, n, p_drop,
                                                                     (uint32_t)seed, (uint32_t)stream_id);
  return tiger_launch_status();
}

// y[i] += alpha * x[i]
__global__ void train_axpy_kernel(float* __restrict__ y, const float* __restrict__ x, const int32_t* __restrict__ count,
                                  int64_t per, int64_t n_cap, float alpha) {
  const int64_t n = seq_rows(count, n_cap, per);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

extern "C" int tiger_train_axpy(float* y, const float* x, const int32_t* count, int64_t per_count, int64_t n, float alpha,
                                void* stream) {
  if (y == nullptr || x == nullptr || n < 0) return TIGER_EINVAL;
  if (n == 0) return TIGER_OK;
  int64_t grid = (n + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  train_axpy_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(y, x, count, per_count > 0 ? per_count : 1, n, alpha);
  return tiger_launch_status();
}

int tiger_seed_step_set_train_seq(const int32_t* p) { return tiger_seed_step_set_here(p); }
