// Large fp32-accurate products of the training step on the tensor cores, both operands PRE-SPLIT and PRE-TILED:
//     C[m, n] (+)= act(alpha * (sum_k opA[m, k] * opW[n, k] + bias[n]))            (tf32x3, see umma.cuh)
//
// Why a second GEMM: the staging kernel of gemm.cu converts its operands while it feeds the tensor core - every
// A tile is loaded, split into tf32 head / tail and written to shared memory once per N tile it meets, every W
// tile once per M tile.  ncu on the training step (profiles/r02_train_kernels.md): the q/k in-projection of the seq
// restarter (11,240 x 1,720 x 860) and its two gradient products ran 390 / 577 / 882 us with the tensor pipe 11-26 %
// active and the issue slots 22-44 % busy - the 16 producer warps are the bottleneck, the more so for the
// transposed operands of the gradients (scalar loads).  Here the conversion happens ONCE per operand:
//   tiger_gemm_pp_pack   one pass over the operand (row-major or transposed, rows / reduction length optionally
//                        bounded by device-side counts): tf32 head / tail planes written as the exact shared-memory
//                        image of every pipeline stage,  pack[tile][k-block][head | tail][kc][row 0..127][4 floats]
//   tiger_sgemm_pp       no producer warps at all: one thread issues two TMA bulk copies per stage (16 KB per operand),
//                        three warps issue the tf32x3 MMA streams into four TMEM accumulators, four warps run the
//                        epilogue (bias / alpha / ReLU, or atomic accumulation for weight gradients with K split
//                        over blockIdx.y); persistent over output tiles.
// Pack traffic is O(operand); what it removes is O(operand x tiles of the other operand).
#include <cstring>

#include "common.cuh"
#include "umma.cuh"

#define PP_BM 128
#define PP_BN 128
#define PP_STAGE_FLOATS (2 * UMMA_KCH * PP_BM * 4)      // one operand, one stage: head + tail planes (16 KB)
#define PP_MAX_STAGES 6
#define PP_THREADS 256                                  // warp 0 TMA, warps 1-3 MMA issuers, warps 4-7 epilogue

// ------------------------------------------------------------------------------------------
// pack
// ------------------------------------------------------------------------------------------
__global__ void gemm_pp_pack_kernel(const float* __restrict__ src, int64_t ld, int trans, int64_t rows, int64_t k_dim,
                                    const int32_t* __restrict__ row_count, const int32_t* __restrict__ k_count,
                                    int64_t per_count, int64_t n_kb, float* __restrict__ out) {
  int64_t rows_eff = rows, k_eff = k_dim;
  if (row_count != nullptr) {
    const int64_t c = (int64_t)(*row_count) * per_count;
    rows_eff = c < rows ? c : rows;
  }
  if (k_count != nullptr) {
    const int64_t c = (int64_t)(*k_count) * per_count;
    k_eff = c < k_dim ? c : k_dim;
  }
  // only the tiles / k-blocks inside the device-side bounds are written: tiger_sgemm_pp applies the same bounds, and a
  // launch sized for a capacity (6,600 restart rows x 40 tokens) must not sweep it (155 us at first)
  const int64_t tiles = (rows_eff + PP_BM - 1) / PP_BM;
  const int64_t kb_eff = (k_eff + UMMA_BK - 1) / UMMA_BK;
  const int64_t total = tiles * kb_eff * UMMA_KCH * PP_BM;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % PP_BM);
    const int kc = (int)((i / PP_BM) % UMMA_KCH);
    const int64_t kb = (i / (PP_BM * UMMA_KCH)) % kb_eff;
    const int64_t t = i / (PP_BM * UMMA_KCH * kb_eff);
    const int64_t row = t * PP_BM + r;
    const int64_t k = kb * UMMA_BK + kc * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows_eff && k < k_eff) {
      if (!trans) {
        const float* p = src + row * ld + k;
        if (k + 4 <= k_eff && ((((uintptr_t)p) & 15) == 0)) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          v.x = __ldg(p);
          if (k + 1 < k_eff) v.y = __ldg(p + 1);
          if (k + 2 < k_eff) v.z = __ldg(p + 2);
          if (k + 3 < k_eff) v.w = __ldg(p + 3);
        }
      } else {                                   // element (row, k) at src[k * ld + row]: coalesced over rows
        const float* p = src + k * ld + row;
        v.x = __ldg(p);
        if (k + 1 < k_eff) v.y = __ldg(p + ld);
        if (k + 2 < k_eff) v.z = __ldg(p + 2 * ld);
        if (k + 3 < k_eff) v.w = __ldg(p + 3 * ld);
      }
    }
    float4 h, l;
    tf32_split(v, h, l);
    float* stage = out + (t * n_kb + kb) * PP_STAGE_FLOATS;
    *reinterpret_cast<float4*>(stage + (kc * PP_BM + r) * 4) = h;
    *reinterpret_cast<float4*>(stage + UMMA_KCH * PP_BM * 4 + (kc * PP_BM + r) * 4) = l;
  }
}

extern "C" int64_t tiger_gemm_pp_pack_bytes(int64_t rows, int64_t k_dim) {
  if (rows <= 0 || k_dim <= 0) return -1;
  const int64_t tiles = (rows + PP_BM - 1) / PP_BM, n_kb = (k_dim + UMMA_BK - 1) / UMMA_BK;
  return tiles * n_kb * PP_STAGE_FLOATS * (int64_t)sizeof(float);
}

extern "C" int tiger_gemm_pp_pack(const float* src, int64_t ld, int trans, int64_t rows, int64_t k_dim,
                                  const int32_t* row_count, const int32_t* k_count, int64_t per_count, float* out,
                                  void* stream) {
  if (src == nullptr || out == nullptr || rows <= 0 || k_dim <= 0 || (!trans && ld < k_dim) || (trans && ld < rows) ||
      (((uintptr_t)out) & 15) != 0)
    return TIGER_EINVAL;
  const int64_t tiles = (rows + PP_BM - 1) / PP_BM, n_kb = (k_dim + UMMA_BK - 1) / UMMA_BK;
  const int64_t total = tiles * n_kb * UMMA_KCH * PP_BM;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  gemm_pp_pack_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(src, ld, trans, rows, k_dim, row_count, k_count,
                                                                    per_count > 0 ? per_count : 1, n_kb, out);
  return tiger_launch_status();
}

// ------------------------------------------------------------------------------------------
// product
// ------------------------------------------------------------------------------------------
struct PPArgs {
  const float* apack;
  const float* wpack;
  const float* bias;
  float* C;
  int64_t ldc, M;
  int N;
  int64_t n_kb_total;          // k-blocks of the packs
  const int32_t* m_count;
  const int32_t* k_count;
  int64_t per_count;
  float alpha;
  int relu, accumulate, kblk_per_part, stages, vec_c;
};

__global__ void __launch_bounds__(PP_THREADS, 1) gemm_pp_kernel(const PPArgs g) {
  extern __shared__ __align__(128) unsigned char pp_smem[];
  const int S = g.stages;
  float* stage0 = reinterpret_cast<float*>(pp_smem);
  const int stage_floats = 2 * PP_STAGE_FLOATS;                 // A then W
  uint64_t* full = reinterpret_cast<uint64_t*>(pp_smem + (size_t)S * stage_floats * sizeof(float));
  uint64_t* empty = full + PP_MAX_STAGES;
  uint64_t* done = empty + PP_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  int64_t M = g.M;
  if (g.m_count != nullptr) {
    const int64_t c = (int64_t)(*g.m_count) * g.per_count;
    M = c < M ? c : M;
  }
  int64_t kb_all = g.n_kb_total;
  if (g.k_count != nullptr) {
    const int64_t c = ((int64_t)(*g.k_count) * g.per_count + UMMA_BK - 1) / UMMA_BK;
    kb_all = c < kb_all ? c : kb_all;
  }
  const int64_t kblk0 = g.kblk_per_part > 0 ? (int64_t)blockIdx.y * g.kblk_per_part : 0;
  if (kblk0 >= kb_all && g.accumulate) return;                  // nothing to add
  int64_t n_blocks = kb_all - kblk0;
  if (g.kblk_per_part > 0 && n_blocks > g.kblk_per_part) n_blocks = g.kblk_per_part;
  const int tiles_n = (g.N + PP_BN - 1) / PP_BN;
  const int64_t tiles_m = (g.M + PP_BM - 1) / PP_BM;
  const int64_t n_tiles = tiles_m * tiles_n;
  if ((int64_t)(blockIdx.x / tiles_n) * PP_BM >= M) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TMEM_COLS = 512;
  auto init_barriers = [&](bool again) {
    if (tid == 0) {
      for (int s = 0; s < S; ++s) {
        if (again) { mbar_inval(full + s); mbar_inval(empty + s); }
        mbar_init(full + s, 1);
        mbar_init(empty + s, UMMA_ISSUERS);
      }
      if (again) mbar_inval(done);
      mbar_init(done, UMMA_ISSUERS);
      fence_mbar_init();
    }
  };
  init_barriers(false);
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = *tmem_slot;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tm = tile / tiles_n;
    const int tn = (int)(tile % tiles_n);
    const int64_t m0 = tm * PP_BM;
    if (m0 >= M) break;
    const int n0 = tn * PP_BN;
    if (tile != (int64_t)blockIdx.x) {
      init_barriers(true);
      __syncthreads();
    }
    if (warp == 0) {
      // ---------------- TMA: two bulk copies per stage ----------------
      if (lane == 0) {
        const float* a_src = g.apack + (tm * g.n_kb_total + kblk0) * PP_STAGE_FLOATS;
        const float* w_src = g.wpack + ((int64_t)tn * g.n_kb_total + kblk0) * PP_STAGE_FLOATS;
        const uint32_t bytes = (uint32_t)PP_STAGE_FLOATS * 4u;
        for (int64_t blk = 0; blk < n_blocks; ++blk) {
          const int s = (int)(blk % S);
          mbar_wait(empty + s, (uint32_t)(((blk / S) & 1) ^ 1));
          mbar_arrive_expect_tx(full + s, 2 * bytes);
          float* dst = stage0 + (size_t)s * stage_floats;
          tma_bulk_load(dst, a_src + blk * PP_STAGE_FLOATS, bytes, full + s);
          tma_bulk_load(dst + PP_STAGE_FLOATS, w_src + blk * PP_STAGE_FLOATS, bytes, full + s);
        }
      }
    } else if (warp < 1 + UMMA_ISSUERS) {
      // ---------------- MMA issuers (roles of umma.cuh) ----------------
      const int role = uniform_warp_idx() - 1;
      const UmmaRole r = umma_role(role, smem_addr_u32(stage0), (uint32_t)stage_floats * 4u, PP_BM, PP_BN, (uint32_t)PP_BN);
      const uint32_t idesc = umma_idesc_tf32(PP_BM, PP_BN);
      const uint32_t tbase = __shfl_sync(0xffffffffu, taddr, 0);
      const uint32_t d_even = tbase + r.acc_even, d_odd = tbase + r.acc_odd;
      int s = 0;
      uint32_t ph = 0, a = r.a_lo, b = r.b_lo;
      for (int64_t blk = 0; blk < n_blocks; ++blk) {
        mbar_wait(full + s, ph);
        tc_fence_after_sync();
        if (elect_one()) {
          umma_tf32_lo(d_even, a, b, idesc, blk > 0 ? 1u : 0u);
          umma_tf32_lo(d_odd, a + r.a_kstep, b + r.b_kstep, idesc, (role == 2 && blk == 0) ? 0u : 1u);
          umma_commit(empty + s);
        }
        __syncwarp();
        a += r.stage_step;
        b += r.stage_step;
        if (++s == S) {
          s = 0;
          ph ^= 1;
          a = r.a_lo;
          b = r.b_lo;
        }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    } else {
      // ---------------- epilogue: warp q reads its 32 TMEM lanes (rows 32 q .. 32 q + 31) ----------------
      mbar_wait(done, 0);
      tc_fence_after_sync();
      const int q = warp & 3;
      const int64_t m = m0 + q * 32 + lane;
      const bool row_ok = m < M;
      const uint32_t tl = taddr + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < PP_BN; c0 += 16) {
        float v[16];
        tmem_ld16(tl + (uint32_t)c0, v);
#pragma unroll
        for (int j = 1; j < UMMA_ACCS; ++j) {
          float t[16];
          tmem_ld16(tl + (uint32_t)(j * PP_BN + c0), t);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += t[e];
        }
        const int nb = n0 + c0;
        if (!row_ok || nb >= g.N) continue;
        float* dst = g.C + m * g.ldc + nb;
        if (g.accumulate) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nb + j < g.N) atomicAdd(dst + j, v[j] * g.alpha);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = nb + j;
            float x = (v[j] + ((g.bias != nullptr && n < g.N) ? __ldg(g.bias + n) : 0.f)) * g.alpha;
            v[j] = g.relu ? fmaxf(x, 0.f) : x;
          }
          if (g.vec_c && nb + 16 <= g.N) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (nb + j < g.N) dst[j] = v[j];
          }
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  if (warp == 4) tmem_dealloc(taddr, TMEM_COLS);
}

// apack / wpack: tiger_gemm_pp_pack images of opA [m_rows, k_dim] and opW [n_cols, k_dim] (same k_dim).
// accumulate != 0: C += alpha * product with atomic adds, K split over k_parts CTAs per tile (bias / relu off).
extern "C" int tiger_sgemm_pp(const float* apack, const float* wpack, const float* bias, float* C, int64_t ldc,
                              int64_t m_rows, int n_cols, int64_t k_dim, const int32_t* m_count, const int32_t* k_count,
                              int64_t per_count, float alpha, int relu, int accumulate, int k_parts, void* stream) {
  if (apack == nullptr || wpack == nullptr || C == nullptr || m_rows <= 0 || n_cols <= 0 || k_dim <= 0 || ldc < n_cols ||
      k_parts < 1 || (accumulate && (bias != nullptr || relu)) || (!accumulate && (k_parts > 1 || k_count != nullptr)) ||
      ((((uintptr_t)apack) | ((uintptr_t)wpack)) & 15) != 0)
    return TIGER_EINVAL;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (cudaFuncSetAttribute(gemm_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             PP_MAX_STAGES * 2 * PP_STAGE_FLOATS * 4 + 256) != cudaSuccess)
      return TIGER_ECUDA;
    sms = n > 0 ? n : 148;
  }
  PPArgs g;
  memset(&g, 0, sizeof(g));
  g.apack = apack; g.wpack = wpack; g.bias = bias; g.C = C; g.ldc = ldc; g.M = m_rows; g.N = n_cols;
  g.n_kb_total = (k_dim + UMMA_BK - 1) / UMMA_BK;
  g.m_count = m_count; g.k_count = k_count; g.per_count = per_count > 0 ? per_count : 1;
  g.alpha = alpha; g.relu = relu; g.accumulate = accumulate ? 1 : 0;
  g.stages = PP_MAX_STAGES;
  g.vec_c = ((((uintptr_t)C) & 15) == 0 && (ldc & 3) == 0) ? 1 : 0;
  const int64_t tiles = ((m_rows + PP_BM - 1) / PP_BM) * ((n_cols + PP_BN - 1) / PP_BN);
  int64_t parts = 1;
  if (accumulate) {
    // K split: enough CTAs for about two waves, at most ~2048 reduction steps per accumulator set (the tensor core
    // truncates when it adds into its fp32 accumulator), but no more - every part adds its whole tile with atomics
    // (98 tiles x 63 parts of the seq restarter's weight gradient were 93 M atomic adds: 514 us)
    int64_t want = (2 * (int64_t)sms + tiles - 1) / tiles;
    const int64_t by_len = (k_dim + 2047) / 2048;
    want = want > by_len ? want : by_len;
    want = want < k_parts ? want : k_parts;
    parts = want < g.n_kb_total ? want : g.n_kb_total;
    if (parts < 1) parts = 1;
    g.kblk_per_part = (int)((g.n_kb_total + parts - 1) / parts);
    parts = (g.n_kb_total + g.kblk_per_part - 1) / g.kblk_per_part;
  }
  dim3 grid((unsigned)(tiles < sms ? tiles : sms), (unsigned)parts);
  gemm_pp_kernel<<<grid, PP_THREADS, (size_t)g.stages * 2 * PP_STAGE_FLOATS * 4 + 256, as_stream(stream)>>>(g);
  return tiger_launch_status();
}
