// fp32 "NT" GEMM with fused epilogue:
//     C[b][m, n] = act(alpha * (sum_k A[b][m, k] * W[b][n, k] + bias[b][n])),  rows with row_zero[m] != 0 -> 0
// A row-major [M, K] (lda), W row-major [N, K] (ldw) - the layout of nn.Linear / in_proj weights, so
// parameters are used as stored (reference: nn.MultiheadAttention in/out projections, nn.Linear and
// MergeLayer of tiger/model/restarters.py:45-50, temporal_agg_modules.py:203-209, basic_modules.py:5-19).
//
// FFMA, not tensor cores: the parity bar is fp32 max-norm 1e-5 against the CPU reference, which
// single-pass TF32 (10-bit mantissa) cannot meet.
//
// Structure: persistent CTAs (grid = a multiple of the SM count) walk the tiles of the ACTUAL
// problem - the row count may live on the device (`count` * rows_per_count), which keeps the restart
// path free of host syncs.  Tiles are 128x128 (8x8 per thread) when that still fills the GPU and
// 32x64 (2x4 per thread) otherwise, chosen at run time.  Operands stream global -> shared with
// cp.async (16-byte LDGSTS, zero-fill for edges) through a 3-stage ring; both operands keep their
// natural [row][k] layout in shared memory (row stride 20 floats: conflict-free LDS.128), and each
// thread owns rows / columns strided by 16 so that warp-wide stores to C are coalesced.
#include "common.cuh"

#define GEMM_THREADS 256
#define GEMM_BK 16
#define GEMM_LDS (GEMM_BK + 4)
#define GEMM_STAGES 3

template <int BM, int BN>
struct FfmaSmem {
  float a[GEMM_STAGES][BM][GEMM_LDS];
  float w[GEMM_STAGES][BN][GEMM_LDS];
};

union FfmaSmemAll {
  FfmaSmem<128, 128> big;
  FfmaSmem<32, 64> small;
};

struct FfmaArgs {
  const float* A;
  const float* W;
  const float* bias;
  float* C;
  const uint8_t* row_zero;
  const int32_t* count;
  int64_t lda, ldw, ldc;
  int64_t stride_a, stride_w, stride_bias, stride_c;
  int64_t M, rows_per_count;
  int batch, N, K;
  float alpha;
  int relu;
  int vec_ok;  // A, W 16-byte aligned with lda, ldw, strides multiples of 4 floats
};

__device__ __forceinline__ void ffma_cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ffma_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ffma_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// one BK slab of a [ROWS x BK] operand tile: global -> shared (rows >= rows_valid and k >= K read as zero)
template <int ROWS>
__device__ __forceinline__ void gemm_load(float (*sm)[GEMM_LDS], const float* __restrict__ base, int64_t ld,
                                          int64_t row0, int64_t rows_valid, int k0, int K, int vec_ok, int tid) {
  constexpr int CHUNKS = ROWS * (GEMM_BK / 4);
#pragma unroll
  for (int c = tid; c < CHUNKS; c += GEMM_THREADS) {
    const int r = c >> 2, kq = (c & 3) << 2;
    const int k = k0 + kq;
    const bool row_ok = row0 + r < rows_valid;
    int valid = row_ok ? (K - k) : 0;            // floats available from k on
    valid = valid < 0 ? 0 : (valid > 4 ? 4 : valid);
    // clamp the address into the allocation even when nothing is read from it
    const float* p = base + (row_ok ? (row0 + r) : row0) * ld + (valid > 0 ? k : 0);
    if (vec_ok) {
      ffma_cp_async16(&sm[r][kq], p, valid * 4);
    } else {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid > 0) v.x = __ldg(p);
      if (valid > 1) v.y = __ldg(p + 1);
      if (valid > 2) v.z = __ldg(p + 2);
      if (valid > 3) v.w = __ldg(p + 3);
      *reinterpret_cast<float4*>(&sm[r][kq]) = v;
    }
  }
}

template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void gemm_tile(const FfmaArgs& g, FfmaSmem<BM, BN>& sm, const float* __restrict__ A,
                                          const float* __restrict__ W, const float* __restrict__ bias,
                                          float* __restrict__ C, int64_t m0, int n0, int64_t M, int tid) {
  constexpr int SX = BN / TN, SY = BM / TM;   // thread grid; thread (ty, tx) owns rows ty + i*SY, cols tx + j*SX
  static_assert(SX * SY == GEMM_THREADS, "thread tiling");
  const int tx = tid % SX, ty = tid / SX;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int n_slabs = (g.K + GEMM_BK - 1) / GEMM_BK;
#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < n_slabs) {
      gemm_load<BM>(sm.a[s], A, g.lda, m0, M, s * GEMM_BK, g.K, g.vec_ok, tid);
      gemm_load<BN>(sm.w[s], W, g.ldw, n0, g.N, s * GEMM_BK, g.K, g.vec_ok, tid);
    }
    ffma_cp_async_commit();
  }
  for (int s = 0; s < n_slabs; ++s) {
    ffma_cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();   // slab s has landed for every thread; the stage refilled below is no longer being read
    const int nxt = s + GEMM_STAGES - 1;
    if (nxt < n_slabs) {
      gemm_load<BM>(sm.a[nxt % GEMM_STAGES], A, g.lda, m0, M, nxt * GEMM_BK, g.K, g.vec_ok, tid);
      gemm_load<BN>(sm.w[nxt % GEMM_STAGES], W, g.ldw, n0, g.N, nxt * GEMM_BK, g.K, g.vec_ok, tid);
    }
    ffma_cp_async_commit();
    const int buf = s % GEMM_STAGES;
#pragma unroll
    for (int kq = 0; kq < GEMM_BK; kq += 4) {
      float4 a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(&sm.a[buf][ty + i * SY][kq]);
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = *reinterpret_cast<const float4*>(&sm.w[buf][tx + j * SX][kq]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
        }
    }
  }
  ffma_cp_async_wait<0>();
  __syncthreads();     // every thread is done with the ring before the next tile's prologue refills it
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty + i * SY;
    if (m >= M) continue;
    const bool zero = g.row_zero != nullptr && g.row_zero[m] != 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + j * SX;
      if (n >= g.N) continue;
      float v = (acc[i][j] + (bias != nullptr ? bias[n] : 0.f)) * g.alpha;
      if (g.relu) v = fmaxf(v, 0.f);
      C[m * g.ldc + n] = zero ? 0.f : v;
    }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 2) sgemm_ffma_kernel(const FfmaArgs g, int sm_count) {
  extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
  FfmaSmemAll& sm = *reinterpret_cast<FfmaSmemAll*>(gemm_smem_raw);
  int64_t M = g.M;
  if (g.count != nullptr) {
    const int64_t c = (int64_t)(*g.count) * g.rows_per_count;
    M = c < M ? c : M;
  }
  if (M <= 0) return;
  const int tid = threadIdx.x;
  const int64_t tm_big = (M + 127) / 128, tn_big = (g.N + 127) / 128;
  if (tm_big * tn_big * g.batch >= (3 * sm_count) / 4) {
    const int64_t per = tm_big * tn_big;
    for (int64_t t = blockIdx.x; t < per * g.batch; t += gridDim.x) {
      const int64_t b = t / per, r = t % per;
      gemm_tile<128, 128, 8, 8>(g, sm.big, g.A + b * g.stride_a, g.W + b * g.stride_w,
                                g.bias != nullptr ? g.bias + b * g.stride_bias : nullptr, g.C + b * g.stride_c,
                                (r / tn_big) * 128, (int)(r % tn_big) * 128, M, tid);
    }
  } else {
    const int64_t tm = (M + 31) / 32, tn = (g.N + 63) / 64;
    const int64_t per = tm * tn;
    for (int64_t t = blockIdx.x; t < per * g.batch; t += gridDim.x) {
      const int64_t b = t / per, r = t % per;
      gemm_tile<32, 64, 2, 4>(g, sm.small, g.A + b * g.stride_a, g.W + b * g.stride_w,
                              g.bias != nullptr ? g.bias + b * g.stride_bias : nullptr, g.C + b * g.stride_c,
                              (r / tn) * 32, (int)(r % tn) * 64, M, tid);
    }
  }
}

static int g_ffma_sms = 0;

extern "C" int tiger_sgemm_ffma_batched(const float* A, int64_t lda, int64_t stride_a, const float* W, int64_t ldw,
                                      int64_t stride_w, const float* bias, int64_t stride_bias, float* C,
                                      int64_t ldc, int64_t stride_c, int batch, int64_t m_rows,
                                      const int32_t* count, int64_t rows_per_count, int n_cols, int k_dim,
                                      float alpha, int relu, const uint8_t* row_zero, void* stream) {
  if (m_rows < 0 || batch <= 0 || n_cols <= 0 || k_dim <= 0 || lda < k_dim || ldw < k_dim || ldc < n_cols)
    return TIGER_EINVAL;
  if (m_rows == 0) return TIGER_OK;
  if (g_ffma_sms == 0) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (cudaFuncSetAttribute(sgemm_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(FfmaSmemAll)) != cudaSuccess)
      return TIGER_ECUDA;
    g_ffma_sms = sms;
  }
  FfmaArgs g;
  g.A = A; g.W = W; g.bias = bias; g.C = C; g.row_zero = row_zero; g.count = count;
  g.lda = lda; g.ldw = ldw; g.ldc = ldc;
  g.stride_a = stride_a; g.stride_w = stride_w; g.stride_bias = stride_bias; g.stride_c = stride_c;
  g.M = m_rows; g.rows_per_count = rows_per_count > 0 ? rows_per_count : 1;
  g.batch = batch; g.N = n_cols; g.K = k_dim; g.alpha = alpha; g.relu = relu;
  const bool strides_ok = batch == 1 || (((stride_a | stride_w) & 3) == 0);
  g.vec_ok = ((((uintptr_t)A | (uintptr_t)W) & 15) == 0 && (lda & 3) == 0 && (ldw & 3) == 0 && strides_ok) ? 1 : 0;
  // enough CTAs for the largest possible problem, never more than two per SM
  const int64_t tiles_small = ((m_rows + 31) / 32) * ((n_cols + 63) / 64) * batch;
  const int64_t grid = tiles_small < 2 * (int64_t)g_ffma_sms ? tiles_small : 2 * (int64_t)g_ffma_sms;
  sgemm_ffma_kernel<<<(unsigned)grid, GEMM_THREADS, sizeof(FfmaSmemAll), as_stream(stream)>>>(g, g_ffma_sms);
  return tiger_launch_status();
}

extern "C" int tiger_sgemm_ffma(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                              float* C, int64_t ldc, int64_t m_rows, const int32_t* count, int64_t rows_per_count,
                              int n_cols, int k_dim, int relu, void* stream) {
  return tiger_sgemm_ffma_batched(A, lda, 0, W, ldw, 0, bias, 0, C, ldc, 0, 1, m_rows, count, rows_per_count, n_cols,
                                k_dim, 1.0f, relu, nullptr, stream);
}
