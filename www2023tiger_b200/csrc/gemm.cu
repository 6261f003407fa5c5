// fp32 "NT" GEMM with fused bias / ReLU:  C[m, n] = act(sum_k A[m, k] * W[n, k] + bias[n])
// A row-major [M, K] (lda), W row-major [N, K] (ldw) - the layout of nn.Linear / in_proj weights,
// so parameters are used as stored (reference: nn.MultiheadAttention in_proj / out_proj,
// nn.Linear and MergeLayer of tiger/model/restarters.py:45-50, basic_modules.py:5-19).
//
// FFMA, not tensor cores: the parity bar is fp32 max-norm 1e-5 against the CPU reference, which
// single-pass TF32 (10-bit mantissa) cannot meet.  The row count may live on the device
// (`count` * rows_per_count) so that the restart path stays free of host syncs: the kernel is
// persistent (grid = a multiple of the SM count) and walks the tiles of the ACTUAL problem,
// choosing a 128x128 tile (8x8 per thread) for tall problems and a 32x64 tile (2x4 per thread)
// for short ones at run time.
#include "common.cuh"

#define GEMM_THREADS 256
#define GEMM_BK 16

template <int BM, int BN>
struct GemmSmem {
  float a[2][GEMM_BK][BM + 4];
  float w[2][GEMM_BK][BN + 4];
};

union GemmSmemAll {
  GemmSmem<128, 128> big;
  GemmSmem<32, 64> small;
};

struct GemmArgs {
  const float* A;
  int64_t lda;
  const float* W;
  int64_t ldw;
  const float* bias;
  float* C;
  int64_t ldc;
  int64_t M;
  const int32_t* count;
  int64_t rows_per_count;
  int N, K;
  int relu;
  int vec_ok;  // A, W 16-byte aligned with lda, ldw, K multiples of 4
};

// global -> registers for one BK slab of a [ROWS x BK] operand tile (rows beyond `rows_valid` and
// k beyond K read as zero)
template <int ROWS, int PER>
__device__ __forceinline__ void gemm_fetch(float4 (&st)[PER], const float* __restrict__ base, int64_t ld,
                                           int64_t row0, int64_t rows_valid, int k0, int K, int vec_ok, int tid) {
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int f = tid + i * GEMM_THREADS;          // float4 slot: row = f / 4, kq = f % 4
    const int r = f >> 2, kq = (f & 3) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < ROWS && row0 + r < rows_valid) {
      const float* p = base + (row0 + r) * ld + k0 + kq;
      if (vec_ok && k0 + kq + 3 < K) {
        v = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        if (k0 + kq + 0 < K) v.x = __ldg(p + 0);
        if (k0 + kq + 1 < K) v.y = __ldg(p + 1);
        if (k0 + kq + 2 < K) v.z = __ldg(p + 2);
        if (k0 + kq + 3 < K) v.w = __ldg(p + 3);
      }
    }
    st[i] = v;
  }
}

template <int ROWS, int PER, int LD>
__device__ __forceinline__ void gemm_stash(const float4 (&st)[PER], float (*sm)[LD], int tid) {
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int f = tid + i * GEMM_THREADS;
    const int r = f >> 2, kq = (f & 3) << 2;
    if (r < ROWS) {
      sm[kq + 0][r] = st[i].x;
      sm[kq + 1][r] = st[i].y;
      sm[kq + 2][r] = st[i].z;
      sm[kq + 3][r] = st[i].w;
    }
  }
}

template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void gemm_tile(const GemmArgs& g, GemmSmem<BM, BN>& sm, int64_t m0, int n0, int64_t M,
                                          int tid) {
  static_assert((BM / TM) * (BN / TN) == GEMM_THREADS, "thread tiling");
  constexpr int PA = (BM * GEMM_BK / 4 + GEMM_THREADS - 1) / GEMM_THREADS;
  constexpr int PW = (BN * GEMM_BK / 4 + GEMM_THREADS - 1) / GEMM_THREADS;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float4 sa[PA], sw[PW];
  const int n_slabs = (g.K + GEMM_BK - 1) / GEMM_BK;
  gemm_fetch<BM, PA>(sa, g.A, g.lda, m0, M, 0, g.K, g.vec_ok, tid);
  gemm_fetch<BN, PW>(sw, g.W, g.ldw, n0, g.N, 0, g.K, g.vec_ok, tid);
  __syncthreads();  // the previous tile has finished reading both buffers
  gemm_stash<BM, PA, BM + 4>(sa, sm.a[0], tid);
  gemm_stash<BN, PW, BN + 4>(sw, sm.w[0], tid);
  __syncthreads();
  for (int s = 0; s < n_slabs; ++s) {
    const int buf = s & 1;
    const bool more = s + 1 < n_slabs;
    if (more) {
      gemm_fetch<BM, PA>(sa, g.A, g.lda, m0, M, (s + 1) * GEMM_BK, g.K, g.vec_ok, tid);
      gemm_fetch<BN, PW>(sw, g.W, g.ldw, n0, g.N, (s + 1) * GEMM_BK, g.K, g.vec_ok, tid);
    }
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      float a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; i += (TM >= 4 ? 4 : TM)) {
        if constexpr (TM >= 4) {
          const float4 t = *reinterpret_cast<const float4*>(&sm.a[buf][kk][ty * TM + i]);
          a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
        } else {
          const float2 t = *reinterpret_cast<const float2*>(&sm.a[buf][kk][ty * TM + i]);
          a[i] = t.x; a[i + 1] = t.y;
        }
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&sm.w[buf][kk][tx * TN + j]);
        w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (more) {
      gemm_stash<BM, PA, BM + 4>(sa, sm.a[buf ^ 1], tid);
      gemm_stash<BN, PW, BN + 4>(sw, sm.w[buf ^ 1], tid);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j] + (g.bias != nullptr ? g.bias[n] : 0.f);
      if (g.relu) v = fmaxf(v, 0.f);
      g.C[m * g.ldc + n] = v;
    }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS) sgemm_nt_kernel(const GemmArgs g) {
  extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
  GemmSmemAll& sm = *reinterpret_cast<GemmSmemAll*>(gemm_smem_raw);
  int64_t M = g.M;
  if (g.count != nullptr) {
    const int64_t c = (int64_t)(*g.count) * g.rows_per_count;
    M = c < M ? c : M;
  }
  if (M <= 0) return;
  const int tid = threadIdx.x;
  if (M > 256) {
    const int64_t tm = (M + 127) / 128, tn = (g.N + 127) / 128;
    for (int64_t t = blockIdx.x; t < tm * tn; t += gridDim.x)
      gemm_tile<128, 128, 8, 8>(g, sm.big, (t / tn) * 128, (int)(t % tn) * 128, M, tid);
  } else {
    const int64_t tm = (M + 31) / 32, tn = (g.N + 63) / 64;
    for (int64_t t = blockIdx.x; t < tm * tn; t += gridDim.x)
      gemm_tile<32, 64, 2, 4>(g, sm.small, (t / tn) * 32, (int)(t % tn) * 64, M, tid);
  }
}

static int g_gemm_sms = 0;

extern "C" int tiger_sgemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                              float* C, int64_t ldc, int64_t m_rows, const int32_t* count, int64_t rows_per_count,
                              int n_cols, int k_dim, int relu, void* stream) {
  if (m_rows < 0 || n_cols <= 0 || k_dim <= 0 || lda < k_dim || ldw < k_dim || ldc < n_cols) return TIGER_EINVAL;
  if (m_rows == 0) return TIGER_OK;
  if (g_gemm_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_gemm_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_gemm_sms <= 0) g_gemm_sms = 148;
    if (cudaFuncSetAttribute(sgemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(GemmSmemAll)) != cudaSuccess)
      return TIGER_ECUDA;
  }
  GemmArgs g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.C = C; g.ldc = ldc;
  g.M = m_rows; g.count = count; g.rows_per_count = rows_per_count > 0 ? rows_per_count : 1;
  g.N = n_cols; g.K = k_dim; g.relu = relu;
  g.vec_ok = ((((uintptr_t)A | (uintptr_t)W) & 15) == 0 && (lda & 3) == 0 && (ldw & 3) == 0) ? 1 : 0;
  // enough CTAs for the largest possible problem, never more than two waves
  const int64_t tiles_small = ((m_rows + 31) / 32) * ((n_cols + 63) / 64);
  int64_t grid = tiles_small < 2 * (int64_t)g_gemm_sms ? tiles_small : 2 * (int64_t)g_gemm_sms;
  sgemm_nt_kernel<<<(unsigned)grid, GEMM_THREADS, sizeof(GemmSmemAll), as_stream(stream)>>>(g);
  return tiger_launch_status();
}
