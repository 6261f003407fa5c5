// K2  fused pending-message gather + GRUCell memory update.
// Reference: LastMessageAggregatorNoGradLastOnly.forward (message_modules.py:150-160) gathers
// msg rows, TIGE.apply_messages (tiger.py:339-356) gathers the update-source memory rows and
// GRUUpdater.forward (update_modules.py:30-37) runs nn.GRUCell (gate order r,z,n):
//   r = s(Wir x + bir + Whr h + bhr)   z = s(Wiz x + biz + Whz h + bhz)
//   n = tanh(Win x + bin + r*(Whn h + bhn))   h' = (h - n)*z + n
//
// fp32 FFMA tiled GEMM (parity tolerance 1e-5 rules out single-pass TF32): CTA tile 64 rows x
// 32 hidden units x (3 gates + the separate Whn accumulator), 256 threads, thread tile 4 rows x
// 2 units, K chunks of 32 double-buffered in shared memory with register-staged prefetch.
// Rows are gathered straight from the node-indexed tables (no [O,M] staging copy in HBM).
#include "common.cuh"

#define GRU_TM 64
#define GRU_TJ 32
#define GRU_KC 32
#define GRU_THREADS 256
#define GRU_APAD 4

struct GruTile {
  float a[2][GRU_TM][GRU_KC + GRU_APAD];
  float w[2][GRU_KC][3 * GRU_TJ];
};

// stage one K-chunk: 8 activation floats + 3 weight float4 per thread
struct GruStage {
  float a[8];
  float4 w[3];
};

__device__ __forceinline__ void gru_load_chunk(GruStage& st, const float* const* row_ptr, int k0, int kdim,
                                               const float* __restrict__ wT, int64_t ldw, int dp, int j0,
                                               int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int k = k0 + lane;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float* rp = row_ptr[warp * 8 + i];
    st.a[i] = (rp != nullptr && k < kdim) ? __ldg(rp + k) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int f = tid + i * GRU_THREADS;  // 0..767 : 32 k x 24 float4
    const int kk = f / 24, c4 = f % 24;
    const int g = c4 >> 3, jj = (c4 & 7) << 2;
    st.w[i] = (k0 + kk < kdim)
                  ? __ldg(reinterpret_cast<const float4*>(wT + (int64_t)(k0 + kk) * ldw + g * dp + j0 + jj))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ void gru_store_chunk(const GruStage& st, GruTile& sm, int buf, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) sm.a[buf][warp * 8 + i][lane] = st.a[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int f = tid + i * GRU_THREADS;
    const int kk = f / 24, c4 = f % 24;
    *reinterpret_cast<float4*>(&sm.w[buf][kk][c4 << 2]) = st.w[i];
  }
}

// acc[g][i][c]: g = 0 (r), 1 (z), 2 (n-part of this phase)
__device__ __forceinline__ void gru_compute_chunk(const GruTile& sm, int buf, int ty, int tx, float (&acc_r)[4][2],
                                                  float (&acc_z)[4][2], float (&acc_n)[4][2]) {
#pragma unroll
  for (int kk = 0; kk < GRU_KC; kk += 4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&sm.a[buf][ty * 4 + i][kk]);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float2 wr = *reinterpret_cast<const float2*>(&sm.w[buf][kk + s][tx * 2]);
      const float2 wz = *reinterpret_cast<const float2*>(&sm.w[buf][kk + s][GRU_TJ + tx * 2]);
      const float2 wn = *reinterpret_cast<const float2*>(&sm.w[buf][kk + s][2 * GRU_TJ + tx * 2]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = s == 0 ? a[i].x : (s == 1 ? a[i].y : (s == 2 ? a[i].z : a[i].w));
        acc_r[i][0] = fmaf(av, wr.x, acc_r[i][0]);
        acc_r[i][1] = fmaf(av, wr.y, acc_r[i][1]);
        acc_z[i][0] = fmaf(av, wz.x, acc_z[i][0]);
        acc_z[i][1] = fmaf(av, wz.y, acc_z[i][1]);
        acc_n[i][0] = fmaf(av, wn.x, acc_n[i][0]);
        acc_n[i][1] = fmaf(av, wn.y, acc_n[i][1]);
      }
    }
  }
}

__device__ __forceinline__ void gru_phase(GruTile& sm, const float* const* row_ptr, int kdim,
                                          const float* __restrict__ wT, int64_t ldw, int dp, int j0, int tid,
                                          int ty, int tx, float (&acc_r)[4][2], float (&acc_z)[4][2],
                                          float (&acc_n)[4][2]) {
  const int n_chunks = (kdim + GRU_KC - 1) / GRU_KC;
  GruStage st;
  gru_load_chunk(st, row_ptr, 0, kdim, wT, ldw, dp, j0, tid);
  __syncthreads();  // previous phase finished reading both buffers
  gru_store_chunk(st, sm, 0, tid);
  __syncthreads();
  for (int c = 0; c < n_chunks; ++c) {
    const bool more = c + 1 < n_chunks;
    if (more) gru_load_chunk(st, row_ptr, (c + 1) * GRU_KC, kdim, wT, ldw, dp, j0, tid);
    gru_compute_chunk(sm, c & 1, ty, tx, acc_r, acc_z, acc_n);
    if (more) gru_store_chunk(st, sm, (c + 1) & 1, tid);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(GRU_THREADS)
gru_update_kernel(const int64_t* __restrict__ node_ids, const int32_t* __restrict__ count, int64_t n_rows,
                  const float* __restrict__ x_table, int64_t x_stride, const float* __restrict__ h_table,
                  int64_t h_stride, int m_dim, int d, const float* __restrict__ wT_ih,
                  const float* __restrict__ wT_hh, int64_t ldw, int dp, const float* __restrict__ b_ih,
                  const float* __restrict__ b_hh, float* __restrict__ h_new, const float* __restrict__ msg_ts,
                  const float* __restrict__ check_mem_ts, int check_equal, uint32_t* __restrict__ err_flags) {
  __shared__ GruTile sm;
  __shared__ const float* x_ptr[GRU_TM];
  __shared__ const float* h_ptr[GRU_TM];
  int64_t n = n_rows;
  if (count != nullptr) {
    const int64_t c = *count;
    n = c < n ? c : n;
  }
  const int64_t row0 = (int64_t)blockIdx.y * GRU_TM;
  if (row0 >= n) return;
  const int j0 = blockIdx.x * GRU_TJ;
  const int tid = threadIdx.x;
  if (tid < GRU_TM) {
    const int64_t r = row0 + tid;
    const float* xp = nullptr;
    const float* hp = nullptr;
    if (r < n) {
      const int64_t u = node_ids != nullptr ? node_ids[r] : r;
      xp = x_table + u * x_stride;
      hp = h_table + u * h_stride;
      if (check_mem_ts != nullptr && blockIdx.x == 0 && err_flags != nullptr) {
        const float mt = msg_ts[u], pt = check_mem_ts[u];
        if (pt > mt) atomicOr(err_flags, TIGER_ERR_MSG_BEFORE_MEM);       // message_modules.py:157-159
        if (check_equal && mt != pt) atomicOr(err_flags, TIGER_ERR_MSG_TS_MISMATCH);  // tiger.py:324-327
      }
    }
    x_ptr[tid] = xp;
    h_ptr[tid] = hp;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float acc_r[4][2] = {}, acc_z[4][2] = {}, acc_in[4][2] = {}, acc_hn[4][2] = {};
  gru_phase(sm, x_ptr, m_dim, wT_ih, ldw, dp, j0, tid, ty, tx, acc_r, acc_z, acc_in);
  gru_phase(sm, h_ptr, d, wT_hh, ldw, dp, j0, tid, ty, tx, acc_r, acc_z, acc_hn);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int j = j0 + tx * 2 + c;
    if (j >= d) continue;
    const float br = b_ih[j] + b_hh[j], bz = b_ih[d + j] + b_hh[d + j];
    const float bin = b_ih[2 * d + j], bhn = b_hh[2 * d + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = ty * 4 + i;
      if (row0 + rl >= n) continue;
      const float r = sigmoidf_acc(acc_r[i][c] + br);
      const float z = sigmoidf_acc(acc_z[i][c] + bz);
      const float nn = tanhf((acc_in[i][c] + bin) + r * (acc_hn[i][c] + bhn));
      const float h = h_ptr[rl][j];
      h_new[(row0 + rl) * d + j] = (h - nn) * z + nn;
    }
  }
}

extern "C" int tiger_gru_update(const int64_t* node_ids, const int32_t* count, int64_t n_rows,
                                const float* x_table, int64_t x_stride, const float* h_table, int64_t h_stride,
                                int m_dim, int d, const float* wT_ih, const float* wT_hh, int64_t ldw,
                                const float* b_ih, const float* b_hh, float* h_new, const float* msg_ts,
                                const float* check_mem_ts, int check_equal, uint32_t* err_flags, void* stream) {
  if (n_rows < 0 || m_dim <= 0 || d <= 0) return TIGER_EINVAL;
  const int dp = (d + 31) / 32 * 32;
  if (ldw < 3 * (int64_t)dp || (ldw & 3) != 0 || (((uintptr_t)wT_ih | (uintptr_t)wT_hh) & 15) != 0)
    return TIGER_EINVAL;
  if (n_rows == 0) return TIGER_OK;
  dim3 grid((unsigned)(dp / GRU_TJ), (unsigned)((n_rows + GRU_TM - 1) / GRU_TM));
  gru_update_kernel<<<grid, GRU_THREADS, 0, as_stream(stream)>>>(
      node_ids, count, n_rows, x_table, x_stride, h_table, h_stride, m_dim, d, wT_ih, wT_hh, ldw, dp, b_ih, b_hh,
      h_new, msg_ts, check_mem_ts, check_equal, err_flags);
  return tiger_launch_status();
}
