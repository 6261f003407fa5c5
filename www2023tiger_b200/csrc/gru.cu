// K2  fused pending-message gather + GRUCell memory update on the tcgen05 tensor cores.
// Reference: LastMessageAggregatorNoGradLastOnly.forward (message_modules.py:150-160) gathers
// msg rows, TIGE.apply_messages (tiger.py:339-356) gathers the update-source memory rows and
// GRUUpdater.forward (update_modules.py:30-37) runs nn.GRUCell (gate order r,z,n):
//   r = s(Wir x + bir + Whr h + bhr)   z = s(Wiz x + biz + Whz h + bhz)
//   n = tanh(Win x + bin + r*(Whn h + bhn))   h' = (h - n)*z + n
//
// One CTA = 128 gathered rows x U hidden units (U = 32): a [128 x 4U] fp32 accumulator in TMEM with
// the column groups [ Win x | r | z | Whn h ].  The K loop runs over the message columns first
// (x phase: one MMA group of N = 3U per k-step into [Win x | r | z], weight rows as stored in
// weight_ih) and then over the state columns (h phase: N = 2U accumulating into [r | z] and N = U into
// [Whn h], weight rows as stored in weight_hh), all in tf32x3 (umma.cuh) so the result keeps fp32
// accuracy: three issuer warps, four partial accumulator sets of 4U columns = all 512 TMEM columns.
// Rows are gathered straight from the node-indexed tables by the producer warps (no [O, M] staging copy
// in HBM; 4 groups of 4 warps keep 4 stages of loads in flight); the same warps then run the gate
// epilogue out of TMEM.  The gate weights are pre-split once per parameter update (tiger_gru_pack) into
// the shared-memory image of each stage, which one TMA bulk copy per stage brings in.
#include "common.cuh"
#include "umma.cuh"

#define GRU_PRODUCER_WARPS 16
#define GRU_THREADS ((GRU_PRODUCER_WARPS + UMMA_ISSUERS + 1) * 32)   // + the TMA warp
#define GRU_BM 128
#define GRU_U 32                       // hidden units per CTA (multiple of 16)
#define GRU_WROWS (3 * GRU_U)          // weight rows per stage
#define GRU_STAGES 6
#define GRU_GROUPS 4                   // producer groups, each fills every 4th stage
#define GRU_GROUP_WARPS (GRU_PRODUCER_WARPS / GRU_GROUPS)
#define GRU_A_PLANE (UMMA_KCH * GRU_BM * 4)      // floats
#define GRU_W_PLANE (UMMA_KCH * GRU_WROWS * 4)
#define GRU_STAGE_FLOATS (2 * GRU_A_PLANE + 2 * GRU_W_PLANE)
#define GRU_ACC_COLS (4 * GRU_U)       // one accumulator set: [ Win x | r | z | Whn h ]
#define GRU_TMEM_COLS (UMMA_ACCS * GRU_ACC_COLS)   // 512: the four tf32x3 partial accumulators (umma.cuh)
#define GRU_NA (GRU_BM / 8 / GRU_GROUP_WARPS)       // A warp-chunks per producer warp
#define GRU_SMEM_BYTES (GRU_STAGES * GRU_STAGE_FLOATS * 4 + 2 * GRU_BM * 8 + 4 * GRU_U * 4 + 128)

struct GruArgs {
  const int64_t* node_ids;
  const int32_t* count;
  int64_t n_rows;
  const float* x_table;
  int64_t x_stride;
  const float* h_table;
  int64_t h_stride;
  int m_dim, d;
  const float* wpack;  // tiger_gru_pack: pre-split gate weights, one contiguous stage image per (unit tile, k-block)
  const float* b_ih;
  const float* b_hh;
  float* h_new;
  const float* msg_ts;
  const float* check_mem_ts;
  int check_equal;
  uint32_t* err_flags;
  int vec_x, vec_h;
};

__global__ void __launch_bounds__(GRU_THREADS, 1) gru_update_kernel(const GruArgs g) {
  extern __shared__ __align__(128) unsigned char gru_smem[];
  float* stage0 = reinterpret_cast<float*>(gru_smem);
  const float** x_ptr = reinterpret_cast<const float**>(gru_smem + (size_t)GRU_STAGES * GRU_STAGE_FLOATS * 4);
  const float** h_ptr = x_ptr + GRU_BM;
  float* bias_s = reinterpret_cast<float*>(h_ptr + GRU_BM);   // [4][GRU_U]: b_r, b_z, b_in, b_hn
  uint64_t* full = reinterpret_cast<uint64_t*>(bias_s + 4 * GRU_U);
  uint64_t* empty = full + GRU_STAGES;
  uint64_t* done = empty + GRU_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  int64_t n = g.n_rows;
  if (g.count != nullptr) {
    const int64_t c = *g.count;
    n = c < n ? c : n;
  }
  const int64_t row0 = (int64_t)blockIdx.y * GRU_BM;
  if (row0 >= n) return;
  const int j0 = blockIdx.x * GRU_U;
  const int d = g.d, m_dim = g.m_dim;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid < GRU_BM) {
    // rows beyond the count read the last valid row: they only feed accumulator rows nobody stores
    int64_t r = row0 + tid;
    const bool live = r < n;
    r = live ? r : n - 1;
    const int64_t u = g.node_ids != nullptr ? g.node_ids[r] : r;
    x_ptr[tid] = g.x_table + u * g.x_stride;
    h_ptr[tid] = g.h_table + u * g.h_stride;
    if (live && g.check_mem_ts != nullptr && blockIdx.x == 0 && g.err_flags != nullptr) {
      const float mt = g.msg_ts[u], pt = g.check_mem_ts[u];
      if (pt > mt) atomicOr(g.err_flags, TIGER_ERR_MSG_BEFORE_MEM);                    // message_modules.py:157-159
      if (g.check_equal && mt != pt) atomicOr(g.err_flags, TIGER_ERR_MSG_TS_MISMATCH);  // tiger.py:324-327
    }
  } else if (tid < GRU_BM + GRU_U) {
    const int jj = tid - GRU_BM, j = j0 + jj;
    const bool ok = j < d;
    bias_s[jj] = ok ? g.b_ih[j] + g.b_hh[j] : 0.f;
    bias_s[GRU_U + jj] = ok ? g.b_ih[d + j] + g.b_hh[d + j] : 0.f;
    bias_s[2 * GRU_U + jj] = ok ? g.b_ih[2 * d + j] : 0.f;
    bias_s[3 * GRU_U + jj] = ok ? g.b_hh[2 * d + j] : 0.f;
  }
  if (tid == GRU_PRODUCER_WARPS * 32) {
    for (int s = 0; s < GRU_STAGES; ++s) {
      mbar_init(full + s, GRU_GROUP_WARPS + 1);
      mbar_init(empty + s, UMMA_ISSUERS);
    }
    mbar_init(done, UMMA_ISSUERS);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, GRU_TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = *tmem_slot;
  const int nbx = (m_dim + UMMA_BK - 1) / UMMA_BK, nbh = (d + UMMA_BK - 1) / UMMA_BK;
  const int n_blocks = nbx + nbh;

  if (warp < GRU_PRODUCER_WARPS) {
    // ---------------- producers ----------------
    // group `grp` fills stages grp, grp + GROUPS, ...: while one group waits for its loads the others
    // convert / publish theirs, so GROUPS stages worth of global loads are always in flight
    const int grp = warp / GRU_GROUP_WARPS, wg = warp % GRU_GROUP_WARPS;
    UmmaChunks<GRU_NA> ca;
    // chunk geometry is phase-independent; only the row pointers change between the x and the h phase
    auto set_phase = [&](bool xph) {
      const float* const* rows = xph ? x_ptr : h_ptr;
#pragma unroll
      for (int i = 0; i < GRU_NA; ++i) {
        int row, kc;
        umma_chunk_pos(wg + GRU_GROUP_WARPS * i, lane, row, kc);
        ca.ptr[i] = rows[row] + kc * 4;
        ca.soff[i] = (kc * GRU_BM + row) * 4;
        ca.kq[i] = kc * 4;
      }
    };
    // two register sets per thread: the loads of this group's next two stages are in flight while the
    // current one is converted and published
    float4 va0[GRU_NA], va1[GRU_NA];
    bool in_x = true;
    set_phase(true);
    auto load = [&](float4 (&va)[GRU_NA], int blk) {
      if (blk >= n_blocks) return;
      const bool xph = blk < nbx;
      if (xph != in_x) {
        set_phase(xph);
        in_x = xph;
      }
      const int k0 = (xph ? blk : blk - nbx) * UMMA_BK;
      umma_chunks_load(va, ca, k0, xph ? m_dim : d, (xph ? g.vec_x : g.vec_h) != 0);
    };
    auto publish = [&](const float4 (&va)[GRU_NA], int blk) {
      const int s = blk % GRU_STAGES;
      float* a_hi = stage0 + (size_t)s * GRU_STAGE_FLOATS;
      float* a_lo = a_hi + GRU_A_PLANE;
      mbar_wait(empty + s, ((blk / GRU_STAGES) & 1) ^ 1);
      umma_chunks_store(a_hi, a_lo, ca, va);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + s);
    };
    load(va0, grp);
    load(va1, grp + GRU_GROUPS);
    for (int blk = grp; blk < n_blocks; blk += 2 * GRU_GROUPS) {
      publish(va0, blk);
      load(va0, blk + 2 * GRU_GROUPS);
      if (blk + GRU_GROUPS < n_blocks) {
        publish(va1, blk + GRU_GROUPS);
        load(va1, blk + 3 * GRU_GROUPS);
      }
    }
    // ---------------- gate epilogue ----------------
    mbar_wait(done, 0);
    tc_fence_after_sync();
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const bool row_ok = row0 + rl < n;
    const float* hp = h_ptr[rl];
    const uint32_t tl = taddr + ((uint32_t)(q * 32) << 16);
    for (int c0 = (warp >> 2) * 16; c0 < GRU_U; c0 += 16 * (GRU_PRODUCER_WARPS / 4)) {
      float an[16], ar[16], az[16], ah[16];
      tmem_ld16(tl + (uint32_t)c0, an);
      tmem_ld16(tl + (uint32_t)(GRU_U + c0), ar);
      tmem_ld16(tl + (uint32_t)(2 * GRU_U + c0), az);
      tmem_ld16(tl + (uint32_t)(3 * GRU_U + c0), ah);
#pragma unroll
      for (int jb = 1; jb < UMMA_ACCS; ++jb) {
        float t[16];
        const uint32_t tb = tl + (uint32_t)(jb * GRU_ACC_COLS + c0);
        tmem_ld16(tb, t);
#pragma unroll
        for (int e = 0; e < 16; ++e) an[e] += t[e];
        tmem_ld16(tb + GRU_U, t);
#pragma unroll
        for (int e = 0; e < 16; ++e) ar[e] += t[e];
        tmem_ld16(tb + 2 * GRU_U, t);
#pragma unroll
        for (int e = 0; e < 16; ++e) az[e] += t[e];
        tmem_ld16(tb + 3 * GRU_U, t);
#pragma unroll
        for (int e = 0; e < 16; ++e) ah[e] += t[e];
      }
      if (!row_ok) continue;
      float* dst = g.h_new + (row0 + rl) * d;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int j = j0 + c0 + jj;
        if (j < d) {
          const float r = sigmoidf_acc(ar[jj] + bias_s[c0 + jj]);
          const float z = sigmoidf_acc(az[jj] + bias_s[GRU_U + c0 + jj]);
          const float nn = tanhf((an[jj] + bias_s[2 * GRU_U + c0 + jj]) + r * (ah[jj] + bias_s[3 * GRU_U + c0 + jj]));
          const float h = hp[j];
          dst[j] = (h - nn) * z + nn;
        }
      }
    }
  } else if (warp == GRU_PRODUCER_WARPS + UMMA_ISSUERS) {
    // ---------------- TMA warp: one bulk copy per stage brings both planes of the weight tile ----------------
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)UMMA_PACK_STAGE_FLOATS(GRU_WROWS) * 4u;
      const float* src = g.wpack + (int64_t)blockIdx.x * n_blocks * UMMA_PACK_STAGE_FLOATS(GRU_WROWS);
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int s = blk % GRU_STAGES;
        mbar_wait(empty + s, ((blk / GRU_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(full + s, bytes);
        tma_bulk_load(stage0 + (size_t)s * GRU_STAGE_FLOATS + 2 * GRU_A_PLANE,
                      src + (int64_t)blk * UMMA_PACK_STAGE_FLOATS(GRU_WROWS), bytes, full + s);
      }
    }
  } else {
    // ---------------- MMA issuers (one elected thread per role, see umma.cuh) ----------------
    // The whole warp runs the loop so that every operand stays warp-uniform; only the tcgen05 instructions
    // are issued by the elected lane.
    const int role = uniform_warp_idx() - GRU_PRODUCER_WARPS;
    const UmmaRole r = umma_role(role, smem_addr_u32(stage0), GRU_STAGE_FLOATS * 4u, GRU_BM, GRU_WROWS, GRU_ACC_COLS);
    const uint32_t id3 = umma_idesc_tf32(GRU_BM, 3 * GRU_U), id2 = umma_idesc_tf32(GRU_BM, 2 * GRU_U),
                   id1 = umma_idesc_tf32(GRU_BM, GRU_U);
    const uint32_t tbase = __shfl_sync(0xffffffffu, taddr, 0);
    const uint32_t d_even = tbase + r.acc_even, d_odd = tbase + r.acc_odd;
    int s = 0;
    uint32_t ph = 0, a = r.a_lo, b = r.b_lo;
    for (int blk = 0; blk < n_blocks; ++blk) {
      mbar_wait(full + s, ph);
      tc_fence_after_sync();
      const uint32_t a1 = a + r.a_kstep, b1 = b + r.b_kstep;
      if (elect_one()) {
        if (blk < nbx) {
          // x phase: weight tile rows [n | r | z] -> columns [0, 3U)
          umma_tf32_lo(d_even, a, b, id3, blk > 0 ? 1u : 0u);
          umma_tf32_lo(d_odd, a1, b1, id3, (role == 2 && blk == 0) ? 0u : 1u);
        } else {
          // h phase: weight tile rows [r | z | n] -> [r | z] accumulate at column U, Whn h starts at column 3U
          const uint32_t fresh = blk == nbx ? 0u : 1u;
          umma_tf32_lo(d_even + GRU_U, a, b, id2, 1u);
          umma_tf32_lo(d_even + 3 * GRU_U, a, b + 2 * GRU_U, id1, fresh);
          umma_tf32_lo(d_odd + GRU_U, a1, b1, id2, 1u);
          umma_tf32_lo(d_odd + 3 * GRU_U, a1, b1 + 2 * GRU_U, id1, role == 2 ? fresh : 1u);
        }
        umma_commit(empty + s);
      }
      __syncwarp();
      a += r.stage_step;
      b += r.stage_step;
      if (++s == GRU_STAGES) {
        s = 0;
        ph ^= 1;
        a = r.a_lo;
        b = r.b_lo;
      }
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(taddr, GRU_TMEM_COLS);
}

// Gate-weight pack: for unit tile t and k-block kb (x phase: kb < nbx over weight_ih, rows [n | r | z]; h phase over
// weight_hh, rows [r | z | n]) the stage image [head plane | tail plane][kc][96 rows][4 floats].
__global__ void gru_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, int m_dim, int d,
                                int tiles, int nbx, int nbh, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_kb = nbx + nbh;
  const int64_t total = (int64_t)tiles * n_kb * UMMA_KCH * GRU_WROWS;
  if (i >= total) return;
  const int r = (int)(i % GRU_WROWS);
  const int kc = (int)((i / GRU_WROWS) % UMMA_KCH);
  const int kb = (int)((i / (GRU_WROWS * UMMA_KCH)) % n_kb);
  const int t = (int)(i / ((int64_t)GRU_WROWS * UMMA_KCH * n_kb));
  const bool xph = kb < nbx;
  const int gi = r / GRU_U, j = t * GRU_U + r % GRU_U;
  const int gate = xph ? (gi == 0 ? 2 : gi - 1) : gi;   // gate order in the weights: r, z, n
  const int kdim = xph ? m_dim : d;
  const float* w = xph ? w_ih : w_hh;
  const int k = (xph ? kb : kb - nbx) * UMMA_BK + kc * 4;
  const float4 v = umma_load_chunk(j < d ? w + (int64_t)(gate * d + j) * kdim : nullptr, k, kdim, false);
  float4 h, l;
  tf32_split(v, h, l);
  float* stage = out + ((int64_t)t * n_kb + kb) * UMMA_PACK_STAGE_FLOATS(GRU_WROWS);
  *reinterpret_cast<float4*>(stage + (kc * GRU_WROWS + r) * 4) = h;
  *reinterpret_cast<float4*>(stage + UMMA_KCH * GRU_WROWS * 4 + (kc * GRU_WROWS + r) * 4) = l;
}

extern "C" int64_t tiger_gru_pack_bytes(int m_dim, int d) {
  if (m_dim <= 0 || d <= 0) return -1;
  const int64_t tiles = (d + GRU_U - 1) / GRU_U;
  const int64_t n_kb = (m_dim + UMMA_BK - 1) / UMMA_BK + (d + UMMA_BK - 1) / UMMA_BK;
  return tiles * n_kb * UMMA_PACK_STAGE_FLOATS(GRU_WROWS) * (int64_t)sizeof(float);
}

extern "C" int tiger_gru_pack(const float* w_ih, const float* w_hh, int m_dim, int d, float* out, void* stream) {
  if (w_ih == nullptr || w_hh == nullptr || out == nullptr || tiger_gru_pack_bytes(m_dim, d) < 0 ||
      (((uintptr_t)out) & 15) != 0)
    return TIGER_EINVAL;
  const int tiles = (d + GRU_U - 1) / GRU_U;
  const int nbx = (m_dim + UMMA_BK - 1) / UMMA_BK, nbh = (d + UMMA_BK - 1) / UMMA_BK;
  const int64_t total = (int64_t)tiles * (nbx + nbh) * UMMA_KCH * GRU_WROWS;
  gru_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(w_ih, w_hh, m_dim, d, tiles, nbx, nbh,
                                                                                 out);
  return tiger_launch_status();
}

extern "C" int tiger_gru_update(const int64_t* node_ids, const int32_t* count, int64_t n_rows,
                                const float* x_table, int64_t x_stride, const float* h_table, int64_t h_stride,
                                int m_dim, int d, const float* wpack, const float* b_ih, const float* b_hh,
                                float* h_new, const float* msg_ts, const float* check_mem_ts, int check_equal,
                                uint32_t* err_flags, void* stream) {
  if (n_rows < 0 || m_dim <= 0 || d <= 0 || wpack == nullptr || (((uintptr_t)wpack) & 15) != 0) return TIGER_EINVAL;
  if (n_rows == 0) return TIGER_OK;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gru_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRU_SMEM_BYTES) !=
        cudaSuccess)
      return TIGER_ECUDA;
    configured = true;
  }
  GruArgs g;
  g.node_ids = node_ids; g.count = count; g.n_rows = n_rows;
  g.x_table = x_table; g.x_stride = x_stride; g.h_table = h_table; g.h_stride = h_stride;
  g.m_dim = m_dim; g.d = d; g.wpack = wpack; g.b_ih = b_ih; g.b_hh = b_hh; g.h_new = h_new;
  g.msg_ts = msg_ts; g.check_mem_ts = check_mem_ts; g.check_equal = check_equal; g.err_flags = err_flags;
  g.vec_x = ((((uintptr_t)x_table) & 15) == 0 && (x_stride & 3) == 0) ? 1 : 0;
  g.vec_h = ((((uintptr_t)h_table) & 15) == 0 && (h_stride & 3) == 0) ? 1 : 0;
  dim3 grid((unsigned)((d + GRU_U - 1) / GRU_U), (unsigned)((n_rows + GRU_BM - 1) / GRU_BM));
  gru_update_kernel<<<grid, GRU_THREADS, GRU_SMEM_BYTES, as_stream(stream)>>>(g);
  return tiger_launch_status();
}
