// K2  fused pending-message gather + GRUCell memory update on the tcgen05 tensor cores.
// Reference: LastMessageAggregatorNoGradLastOnly.forward (message_modules.py:150-160) gathers
// msg rows, TIGE.apply_messages (tiger.py:339-356) gathers the update-source memory rows and
// GRUUpdater.forward (update_modules.py:30-37) runs nn.GRUCell (gate order r,z,n):
//   r = s(Wir x + bir + Whr h + bhr)   z = s(Wiz x + biz + Whz h + bhz)
//   n = tanh(Win x + bin + r*(Whn h + bhn))   h' = (h - n)*z + n
//
// One CTA = 128 gathered rows x U = 32 hidden units.  Accumulator set in TMEM: [128 x 4U] fp32 with the
// column groups [ Win x | r | z | Whn h ].  The K loop runs over the message columns first (x phase: one
// MMA of N = 3U per k-step into [Win x | r | z], weight tile rows [n | r | z] of weight_ih) and then over
// the state columns (h phase: N = 2U accumulating into [r | z] and N = U into [Whn h], weight tile rows
// [r | z | n] of weight_hh), in tf32x3 (umma.cuh) so the result keeps fp32 accuracy.
//
// Data movement (everything the tensor core reads is produced without shared-memory traffic from the SM):
//   activations  the 16 producer warps (2 groups x 8 warps; two threads per row, 64 bytes each) gather
//                one 128-byte line of every row per stage straight from the node-indexed tables (no
//                [O, M] staging copy in HBM), split it into tf32 head / tail and write it into a 2-stage
//                ring in TENSOR MEMORY (tcgen05.st); the MMAs take A from TMEM and run at the math floor
//                (tools/umma_bench.cu).  A stage is always refilled by the group that filled it before:
//                an mbarrier wait only distinguishes two phases, so a group that skipped a phase of a
//                stage's barrier could run a full round ahead of the tensor core.
//   weights      pre-split once per parameter update (tiger_gru_pack) into the shared-memory image of
//                every stage; one TMA bulk copy per stage into a 6-stage ring
//   MMA issue    two warps (one elected thread each): X issues the cross terms tail*head + head*tail
//                into accumulator set 0, Y the head*head terms alternating between sets 1 and 2 (the
//                tensor core truncates when it accumulates; spreading the large products over two
//                accumulators halves that bias); 3 sets x 128 + 2 x 64 activation columns = 512 TMEM columns
//   epilogue     the producer warps add the three partial sets and apply the gates out of TMEM
#include "common.cuh"
#include "umma.cuh"

#define GRU_PRODUCER_WARPS 16
#define GRU_GROUPS 2                   // producer groups: group g fills activation stage g (GROUPS <= A_STAGES, see below)
#define GRU_GROUP_WARPS (GRU_PRODUCER_WARPS / GRU_GROUPS)
#define GRU_ISSUERS 2
#define GRU_THREADS ((GRU_PRODUCER_WARPS + GRU_ISSUERS + 1) * 32)   // + the TMA warp
#define GRU_BM 128
#define GRU_U 32                       // hidden units per CTA (multiple of 16)
#define GRU_WROWS (3 * GRU_U)          // weight rows per stage
#define GRU_ACC_COLS (4 * GRU_U)       // one accumulator set: [ Win x | r | z | Whn h ]
#define GRU_SETS 3
#define GRU_A_STAGES 2
#define GRU_W_STAGES 6
#define GRU_A_RING (GRU_SETS * GRU_ACC_COLS)                          // first TMEM column of the activation ring
#define GRU_TMEM_COLS 512
#define GRU_W_STAGE_FLOATS UMMA_PACK_STAGE_FLOATS(GRU_WROWS)
#define GRU_SMEM_BYTES (GRU_W_STAGES * GRU_W_STAGE_FLOATS * 4 + 4 * GRU_U * 4 + 256)

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
  float* wstage0 = reinterpret_cast<float*>(gru_smem);
  float* bias_s = wstage0 + (size_t)GRU_W_STAGES * GRU_W_STAGE_FLOATS;   // [4][GRU_U]: b_r, b_z, b_in, b_hn
  uint64_t* a_full = reinterpret_cast<uint64_t*>(bias_s + 4 * GRU_U);
  uint64_t* a_empty = a_full + GRU_A_STAGES;
  uint64_t* w_full = a_empty + GRU_A_STAGES;
  uint64_t* w_empty = w_full + GRU_W_STAGES;
  uint64_t* done = w_empty + GRU_W_STAGES;
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

  if (tid < GRU_U) {
    const int j = j0 + tid;
    const bool ok = j < d;
    bias_s[tid] = ok ? g.b_ih[j] + g.b_hh[j] : 0.f;
    bias_s[GRU_U + tid] = ok ? g.b_ih[d + j] + g.b_hh[d + j] : 0.f;
    bias_s[2 * GRU_U + tid] = ok ? g.b_ih[2 * d + j] : 0.f;
    bias_s[3 * GRU_U + tid] = ok ? g.b_hh[2 * d + j] : 0.f;
  }
  if (tid == GRU_PRODUCER_WARPS * 32) {
    for (int s = 0; s < GRU_A_STAGES; ++s) {
      mbar_init(a_full + s, GRU_GROUP_WARPS);
      mbar_init(a_empty + s, GRU_ISSUERS);
    }
    for (int s = 0; s < GRU_W_STAGES; ++s) {
      mbar_init(w_full + s, 1);
      mbar_init(w_empty + s, GRU_ISSUERS);
    }
    mbar_init(done, GRU_ISSUERS);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, GRU_TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = *tmem_slot;
  const int nbx = (m_dim + TS_BK - 1) / TS_BK, nbh = (d + TS_BK - 1) / TS_BK;
  const int n_blocks = nbx + nbh;

  if (warp < GRU_PRODUCER_WARPS) {
    // ---------------- producers: two threads per row ----------------
    const int grp = warp / GRU_GROUP_WARPS, wg = warp % GRU_GROUP_WARPS;
    const int q = wg & 3, half = wg >> 2;        // TMEM lane quadrant of this warp, which 16 floats of the stage
    const int rl = q * 32 + lane;
    // rows beyond the count read the last valid row: they only feed accumulator rows nobody stores
    int64_t r = row0 + rl;
    const bool live = r < n;
    r = live ? r : n - 1;
    const int64_t u = g.node_ids != nullptr ? g.node_ids[r] : r;
    const float* xp = g.x_table + u * g.x_stride;
    const float* hp = g.h_table + u * g.h_stride;
    if (live && warp < 4 && g.check_mem_ts != nullptr && blockIdx.x == 0 && g.err_flags != nullptr) {
      const float mt = g.msg_ts[u], pt = g.check_mem_ts[u];
      if (pt > mt) atomicOr(g.err_flags, TIGER_ERR_MSG_BEFORE_MEM);                    // message_modules.py:157-159
      if (g.check_equal && mt != pt) atomicOr(g.err_flags, TIGER_ERR_MSG_TS_MISMATCH);  // tiger.py:324-327
    }
    const uint32_t tl = taddr + GRU_A_RING + ((uint32_t)(q * 32) << 16) + 16u * half;
    float4 v[4];
    auto load = [&](int blk) {
      const bool xph = blk < nbx;
      const float* p = xph ? xp : hp;
      const int k0 = (xph ? blk : blk - nbx) * TS_BK + 16 * half, kdim = xph ? m_dim : d;
      if ((xph ? g.vec_x : g.vec_h) != 0 && k0 + 16 <= kdim) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(p + k0) + i);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = umma_load_chunk(p, k0 + 4 * i, kdim, false);
      }
    };
    if (grp < n_blocks) load(grp);
    for (int blk = grp; blk < n_blocks; blk += GRU_GROUPS) {
      const int s = blk % GRU_A_STAGES;
      mbar_wait(a_empty + s, ((blk / GRU_A_STAGES) & 1) ^ 1);
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 h, l;
        tf32_split(v[i], h, l);
        hi[4 * i] = h.x; hi[4 * i + 1] = h.y; hi[4 * i + 2] = h.z; hi[4 * i + 3] = h.w;
        lo[4 * i] = l.x; lo[4 * i + 1] = l.y; lo[4 * i + 2] = l.z; lo[4 * i + 3] = l.w;
      }
      const uint32_t ts = tl + (uint32_t)(s * 2 * TS_BK);
      tmem_st16(ts, hi);
      tmem_st16(ts + TS_BK, lo);
      tmem_wait_st();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + s);
      // this group's next line is in flight while the other group's stage is written / consumed
      if (blk + GRU_GROUPS < n_blocks) load(blk + GRU_GROUPS);
    }
    // ---------------- gate epilogue ----------------
    mbar_wait(done, 0);
    tc_fence_after_sync();
    const uint32_t tacc = taddr + ((uint32_t)(q * 32) << 16);   // q == warp & 3: same rows as in the producer role
    for (int c0 = (warp >> 2) * 16; c0 < GRU_U; c0 += 16 * (GRU_PRODUCER_WARPS / 4)) {
      float an[16], ar[16], az[16], ah[16];
      tmem_ld16(tacc + (uint32_t)c0, an);
      tmem_ld16(tacc + (uint32_t)(GRU_U + c0), ar);
      tmem_ld16(tacc + (uint32_t)(2 * GRU_U + c0), az);
      tmem_ld16(tacc + (uint32_t)(3 * GRU_U + c0), ah);
#pragma unroll
      for (int jb = 1; jb < GRU_SETS; ++jb) {
        float t[16];
        const uint32_t tb = tacc + (uint32_t)(jb * GRU_ACC_COLS + c0);
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
      if (!live) continue;
      float* dst = g.h_new + (row0 + rl) * d;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int j = j0 + c0 + jj;
        if (j < d) {
          const float rg = sigmoidf_acc(ar[jj] + bias_s[c0 + jj]);
          const float zg = sigmoidf_acc(az[jj] + bias_s[GRU_U + c0 + jj]);
          const float nn = tanhf((an[jj] + bias_s[2 * GRU_U + c0 + jj]) + rg * (ah[jj] + bias_s[3 * GRU_U + c0 + jj]));
          const float h = hp[j];
          dst[j] = (h - nn) * zg + nn;
        }
      }
    }
  } else if (warp == GRU_PRODUCER_WARPS + GRU_ISSUERS) {
    // ---------------- TMA warp: one bulk copy per stage brings both planes of the weight tile ----------------
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)GRU_W_STAGE_FLOATS * 4u;
      const float* src = g.wpack + (int64_t)blockIdx.x * n_blocks * GRU_W_STAGE_FLOATS;
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int s = blk % GRU_W_STAGES;
        mbar_wait(w_empty + s, ((blk / GRU_W_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(w_full + s, bytes);
        tma_bulk_load(wstage0 + (size_t)s * GRU_W_STAGE_FLOATS, src + (int64_t)blk * GRU_W_STAGE_FLOATS, bytes,
                      w_full + s);
      }
    }
  } else {
    // ---------------- MMA issuers: X (cross terms) and Y (head*head) ----------------
    // The whole warp runs the loop so that every operand stays warp-uniform; only the tcgen05 instructions
    // are issued by the elected lane.
    const bool is_x = uniform_warp_idx() == GRU_PRODUCER_WARPS;
    const uint32_t id3 = umma_idesc_tf32(GRU_BM, 3 * GRU_U), id2 = umma_idesc_tf32(GRU_BM, 2 * GRU_U),
                   id1 = umma_idesc_tf32(GRU_BM, GRU_U);
    const uint32_t tbase = __shfl_sync(0xffffffffu, taddr, 0);
    const uint32_t a_ring = tbase + GRU_A_RING;
    const uint32_t b_hi0 = umma_desc_lo(smem_addr_u32(wstage0), GRU_WROWS);
    const uint32_t b_plane = (uint32_t)(TS_KCH * GRU_WROWS * 16) >> 4;      // head plane -> tail plane, 16-byte units
    const uint32_t b_stage = (uint32_t)(GRU_W_STAGE_FLOATS * 4) >> 4, b_kstep = 2u * GRU_WROWS;
    bool ready_a = mbar_test(a_full, 0), ready_w = mbar_test(w_full, 0);
    for (int blk = 0; blk < n_blocks; ++blk) {
      const int sa = blk % GRU_A_STAGES, sw = blk % GRU_W_STAGES;
      mbar_wait_probed(ready_a, a_full + sa, (blk / GRU_A_STAGES) & 1);
      mbar_wait_probed(ready_w, w_full + sw, (blk / GRU_W_STAGES) & 1);
      tc_fence_after_sync();
      // probe the next stage now: the probes' latency overlaps the MMA issue below
      const int nb = blk + 1;
      ready_a = nb < n_blocks ? mbar_test(a_full + nb % GRU_A_STAGES, (nb / GRU_A_STAGES) & 1) : true;
      ready_w = nb < n_blocks ? mbar_test(w_full + nb % GRU_W_STAGES, (nb / GRU_W_STAGES) & 1) : true;
      if (elect_one()) {
        const uint32_t a_hi = a_ring + (uint32_t)(sa * 2 * TS_BK), a_lo = a_hi + TS_BK;
        const uint32_t b_hi = b_hi0 + (uint32_t)sw * b_stage, b_lo = b_hi + b_plane;
        const bool xph = blk < nbx;
        const bool first = xph ? blk == 0 : blk == nbx;
#pragma unroll
        for (int j = 0; j < TS_BK / 8; ++j) {
          const uint32_t ah = a_hi + 8u * j, al = a_lo + 8u * j, bh = b_hi + b_kstep * j, bl = b_lo + b_kstep * j;
          if (is_x) {
            const uint32_t acc = tbase;                                  // set 0
            const uint32_t fresh = (first && j == 0) ? 0u : 1u;
            if (xph) {
              // weight tile rows [n | r | z] -> columns [0, 3U)
              umma_tf32_ts(acc, al, bh, id3, fresh);
              umma_tf32_ts(acc, ah, bl, id3, 1u);
            } else {
              // weight tile rows [r | z | n] -> [r | z] accumulate at column U, Whn h starts at column 3U
              umma_tf32_ts(acc + GRU_U, al, bh, id2, 1u);
              umma_tf32_ts(acc + GRU_U, ah, bl, id2, 1u);
              umma_tf32_ts(acc + 3 * GRU_U, al, bh + 2 * GRU_U, id1, fresh);
              umma_tf32_ts(acc + 3 * GRU_U, ah, bl + 2 * GRU_U, id1, 1u);
            }
          } else {
            const uint32_t acc = tbase + (uint32_t)((1 + (j & 1)) * GRU_ACC_COLS);   // sets 1 / 2
            const uint32_t fresh = (first && j < 2) ? 0u : 1u;
            if (xph) {
              umma_tf32_ts(acc, ah, bh, id3, fresh);
            } else {
              umma_tf32_ts(acc + GRU_U, ah, bh, id2, 1u);
              umma_tf32_ts(acc + 3 * GRU_U, ah, bh + 2 * GRU_U, id1, fresh);
            }
          }
        }
        umma_commit(a_empty + sa);
        umma_commit(w_empty + sw);
      }
      __syncwarp();
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
  const int64_t total = (int64_t)tiles * n_kb * TS_KCH * GRU_WROWS;
  if (i >= total) return;
  const int r = (int)(i % GRU_WROWS);
  const int kc = (int)((i / GRU_WROWS) % TS_KCH);
  const int kb = (int)((i / (GRU_WROWS * TS_KCH)) % n_kb);
  const int t = (int)(i / ((int64_t)GRU_WROWS * TS_KCH * n_kb));
  const bool xph = kb < nbx;
  const int gi = r / GRU_U, j = t * GRU_U + r % GRU_U;
  const int gate = xph ? (gi == 0 ? 2 : gi - 1) : gi;   // gate order in the weights: r, z, n
  const int kdim = xph ? m_dim : d;
  const float* w = xph ? w_ih : w_hh;
  const int k = (xph ? kb : kb - nbx) * TS_BK + kc * 4;
  const float4 v = umma_load_chunk(j < d ? w + (int64_t)(gate * d + j) * kdim : nullptr, k, kdim, false);
  float4 h, l;
  tf32_split(v, h, l);
  float* stage = out + ((int64_t)t * n_kb + kb) * GRU_W_STAGE_FLOATS;
  *reinterpret_cast<float4*>(stage + (kc * GRU_WROWS + r) * 4) = h;
  *reinterpret_cast<float4*>(stage + TS_KCH * GRU_WROWS * 4 + (kc * GRU_WROWS + r) * 4) = l;
}

extern "C" int64_t tiger_gru_pack_bytes(int m_dim, int d) {
  if (m_dim <= 0 || d <= 0) return -1;
  const int64_t tiles = (d + GRU_U - 1) / GRU_U;
  const int64_t n_kb = (m_dim + TS_BK - 1) / TS_BK + (d + TS_BK - 1) / TS_BK;
  return tiles * n_kb * GRU_W_STAGE_FLOATS * (int64_t)sizeof(float);
}

extern "C" int tiger_gru_pack(const float* w_ih, const float* w_hh, int m_dim, int d, float* out, void* stream) {
  if (w_ih == nullptr || w_hh == nullptr || out == nullptr || tiger_gru_pack_bytes(m_dim, d) < 0 ||
      (((uintptr_t)out) & 15) != 0)
    return TIGER_EINVAL;
  const int tiles = (d + GRU_U - 1) / GRU_U;
  const int nbx = (m_dim + TS_BK - 1) / TS_BK, nbh = (d + TS_BK - 1) / TS_BK;
  const int64_t total = (int64_t)tiles * (nbx + nbh) * TS_KCH * GRU_WROWS;
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
