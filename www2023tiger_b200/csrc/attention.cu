// K3  fused temporal graph attention (n_layers = 1), eval mode.
// Reference: GraphEmbedding.compute_embedding_with_computation_graph
// (tiger/model/temporal_agg_modules.py:29-83) + TemporalAttention.forward (:210-235) + torch
// F.multi_head_attention_forward (need_weights branch) + MergeLayer (basic_modules.py:16-19).
//
// One CTA handles G queries.  Math per query (h = head, hd = E / n_head, scale = sqrt(1/hd)):
//   q    = scale * (Wq [c | cos(b)] + bq)                       c = repr(center) + nf(center)
//   qk_h = Wk_h^T q_h  (length C)      qb_h = q_h . bk_h        <- "folded" key projection
//   s_hj = qk_h . kv_j + qb_h          kv_j = [repr(n_j)+nf | ef(e_j) | cos((t - t_j) w + b)]
//   p_h  = softmax_j(s_hj) over unmasked slots
//   o_h  = Wv_h (sum_j p_hj kv_j) + bv_h                         <- "folded" value projection
//   out  = Wo [o_1 | o_2 ..] + bo ; 0 if every slot is padding
//   z    = W2 relu(W1 [out | c] + b1) + b2
// which equals the reference's  softmax(q K^T) V  exactly (linearity; sum_j p_hj = 1) and
// needs ~9x fewer flops than projecting all K neighbors.
//
// Blackwell mapping: warp 0 is a producer that streams the seven k-major weight matrices
// (2.8 MB, L2 resident) through a 3-stage shared-memory ring with cp.async.bulk (TMA 1-D bulk
// copies, SASS UBLKCP) completing on mbarriers; it runs ahead across operator boundaries, so
// weight latency is hidden behind the gathers and the math.  Consumer warps own output
// columns (thread n -> columns n, n+TL, n+2TL) and keep G accumulators per column in registers.
// Neighbor rows are read straight from the node tables (L2) in two passes (scores, pooling);
// they are not staged, which keeps shared memory at ~110 KB so two CTAs fit per SM.
#include "common.cuh"

#define ATT_NST 3                 // ring stages
#define ATT_STAGE_FLOATS 4096     // 16 KB per stage
#define ATT_MAXH 8

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_sync(int n_threads) {
  asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

// one streamed matrix: `rows` k-rows of `ld` floats, contiguous
struct AttMat {
  const float* ptr;
  int rows;
  int ld;
};

struct AttArgs {
  // gather mode
  const int64_t* center_nids;
  const float* q_ts;
  int64_t ts_period;
  const int64_t* neigh_nids;
  const int64_t* neigh_eids;
  const float* neigh_ts;
  const float* rows_a;
  const float* rows_b;
  const void* sel;
  int sel_is_i64;
  const float* nfeats;
  const float* efeats;
  // dense mode (TemporalAttention.forward signature)
  const float* qx;    // [n,d]
  const float* qt;    // [n,d]
  const float* kx;    // [n,K,d]
  const float* ky;    // [n,K,de]
  const float* kt;    // [n,K,d]
  const uint8_t* pad; // [n,K]
  int dense;
  int64_t n_query;
  int k, d, de, n_head;
  int ldE, ldC, ldD;  // padded leading dims of the packed matrices
  tiger_attn_params p;
  float* out;
};

template <int G>
struct GS_ {
  static const int v = G <= 1 ? 1 : (G <= 2 ? 2 : (G <= 4 ? 4 : 8));
};

template <int G>
__device__ __forceinline__ void load_xg(float (&xv)[G], const float* x) {
  constexpr int S = GS_<G>::v;
  if constexpr (S == 1) {
    xv[0] = x[0];
  } else if constexpr (S == 2) {
    const float2 t = *reinterpret_cast<const float2*>(x);
    xv[0] = t.x;
    if constexpr (G > 1) xv[1] = t.y;
  } else {
    const float4 t = *reinterpret_cast<const float4*>(x);
    xv[0] = t.x; xv[1] = t.y; xv[2] = t.z;
    if constexpr (G > 3) xv[3] = t.w;
    if constexpr (S == 8) {
      const float4 u = *reinterpret_cast<const float4*>(x + 4);
      if constexpr (G > 4) xv[4] = u.x;
      if constexpr (G > 5) xv[5] = u.y;
      if constexpr (G > 6) xv[6] = u.z;
      if constexpr (G > 7) xv[7] = u.w;
    }
  }
}

struct Ring {
  uint64_t* full;
  uint64_t* empty;
  const float* stage;
  uint32_t it;  // chunks consumed so far
};

// y[n*ys_n + g*ys_g] = act(scale * (sum_k W[k][n] x_n[k][g] + bias[n]))  for n < n_out
// W streams through the ring in chunks of `rpc` rows; x_n = x + (n / hd) * x_head_stride.
template <int G, int NI>
__device__ __forceinline__ void matvec_stream(Ring& ring, int ld, int k_in, int n_out, const float* x,
                                              int x_head_stride, int hd, float* y, int ys_n, int ys_g,
                                              const float* __restrict__ bias, float scale, bool relu, int gvalid,
                                              int ctid, int tl) {
  constexpr int S = GS_<G>::v;
  int n[NI];
  bool ok[NI];
  const float* xn[NI];
  float acc[NI][G];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    n[i] = ctid + i * tl;
    ok[i] = n[i] < n_out;
    if (!ok[i]) n[i] = n_out - 1;
    xn[i] = x + (x_head_stride ? (n[i] / hd) * x_head_stride : 0);
#pragma unroll
    for (int g = 0; g < G; ++g) acc[i][g] = 0.f;
  }
  const int rpc = ATT_STAGE_FLOATS / ld;
  const int lane = ctid & 31;
  for (int k0 = 0; k0 < k_in; k0 += rpc) {
    const int rows = (k_in - k0) < rpc ? (k_in - k0) : rpc;
    const int s = ring.it % ATT_NST;
    mbar_wait(ring.full + s, (ring.it / ATT_NST) & 1);
    const float* ws = ring.stage + s * ATT_STAGE_FLOATS;
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
      float w[NI];
#pragma unroll
      for (int i = 0; i < NI; ++i) w[i] = ws[r * ld + n[i]];
      if (x_head_stride == 0) {
        float xv[G];
        load_xg<G>(xv, xn[0] + (k0 + r) * S);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
          for (int g = 0; g < G; ++g) acc[i][g] = fmaf(w[i], xv[g], acc[i][g]);
      } else {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          float xv[G];
          load_xg<G>(xv, xn[i] + (k0 + r) * S);
#pragma unroll
          for (int g = 0; g < G; ++g) acc[i][g] = fmaf(w[i], xv[g], acc[i][g]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ring.empty + s);
    ++ring.it;
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    if (!ok[i]) continue;
    const float b = bias != nullptr ? bias[n[i]] : 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (g >= gvalid) break;
      float v = (acc[i][g] + b) * scale;
      if (relu) v = fmaxf(v, 0.f);
      y[n[i] * ys_n + g * ys_g] = v;
    }
  }
}

template <int G>
__device__ __forceinline__ void matvec_dispatch(Ring& ring, int ld, int k_in, int n_out, const float* x,
                                                int x_head_stride, int hd, float* y, int ys_n, int ys_g,
                                                const float* bias, float scale, bool relu, int gvalid, int ctid,
                                                int tl) {
  const int ni = (n_out + tl - 1) / tl;
  if (ni <= 1)
    matvec_stream<G, 1>(ring, ld, k_in, n_out, x, x_head_stride, hd, y, ys_n, ys_g, bias, scale, relu, gvalid, ctid, tl);
  else if (ni == 2)
    matvec_stream<G, 2>(ring, ld, k_in, n_out, x, x_head_stride, hd, y, ys_n, ys_g, bias, scale, relu, gvalid, ctid, tl);
  else
    matvec_stream<G, 3>(ring, ld, k_in, n_out, x, x_head_stride, hd, y, ys_n, ys_g, bias, scale, relu, gvalid, ctid, tl);
}

// producer: stream `rows` x `ld` floats in ring-sized chunks
__device__ __forceinline__ void produce_matrix(const AttMat& m, uint64_t* full, uint64_t* empty, float* stage,
                                               uint32_t& it) {
  const int rpc = ATT_STAGE_FLOATS / m.ld;
  for (int k0 = 0; k0 < m.rows; k0 += rpc) {
    const int rows = (m.rows - k0) < rpc ? (m.rows - k0) : rpc;
    const int s = it % ATT_NST;
    if (it >= ATT_NST) mbar_wait(empty + s, ((it / ATT_NST) - 1) & 1);
    const uint32_t bytes = (uint32_t)rows * (uint32_t)m.ld * 4u;
    mbar_arrive_expect_tx(full + s, bytes);
    bulk_load(stage + s * ATT_STAGE_FLOATS, m.ptr + (int64_t)k0 * m.ld, bytes, full + s);
    ++it;
  }
}

__device__ __forceinline__ const float* resolve_row(const AttArgs& a, int64_t u) {
  int64_t r;
  if (a.sel_is_i64)
    r = reinterpret_cast<const int64_t*>(a.sel)[u];
  else
    r = reinterpret_cast<const int32_t*>(a.sel)[u];
  if (a.rows_a == nullptr || r >= 0) return a.rows_b + r * a.d;
  return a.rows_a + u * a.d;
}

// shared-memory layout in floats; every region is rounded up to 4 floats (16 bytes)
struct AttSmem {
  int stage, xq, q, qk, kvbar, mo, hdn, sc, qb, tq, nts, ctr, nbr, eid, full, empty, invalid, total;
};

template <int G>
__host__ __device__ inline AttSmem att_layout(int d, int de, int K, int H) {
  const int S = GS_<G>::v;
  const int E = 2 * d, C = 2 * d + de;
  AttSmem L;
  int off = 0;
  auto take = [&off](int n) { const int o = off; off += (n + 3) & ~3; return o; };
  L.stage = take(ATT_NST * ATT_STAGE_FLOATS);
  L.xq = take(E * S);
  L.q = take(E * S);
  L.qk = take(H * G * C);
  L.kvbar = take(H * C * S);
  L.mo = take((E + d) * S);
  L.hdn = take(d * S);
  L.sc = take(G * H * K);
  L.qb = take(H * G);
  L.tq = take(G);
  L.nts = take(G * K);
  L.ctr = take(2 * G);
  L.nbr = take(2 * G * K);
  L.eid = take(2 * G * K);
  L.full = take(2 * ATT_NST);
  L.empty = take(2 * ATT_NST);
  L.invalid = take(G);
  L.total = off;
  return L;
}

template <int G>
__global__ void __launch_bounds__(544)
temporal_attention_kernel(const AttArgs a) {
  constexpr int S = GS_<G>::v;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = a.d, de = a.de, K = a.k, H = a.n_head;
  const int E = 2 * d, C = 2 * d + de, hd = E / H;
  const int tl = blockDim.x - 32;
  const int n_cwarps = tl >> 5;
  // ---- shared memory carve-up (every region starts on a 16-byte boundary) ----
  const AttSmem L = att_layout<G>(d, de, K, H);
  float* base = reinterpret_cast<float*>(smem_raw);
  float* stage = base + L.stage;     // NST * STAGE
  float* xq = base + L.xq;           // [E][S]   later reused for o
  float* q = base + L.q;             // [E][S]
  float* qk = base + L.qk;           // [H][G][C]
  float* kvbar = base + L.kvbar;     // [H][C][S]
  float* mo = base + L.mo;           // [E+d][S]
  float* hdn = base + L.hdn;         // [d][S]
  float* sc = base + L.sc;           // [G][H][K]
  float* qb = base + L.qb;           // [H][G]
  float* tq = base + L.tq;           // [G]
  float* nts = base + L.nts;         // [G][K]
  int64_t* ctr = reinterpret_cast<int64_t*>(base + L.ctr);   // [G]
  int64_t* nbr = reinterpret_cast<int64_t*>(base + L.nbr);   // [G][K]
  int64_t* eid = reinterpret_cast<int64_t*>(base + L.eid);   // [G][K]
  uint64_t* full = reinterpret_cast<uint64_t*>(base + L.full);    // [NST]
  uint64_t* empty = reinterpret_cast<uint64_t*>(base + L.empty);  // [NST]
  int* invalid = reinterpret_cast<int*>(base + L.invalid);        // [G]

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < ATT_NST; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, n_cwarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t q0 = (int64_t)blockIdx.x * G;
  const int gvalid = (int)((a.n_query - q0) < G ? (a.n_query - q0) : G);

  if (tid < 32) {
    // =========================== producer warp ===========================
    if (tid == 0) {
      uint32_t it = 0;
      AttMat m;
      m = {a.p.wqT, E, a.ldE};  produce_matrix(m, full, empty, stage, it);
      for (int h = 0; h < H; ++h) {
        m = {a.p.wk + (int64_t)h * hd * a.ldC, hd, a.ldC};
        produce_matrix(m, full, empty, stage, it);
      }
      m = {a.p.wvT, C, a.ldE};  produce_matrix(m, full, empty, stage, it);
      m = {a.p.woT, E, a.ldE};  produce_matrix(m, full, empty, stage, it);
      m = {a.p.fc1T, E + d, a.ldD};  produce_matrix(m, full, empty, stage, it);
      m = {a.p.fc2T, d, a.ldD};  produce_matrix(m, full, empty, stage, it);
    }
    return;
  }
  // =========================== consumer warps ===========================
  const int ctid = tid - 32;
  const int lane = ctid & 31, cw = ctid >> 5;
  Ring ring = {full, empty, stage, 0u};
  const float scale = sqrtf(1.0f / (float)hd);

  // ---- P1: metadata, center rows, query time code ----
  for (int i = ctid; i < G * K; i += tl) {
    const int g = i / K, j = i % K;
    int64_t nb = 0, ei = 0;
    float tv = 0.f;
    if (g < gvalid) {
      const int64_t o = (q0 + g) * K + j;
      if (a.dense) {
        nb = a.pad[o] ? 0 : 1;   // only the mask matters in dense mode
      } else {
        nb = a.neigh_nids[o];
        ei = a.neigh_eids[o];
        tv = a.neigh_ts[o];
      }
    }
    nbr[i] = nb; eid[i] = ei; nts[i] = tv;
  }
  if (ctid < G) {
    ctr[ctid] = (ctid < gvalid && !a.dense) ? a.center_nids[q0 + ctid] : 0;
    tq[ctid] = (ctid < gvalid && !a.dense) ? a.q_ts[(q0 + ctid) % a.ts_period] : 0.f;
  }
  consumer_sync(tl);
  if (ctid < G) {
    int any = 0;
    for (int j = 0; j < K; ++j) any |= (nbr[ctid * K + j] != 0);
    invalid[ctid] = !any;
  }
  for (int g = cw; g < G; g += n_cwarps) {
    if (g < gvalid) {
      const float* rp;
      const float* nfp = nullptr;
      const float* qtp = nullptr;
      if (a.dense) {
        rp = a.qx + (q0 + g) * d;
        qtp = a.qt + (q0 + g) * d;
      } else {
        const int64_t u = ctr[g];
        rp = resolve_row(a, u);
        if (a.nfeats != nullptr) nfp = a.nfeats + u * d;
      }
      for (int c = lane; c < d; c += 32) {
        const float v = rp[c] + (nfp != nullptr ? nfp[c] : 0.f);
        xq[c * S + g] = v;
        mo[(E + c) * S + g] = v;
        xq[(d + c) * S + g] = qtp != nullptr ? qtp[c] : time_enc(0.f, a.p.time_w[c], a.p.time_b[c]);
      }
    } else {
      for (int c = lane; c < d; c += 32) {
        xq[c * S + g] = 0.f;
        mo[(E + c) * S + g] = 0.f;
        xq[(d + c) * S + g] = 0.f;
      }
    }
  }
  consumer_sync(tl);

  // ---- P2: q = scale * (Wq xq + bq) ----
  matvec_dispatch<G>(ring, a.ldE, E, E, xq, 0, hd, q, S, 1, a.p.in_bias, scale, false, G, ctid, tl);
  consumer_sync(tl);

  // ---- P3: qk_h = Wk_h^T q_h, qb_h = q_h . bk_h ----
  for (int h = 0; h < H; ++h)
    matvec_dispatch<G>(ring, a.ldC, hd, C, q + h * hd * S, 0, hd, qk + h * G * C, 1, C, nullptr, 1.0f, false, G, ctid,
                       tl);
  for (int i = cw; i < H * G; i += n_cwarps) {
    const int h = i / G, g = i % G;
    float s = 0.f;
    for (int c = lane; c < hd; c += 32) s += q[(h * hd + c) * S + g] * a.p.in_bias[E + h * hd + c];
    s = warp_sum(s);
    if (lane == 0) qb[h * G + g] = s;
  }
  consumer_sync(tl);

  // ---- P4: scores, one warp per (query, slot) ----
  for (int i = cw; i < G * K; i += n_cwarps) {
    const int g = i / K, j = i % K;
    float s[ATT_MAXH];
#pragma unroll
    for (int h = 0; h < ATT_MAXH; ++h) s[h] = 0.f;
    const bool live = g < gvalid && nbr[i] != 0;
    if (live) {
      const float* qkg = qk + g * C;
      if (a.dense) {
        const int64_t o = (q0 + g) * K + j;
        const float* px = a.kx + o * d;
        const float* py = a.ky + o * de;
        const float* pt = a.kt + o * d;
        for (int c = lane; c < d; c += 32) {
          const float v = px[c];
#pragma unroll
          for (int h = 0; h < ATT_MAXH; ++h) if (h < H) s[h] = fmaf(qkg[h * G * C + c], v, s[h]);
        }
        for (int c = lane; c < de; c += 32) {
          const float v = py[c];
#pragma unroll
          for (int h = 0; h < ATT_MAXH; ++h) if (h < H) s[h] = fmaf(qkg[h * G * C + d + c], v, s[h]);
        }
        for (int c = lane; c < d; c += 32) {
          const float v = pt[c];
#pragma unroll
          for (int h = 0; h < ATT_MAXH; ++h) if (h < H) s[h] = fmaf(qkg[h * G * C + d + de + c], v, s[h]);
        }
      } else {
        const int64_t u = nbr[i];
        const float* rp = resolve_row(a, u);
        const float* nfp = a.nfeats != nullptr ? a.nfeats + u * d : nullptr;
        for (int c = lane; c < d; c += 32) {
          const float v = rp[c] + (nfp != nullptr ? nfp[c] : 0.f);
#pragma unroll
          for (int h = 0; h < ATT_MAXH; ++h) if (h < H) s[h] = fmaf(qkg[h * G * C + c], v, s[h]);
        }
        if (a.efeats != nullptr) {
          const float* ep = a.efeats + eid[i] * de;
          for (int c = lane; c < de; c += 32) {
            const float v = ep[c];
#pragma unroll
            for (int h = 0; h < ATT_MAXH; ++h) if (h < H) s[h] = fmaf(qkg[h * G * C + d + c], v, s[h]);
          }
        }
        const float dt = tq[g] - nts[i];
        for (int c = lane; c < d; c += 32) {
          const float v = time_enc(dt, a.p.time_w[c], a.p.time_b[c]);
#pragma unroll
          for (int h = 0; h < ATT_MAXH; ++h) if (h < H) s[h] = fmaf(qkg[h * G * C + d + de + c], v, s[h]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < ATT_MAXH; ++h) {
      if (h < H) {
        const float t = warp_sum(s[h]);
        if (lane == 0) sc[(g * H + h) * K + j] = live ? t + qb[h * G + g] : -INFINITY;
      }
    }
  }
  consumer_sync(tl);

  // ---- P5: masked softmax over the K slots ----
  if (ctid < G * H) {
    float* row = sc + ctid * K;
    float m = -INFINITY;
    for (int j = 0; j < K; ++j) m = fmaxf(m, row[j]);
    if (m == -INFINITY) {
      for (int j = 0; j < K; ++j) row[j] = 0.f;   // every slot is padding: output is zero-filled below
    } else {
      float sum = 0.f;
      for (int j = 0; j < K; ++j) {
        const float e = expf(row[j] - m);
        row[j] = e;
        sum += e;
      }
      for (int j = 0; j < K; ++j) row[j] = row[j] / sum;
    }
  }
  consumer_sync(tl);

  // ---- P6: pooled keys kvbar[h][c][g] = sum_j p[g][h][j] kv[g][j][c] ----
  for (int i = ctid; i < G * C; i += tl) {
    const int g = i / C, c = i % C;
    float acc[ATT_MAXH];
#pragma unroll
    for (int h = 0; h < ATT_MAXH; ++h) acc[h] = 0.f;
    if (g < gvalid) {
      for (int j = 0; j < K; ++j) {
        const int gi = g * K + j;
        if (nbr[gi] == 0) continue;
        float v;
        if (a.dense) {
          const int64_t o = (q0 + g) * K + j;
          v = c < d ? a.kx[o * d + c] : (c < d + de ? a.ky[o * de + (c - d)] : a.kt[o * d + (c - d - de)]);
        } else if (c < d) {
          const int64_t u = nbr[gi];
          v = resolve_row(a, u)[c] + (a.nfeats != nullptr ? a.nfeats[u * d + c] : 0.f);
        } else if (c < d + de) {
          v = a.efeats != nullptr ? a.efeats[eid[gi] * de + (c - d)] : 0.f;
        } else {
          v = time_enc(tq[g] - nts[gi], a.p.time_w[c - d - de], a.p.time_b[c - d - de]);
        }
#pragma unroll
        for (int h = 0; h < ATT_MAXH; ++h) if (h < H) acc[h] = fmaf(sc[(g * H + h) * K + j], v, acc[h]);
      }
    }
#pragma unroll
    for (int h = 0; h < ATT_MAXH; ++h) if (h < H) kvbar[(h * C + c) * S + g] = acc[h];
  }
  consumer_sync(tl);

  // ---- P7: o = Wv kvbar_h(n) + bv  (into xq) ----
  float* o = xq;
  matvec_dispatch<G>(ring, a.ldE, C, E, kvbar, C * S, hd, o, S, 1, a.p.in_bias + 2 * E, 1.0f, false, G, ctid, tl);
  consumer_sync(tl);

  // ---- P8: out = Wo o + bo, zero for all-padding rows (into mo[0:E]) ----
  matvec_dispatch<G>(ring, a.ldE, E, E, o, 0, hd, mo, S, 1, a.p.out_bias, 1.0f, false, G, ctid, tl);
  consumer_sync(tl);
  for (int i = ctid; i < E * G; i += tl) {
    const int g = i % G, n = i / G;
    if (invalid[g]) mo[n * S + g] = 0.f;
  }
  consumer_sync(tl);

  // ---- P9: hidden = relu(W1 [out | c] + b1) ----
  matvec_dispatch<G>(ring, a.ldD, E + d, d, mo, 0, hd, hdn, S, 1, a.p.fc1_b, 1.0f, true, G, ctid, tl);
  consumer_sync(tl);

  // ---- P10: z = W2 hidden + b2 -> global ----
  matvec_dispatch<G>(ring, a.ldD, d, d, hdn, 0, hd, a.out + q0 * d, 1, d, a.p.fc2_b, 1.0f, false, gvalid, ctid, tl);
}

template <int G>
static size_t att_smem_bytes(int d, int de, int K, int H) {
  return (size_t)att_layout<G>(d, de, K, H).total * sizeof(float);
}

static int g_num_sms = 0;

template <int G>
static int launch_attention(const AttArgs& a, cudaStream_t st) {
  const int C = 2 * a.d + a.de, E = 2 * a.d;
  const int mx = C > E ? C : E;
  int tl = ((mx + 2) / 3 + 31) / 32 * 32;
  if (tl < 64) tl = 64;
  if (tl > 512) return TIGER_EINVAL;
  const size_t smem = att_smem_bytes<G>(a.d, a.de, a.k, a.n_head);
  if (smem > 227 * 1024) return TIGER_EINVAL;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(temporal_attention_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return TIGER_ECUDA;
    configured = smem;
  }
  const unsigned grid = (unsigned)((a.n_query + G - 1) / G);
  temporal_attention_kernel<G><<<grid, tl + 32, smem, st>>>(a);
  return tiger_launch_status();
}

static int attention_entry(AttArgs& a, cudaStream_t st) {
  if (a.n_query < 0 || a.k <= 0 || a.d <= 0 || a.de <= 0 || a.n_head <= 0 || a.n_head > ATT_MAXH) return TIGER_EINVAL;
  if ((2 * a.d) % a.n_head != 0) return TIGER_EINVAL;
  if (a.n_query == 0) return TIGER_OK;
  const int E = 2 * a.d, C = 2 * a.d + a.de;
  a.ldE = (E + 3) / 4 * 4;
  a.ldC = (C + 3) / 4 * 4;
  a.ldD = (a.d + 3) / 4 * 4;
  if (a.ldC > ATT_STAGE_FLOATS) return TIGER_EINVAL;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  // queries per CTA: 4 keeps two CTAs per SM resident at d = 172; fewer for tiny launches
  if (a.n_query <= g_num_sms) return launch_attention<1>(a, st);
  if (a.n_query <= 2 * g_num_sms) return launch_attention<2>(a, st);
  return launch_attention<4>(a, st);
}

extern "C" int tiger_temporal_attention(const int64_t* center_nids, const float* q_ts, int64_t n_query,
                                        int64_t ts_period, const int64_t* neigh_nids, const int64_t* neigh_eids,
                                        const float* neigh_ts, int k, const float* rows_a, const float* rows_b,
                                        const void* sel, int sel_is_i64, const float* nfeats, const float* efeats,
                                        int d, int de, int n_head, const tiger_attn_params* params, float* out,
                                        void* stream) {
  if (params == nullptr || rows_b == nullptr || sel == nullptr) return TIGER_EINVAL;
  AttArgs a = {};
  a.center_nids = center_nids; a.q_ts = q_ts; a.ts_period = ts_period > 0 ? ts_period : n_query;
  a.neigh_nids = neigh_nids; a.neigh_eids = neigh_eids; a.neigh_ts = neigh_ts;
  a.rows_a = rows_a; a.rows_b = rows_b; a.sel = sel; a.sel_is_i64 = sel_is_i64;
  a.nfeats = nfeats; a.efeats = efeats; a.dense = 0;
  a.n_query = n_query; a.k = k; a.d = d; a.de = de; a.n_head = n_head; a.p = *params; a.out = out;
  return attention_entry(a, as_stream(stream));
}

extern "C" int tiger_temporal_attention_dense(const float* qx, const float* qt, const float* kx, const float* ky,
                                              const float* kt, const uint8_t* padding_mask, int64_t n_query, int k,
                                              int d, int de, int n_head, const tiger_attn_params* params, float* out,
                                              void* stream) {
  if (params == nullptr) return TIGER_EINVAL;
  AttArgs a = {};
  a.qx = qx; a.qt = qt; a.kx = kx; a.ky = ky; a.kt = kt; a.pad = padding_mask; a.dense = 1; a.ts_period = 1;
  a.n_query = n_query; a.k = k; a.d = d; a.de = de; a.n_head = n_head; a.p = *params; a.out = out;
  return attention_entry(a, as_stream(stream));
}
