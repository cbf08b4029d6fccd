// Fused spectral graph-convolution layer (reference: model/lanczos_net.py:157-182):
//     msg = [ V diag(f_s) V^T X  (s < S) ] ++ [ L_e X  (e <= E) ];   X' = ReLU(cat(msg) W^T + b)
// as ONE persistent tcgen05 kernel (skeleton: tc_gemm.cuh).  Nothing of the reference's
// intermediate tensors exists in HBM: not the N x N filters, not the [B*N, C*D] message matrix.
//
// Tiles are PACKED: lnb_graph_prepare measures every graph's real extent (n_eff rows/columns
// of the operators that are not identically zero, k_eff non-zero Ritz vectors) and assigns
// consecutive graphs to 128-row tiles by next-fit (sum n_eff <= 128, sum ceil4(k_eff) <= 128,
// <= 32 graphs).  A QM8-shaped batch of 1024 molecules (16 real atoms on average, padded to 26)
// becomes ~137 tiles -- one wave of the 148 SMs -- instead of 256 fixed-slot tiles.  Rows that
// are pure padding are never multiplied; their (constant) output act(b) is written directly.
// Dropping exact zeros is exact, so the result equals the dense reference for arbitrary inputs.
//
// Two accumulator lifetimes ("steps") per tile:
//   step 0  Z = sum_s (f_s . U) W_s^T      rows = (graph, Ritz index): the producer only scales
//           rows of U_g = V_g^T X_g (shared memory) by the filter coefficient; Z is drained to
//           shared memory;
//   step 1  E = sum_e (L_e X) W_e^T        rows = (graph, node): sparse ELL rows of the operators
//           times X; epilogue: out = act(E + V Z + b), written back as full 512-byte rows.
// Algebra: sum_s V diag(f_s) V^T X W_s^T = V [ sum_s diag(f_s) (V^T X) W_s^T ].
#include "tc_gemm.cuh"

namespace {

constexpr int KMAX = 32;        // max Ritz pairs
constexpr int GMAX = 32;        // max graphs per tile
constexpr int RMAX = 128;       // rows per tile
constexpr int EMAX = 16;        // max operator channels
constexpr int FR = 8;           // filter coefficients held in registers per row

// --------------------------------------------------------------------------------------------
// Per-forward operator compression and extents.
//   ell_val/ell_idx [B, E1, N(t), N(n)]: the t-th non-zero of row n of channel e (t-major so a
//   warp of consecutive rows reads consecutive addresses; the diagonal entry first, then by
//   column); ell_max[b,e] = max non-zeros per row.
//   gext[b] = {n_eff, k_eff}: the operators are zero outside their leading n_eff rows/columns,
//   Q[b] is zero outside its leading n_eff rows / k_eff columns.
// --------------------------------------------------------------------------------------------
template <bool STAGE>   // STAGE: this graph's operators fit in shared memory
__global__ void __launch_bounds__(256)
graph_prepare_kernel(const float* __restrict__ L, const float* __restrict__ Q, int N, int E1, int K,
                     float* __restrict__ ell_val, uint8_t* __restrict__ ell_idx,
                     int32_t* __restrict__ ell_max, int32_t* __restrict__ gext, int binarize) {
  extern __shared__ __align__(16) float gp_smem[];     // [N*N*E1] this graph's operators (optional)
  __shared__ int s_max[EMAX];
  __shared__ int s_ext[2];
  __shared__ uint8_t cnt_s[256 * EMAX];                // non-zeros per (row, channel); N <= 255
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < EMAX) s_max[tid] = 0;
  if (tid < 2) s_ext[tid] = 0;
  const int64_t per = (int64_t)N * N * E1;
  const float* Lg = L + b * per;
  if (STAGE) {                                         // coalesced async copy; strided reads then hit smem
    if ((per & 3) == 0) {
      for (int i = tid; i < (int)(per >> 2); i += 256) tc05::cp_async_16(gp_smem + 4 * i, Lg + 4 * i);
    } else {
      for (int i = tid; i < (int)per; i += 256) tc05::cp_async_4(gp_smem + i, Lg + i);
    }
  }
  const float* Lb = STAGE ? gp_smem : Lg;
  // extents of Q while the operator copy is in flight
  const float* Qb = Q + (int64_t)b * N * K;
  int ne = 0, ke = 0;
  for (int i = tid; i < N * K; i += 256) {
    if (__ldg(Qb + i) != 0.f) {
      ne = max(ne, i / K + 1);
      ke = max(ke, i % K + 1);
    }
  }
  if (STAGE) tc05::cp_async_wait_all();
  __syncthreads();
  const int pairs = N * E1;
  // thread <-> (row n, channel e), p = n*E1 + e: one pass over the row, compacting as it goes
  for (int p = tid; p < pairs; p += 256) {
    const int n = p / E1, e = p - n * E1;
    const float* row = Lb + (n * N) * E1 + e;
    float* val = ell_val + ((int64_t)(b * E1 + e) * N) * N + n;
    uint8_t* idx = ell_idx + ((int64_t)(b * E1 + e) * N) * N + n;
    int cnt = 0, far = 0;
    // the diagonal entry goes first: consecutive rows then gather consecutive rows of X for
    // entry 0 (conflict-free in the fused kernel), the other entries follow in column order
    const float dg = row[n * E1];
    if (dg != 0.f) {
      val[0] = binarize ? 1.f : dg;
      idx[0] = (uint8_t)n;
      cnt = 1;
      far = n + 1;
    }
#pragma unroll 2
    for (int i = 0; i < N; ++i) {
      const float v = row[i * E1];
      if (v != 0.f && i != n) {
        val[cnt * N] = binarize ? 1.f : v;
        idx[cnt * N] = (uint8_t)i;
        ++cnt;
        far = max(far, i + 1);
      }
    }
    cnt_s[p] = (uint8_t)cnt;
    if (cnt) {
      atomicMax(&s_max[e], cnt);
      ne = max(ne, max(n + 1, far));
    }
  }
  if (ne) atomicMax(&s_ext[0], ne);
  if (ke) atomicMax(&s_ext[1], ke);
  __syncthreads();
  // zero-fill the tail of every row up to the channel maximum of this graph, so consumers can
  // run all rows of a (graph, channel) to the same length without per-row guards
  for (int pr = tid; pr < pairs; pr += 256) {
    const int n = pr / E1, e = pr - n * E1;
    float* val = ell_val + ((int64_t)(b * E1 + e) * N) * N + n;
    uint8_t* idx = ell_idx + ((int64_t)(b * E1 + e) * N) * N + n;
    for (int t = cnt_s[pr]; t < s_max[e]; ++t) {
      val[(int64_t)t * N] = 0.f;
      idx[(int64_t)t * N] = 0;
    }
  }
  if (tid < E1) ell_max[b * E1 + tid] = s_max[tid];
  if (tid < 2) gext[b * 2 + tid] = s_ext[tid];
}

// Next-fit assignment of consecutive graphs to tiles.  tiles[0] = T, tiles[1 + t] = first graph
// of tile t, tiles[1 + T] = B.  One CTA, all of it parallel: inclusive prefix sums of the row /
// Ritz-row counts; for every graph i the end NX[i] of the tile that would start at i (a window of
// <= 32 graphs); the tile starts are the graphs reachable from 0 along NX, found by pointer
// jumping (round k marks the starts 2^k .. 2^(k+1)-1 hops away and squares the jump table); a
// prefix sum over the marks numbers the tiles.  The arrays live in shared memory (6 (B+1) ints);
// batches too large for that use the global scratch and a serial walk.
template <bool in_smem>
__global__ void __launch_bounds__(1024)
tile_assign_kernel(const int32_t* __restrict__ gext, int B, int K, int32_t* __restrict__ tiles,
                   int32_t* __restrict__ scratch /* [3 * B] */, int32_t* __restrict__ rowmap,
                   int32_t* __restrict__ nrows) {
  extern __shared__ int32_t ta_smem[];
  __shared__ int warp_n[32], warp_k[32], warp_r[32];
  __shared__ int run_n, run_k, run_r;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t* PN = in_smem ? ta_smem : scratch;                     // inclusive prefix of n_eff
  int32_t* PK = PN + B;                                          // inclusive prefix of ceil4(k_eff)
  int32_t* NX = PK + B;                                          // [B + 1] end of the tile starting at i
  int32_t* PR = NX + 3 * (B + 1);                                // (shared memory only) inclusive prefix of k_eff
  if (tid == 0) { run_n = 0; run_k = 0; run_r = 0; }
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    const int n = (b < B) ? gext[b * 2] : 0;
    const int kr = (b < B) ? min(gext[b * 2 + 1], K) : 0;      // rows of the compact Ritz row list
    const int k = (b < B) ? ((gext[b * 2 + 1] + 3) & ~3) : 0;
    int in = n, ik = k, ir = kr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tn = __shfl_up_sync(0xffffffffu, in, o), tk = __shfl_up_sync(0xffffffffu, ik, o);
      const int tr = __shfl_up_sync(0xffffffffu, ir, o);
      if (lane >= o) { in += tn; ik += tk; ir += tr; }
    }
    if (lane == 31) { warp_n[warp] = in; warp_k[warp] = ik; warp_r[warp] = ir; }
    __syncthreads();
    if (warp == 0) {
      int wn = warp_n[lane], wk = warp_k[lane], wr = warp_r[lane], sn = wn, sk = wk, sr = wr;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tn = __shfl_up_sync(0xffffffffu, sn, o), tk = __shfl_up_sync(0xffffffffu, sk, o);
        const int tr = __shfl_up_sync(0xffffffffu, sr, o);
        if (lane >= o) { sn += tn; sk += tk; sr += tr; }
      }
      warp_n[lane] = sn - wn;
      warp_k[lane] = sk - wk;
      warp_r[lane] = sr - wr;
    }
    __syncthreads();
    if (b < B) {
      PN[b] = run_n + warp_n[warp] + in;
      PK[b] = run_k + warp_k[warp] + ik;
      if (in_smem) {
        PR[b] = run_r + warp_r[warp] + ir;
      } else if (rowmap) {                // {b*K + k : k < k_eff(b)}, see lnb_ritz_rowmap
        const int base = run_r + warp_r[warp] + ir - kr;
        for (int i = 0; i < kr; ++i) rowmap[base + i] = b * K + i;
      }
    }
    __syncthreads();
    if (tid == 1023) { run_n += warp_n[31] + in; run_k += warp_k[31] + ik; run_r += warp_r[31] + ir; }
    __syncthreads();
  }
  if (tid == 0 && nrows) nrows[0] = run_r;
  __threadfence_block();
  if (in_smem && rowmap) {
    // coalesced expansion of the row list: a warp writes the k_eff consecutive entries of a graph
    for (int b = warp; b < B; b += 32) {
      const int base = b ? PR[b - 1] : 0, kr = PR[b] - base;
      for (int i = lane; i < kr; i += 32) rowmap[base + i] = b * K + i;
    }
  }
  for (int i = tid; i < B; i += 1024) {
    const int pn0 = i ? PN[i - 1] : 0, pk0 = i ? PK[i - 1] : 0;
    int j = i + 1;                                   // the first graph always fits (n_eff <= 128)
    const int jmax = min(B, i + GMAX);
    while (j < jmax && PN[j] - pn0 <= RMAX && PK[j] - pk0 <= RMAX) ++j;
    NX[i] = j;
  }
  __syncthreads();
  if (!in_smem) {                                    // huge batch: serial walk over the jump table
    if (tid == 0) {
      int T = 0, i = 0;
      while (i < B) { tiles[1 + T] = i; ++T; i = NX[i]; }
      tiles[0] = T;
      tiles[1 + T] = B;
    }
    return;
  }
  int32_t* Ja = NX;
  int32_t* Jb = NX + (B + 1);
  int32_t* MK = Jb + (B + 1);
  for (int i = tid; i <= B; i += 1024) MK[i] = (i == 0) ? 1 : 0;
  if (tid == 0) { Ja[B] = B; Jb[B] = B; }
  __syncthreads();
  for (int span = 1; span < B; span <<= 1) {
    // starts fewer than `span` hops from graph 0 are marked; Ja = NX applied `span` times
    for (int i = tid; i < B; i += 1024)
      if (MK[i]) MK[Ja[i]] = 1;                      // late marks only add true starts (idempotent)
    for (int i = tid; i < B; i += 1024) Jb[i] = Ja[Ja[i]];
    __syncthreads();
    int32_t* t = Ja; Ja = Jb; Jb = t;
  }
  // number the starts: exclusive prefix sum of the marks
  if (tid == 0) run_n = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    const int m = (b < B) ? MK[b] : 0;
    int in = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tn = __shfl_up_sync(0xffffffffu, in, o);
      if (lane >= o) in += tn;
    }
    if (lane == 31) warp_n[warp] = in;
    __syncthreads();
    if (warp == 0) {
      int wn = warp_n[lane], sn = wn;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tn = __shfl_up_sync(0xffffffffu, sn, o);
        if (lane >= o) sn += tn;
      }
      warp_n[lane] = sn - wn;
    }
    __syncthreads();
    if (m) tiles[1 + run_n + warp_n[warp] + in - 1] = b;
    __syncthreads();
    if (tid == 1023) run_n += warp_n[31] + in;
    __syncthreads();
  }
  if (tid == 0) {
    tiles[0] = run_n;
    tiles[1 + run_n] = B;
  }
}

// --------------------------------------------------------------------------------------------
struct SpectralPolicy {
  static constexpr int kStagesB = 2;      // two W stages: shared memory goes to the packed tile
  static constexpr int LMAX = 8;          // layers run by one launch
  struct Params {
    const float* X;         // [B, N, Din0] input state, or nullptr with node_ids/emb (embedding)
    const int64_t* node_ids;// [B, N]
    const float* emb;       // [emb_rows, Din0]
    const float* Q;         // [B, N, K]
    const float* coeff;     // layer l: coeff + l * coeff_stride -> [B, K, S]
    int64_t coeff_stride;
    const float* ell_val;   // [B, E1, N, N]
    const uint8_t* ell_idx; // [B, E1, N, N]
    const int32_t* ell_max; // [B, E1]
    const int32_t* gext;    // [B, 2]
    const int32_t* tiles;   // [B + 2]
    const float* bias;      // layer l: bias + l * H (may be null)
    float* out;             // [B, N, H] final state (may be null when the readout is fused)
    // fused readout (model/lanczos_net.py:185-194); score == nullptr disables it
    const float* W_out;     // [P, H]
    const float* b_out;     // [P]
    const float* w_att;     // [H]
    const float* b_att;     // [1]
    const uint8_t* mask;    // [B, N] or null (mean over all N nodes)
    float* score;           // [B, P]
    int P, emb_rows;
    int L;                  // number of layers in this launch
    int Din[LMAX];          // input width of each layer (Din[l>0] == H)
    int B, N, E1, K, S, H, relu;
    int LB;                 // ELL lines (channel, t) that fit in shared memory
    int write_pad;          // also write the constant rows of padded nodes of `out`
    int dbg;                // debug experiment flags (LNB_DBG), 0 in production
  };
  // sub = layer * 2 + step  (step 0: Z accumulation, step 1: edge accumulation + epilogue)
  static __device__ __forceinline__ int num_steps(const Params& p, int cta, int ncta) {
    const int T = __ldg(p.tiles);
    const int mine = T > cta ? (T - cta + ncta - 1) / ncta : 0;
    return (p.S > 0 ? 2 : 1) * p.L * mine;
  }
  static __device__ __forceinline__ void decode(const Params& p, int cta, int ncta, int it,
                                                int& m_tile, int& sub) {
    const int per = (p.S > 0 ? 2 : 1) * p.L;
    const int rem = it % per;
    m_tile = cta + (it / per) * ncta;
    sub = p.S > 0 ? rem : rem * 2 + 1;
  }
  static __device__ __forceinline__ int num_kblocks(const Params& p, int sub) {
    return ((sub & 1) == 0 ? p.S : p.E1) * p.Din[sub >> 1] / tcg::BK;
  }
  static __device__ __forceinline__ void w_coords(const Params& p, int sub, int kb, int& col0, int& row0) {
    col0 = ((sub & 1) == 0 ? 0 : p.S * p.Din[sub >> 1]) + kb * tcg::BK;
    row0 = (sub >> 1) * p.H;
  }

  struct Tables {
    int gs, ng, Rtot, Ztot, nquads, nlines;
    int nbase[GMAX + 1], kbase[GMAX + 1], gn[GMAX], gk[GMAX];
    int cnt_e[EMAX], base_e[EMAX], tmax_e[EMAX];
    uint8_t emax[GMAX][EMAX];
    uint8_t row_g[RMAX], row_n[RMAX], z_g[RMAX], z_k[RMAX];
    uint8_t q_g[RMAX / 4 + GMAX], q_n0[RMAX / 4 + GMAX];
    uint8_t line_e[256], line_t[256];
  };

  const Params& p;
  const int tid, r;
  const int N, K, S, E1, H, XP;           // hot parameters in registers
  int Din;                                // input width of the current layer
  float* Xs;                // X rows [RMAX][XP]; reused for V Z + the finished output rows
  float* UZ;                // U rows (graph,k) [RMAX][XP] during step 0, Z afterwards
  float* Qs;                // [RMAX][K]
  Tables* tb;
  float* Ev;                // [LB][RMAX] staged ELL values
  uint8_t* Ei;              // [LB][RMAX] staged ELL columns as tile-local row indices
  float fr[FR];             // this (graph, k) row's filter coefficients f[k, 0..S)
  tcg::PhaseTimer* ptm = nullptr;   // profiling aid: the current step's timer

  static __host__ __device__ constexpr size_t tables_bytes() { return (sizeof(Tables) + 15) & ~size_t(15); }

  __device__ SpectralPolicy(const Params& p_, uint8_t* smem, int tid_)
      : p(p_), tid(tid_), r(tid_ & 127), N(p_.N), K(p_.K), S(p_.S), E1(p_.E1),
        H(p_.H), XP((p_.Din[0] > p_.H ? p_.Din[0] : p_.H) + 4), Din(p_.Din[0]) {
    Xs = reinterpret_cast<float*>(smem);
    UZ = Xs + (size_t)RMAX * XP;
    Qs = UZ + (size_t)RMAX * XP;
    uint8_t* t8 = reinterpret_cast<uint8_t*>(Qs + (size_t)RMAX * K);
    tb = reinterpret_cast<Tables*>(t8);
    Ev = reinterpret_cast<float*>(t8 + tables_bytes());
    Ei = reinterpret_cast<uint8_t*>(Ev + (size_t)p.LB * RMAX);
  }
  static size_t smem_fixed(int Din, int K, int H) {
    const int W = (Din > H ? Din : H) + 4;
    return (size_t)2 * RMAX * W * 4 + (size_t)RMAX * K * 4 + tables_bytes() + 16;
  }
  static size_t ell_line_bytes() { return (size_t)RMAX * 5; }

  // ------------------------------------------------------------------------------------------
  __device__ void build_tables(int m_tile) {
    // executed by warp 0: lane j <-> j-th graph of the tile
    const int lane = tid & 31;
    const int gs = __ldg(p.tiles + 1 + m_tile), ge = __ldg(p.tiles + 2 + m_tile);
    const int ng = ge - gs;
    int n = 0, k = 0;
    if (lane < ng) {
      n = __ldg(p.gext + (gs + lane) * 2);
      k = __ldg(p.gext + (gs + lane) * 2 + 1);
    }
    const int kp = (k + 3) & ~3, nq = (n + 3) >> 2;
    int pn = n, pk = kp, pq = nq;                     // inclusive prefix sums
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, pn, o), bq = __shfl_up_sync(0xffffffffu, pk, o);
      const int c = __shfl_up_sync(0xffffffffu, pq, o);
      if (lane >= o) { pn += a; pk += bq; pq += c; }
    }
    const int nb = pn - n, kb = pk - kp, qb = pq - nq;
    if (lane < ng) {
      tb->nbase[lane] = nb; tb->kbase[lane] = kb; tb->gn[lane] = n; tb->gk[lane] = k;
      for (int i = 0; i < n; ++i) { tb->row_g[nb + i] = (uint8_t)lane; tb->row_n[nb + i] = (uint8_t)i; }
      for (int i = 0; i < kp; ++i) { tb->z_g[kb + i] = (uint8_t)lane; tb->z_k[kb + i] = (uint8_t)i; }
      for (int i = 0; i < nq; ++i) { tb->q_g[qb + i] = (uint8_t)lane; tb->q_n0[qb + i] = (uint8_t)(4 * i); }
    }
    const int Rtot = __shfl_sync(0xffffffffu, pn, 31), Ztot = __shfl_sync(0xffffffffu, pk, 31);
    const int nquads = __shfl_sync(0xffffffffu, pq, 31);
    // per-channel maximum row length over the tile's graphs, per-graph row lengths
    // (all loads issued before the first reduction: one global round trip instead of E1)
    int em[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e)
      em[e] = (e < E1 && lane < ng) ? __ldg(p.ell_max + (gs + lane) * E1 + e) : 0;
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (e < E1) {
        int m = em[e];
        if (lane < ng) tb->emax[lane][e] = (uint8_t)m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) tb->tmax_e[e] = m;
      }
    }
    __syncwarp();
    if (lane == 0) {
      tb->gs = gs; tb->ng = ng; tb->Rtot = Rtot; tb->Ztot = Ztot; tb->nquads = nquads;
      tb->nbase[ng] = Rtot; tb->kbase[ng] = Ztot;
      int left = p.LB, base = 0;
      for (int e = 0; e < E1; ++e) {            // staged lines per channel, in channel order
        const int c = min(tb->tmax_e[e], left);
        tb->cnt_e[e] = c; tb->base_e[e] = base;
        for (int t = 0; t < c; ++t) { tb->line_e[base + t] = (uint8_t)e; tb->line_t[base + t] = (uint8_t)t; }
        base += c; left -= c;
      }
      tb->nlines = base;
    }
  }

  __device__ void step_begin(int m_tile, int sub, int /*kb_first*/, tcg::PhaseTimer& tm) {
    tcg::producers_sync();              // previous step's smem readers / writers are done
    const int layer = sub >> 1, step = sub & 1;
    if (step == 1 && S > 0) return;     // tile state was staged by step 0 of this layer
    Din = p.Din[layer];
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = tcg::PRODUCER_THREADS / 32;
    const int dv = Din / 4;
    ptm = &tm;
    if (layer == 0) {
    if (warp == 0) build_tables(m_tile);
    tcg::producers_sync();
    tm.lap(16);
    const int gs = tb->gs, Rtot = tb->Rtot;
    // ---- phase A: asynchronous copies of the real rows of X and Q (one warp per row) --------
    int64_t my_id = 0;                           // lane j: embedding id of this warp's j-th row
    if (!p.X && warp + lane * NW < Rtot) {
      const int row = warp + lane * NW;
      my_id = __ldg(p.node_ids + (int64_t)(gs + tb->row_g[row]) * N + tb->row_n[row]);
    }
    for (int row = warp, j = 0; row < Rtot; row += NW, ++j) {
      const int64_t src_row = (int64_t)(gs + tb->row_g[row]) * N + tb->row_n[row];
      const float* xsrc;
      if (p.X) {
        xsrc = p.X + src_row * Din;
      } else {                                  // embedding rows (model/lanczos_net.py:154)
        int64_t id = __shfl_sync(0xffffffffu, my_id, j);
        id = id < 0 ? 0 : (id >= p.emb_rows ? p.emb_rows - 1 : id);
        xsrc = p.emb + id * Din;
      }
      float* xd = Xs + (size_t)row * XP;
      for (int q4 = lane; q4 < dv; q4 += 32) tc05::cp_async_16(xd + 4 * q4, xsrc + 4 * q4);
      const float* qsrc = p.Q + src_row * K;            // K % 4 == 0: 16-byte pieces
      float* qd = Qs + (size_t)row * K;
      for (int k4 = lane; k4 < (K >> 2); k4 += 32) tc05::cp_async_16(qd + 4 * k4, qsrc + 4 * k4);
    }
    tm.lap(17);
    // ---- staged ELL lines: line l <-> (channel e, entry t); a warp per line, batched loads ---
    {
      constexpr int ELL_BATCH = 4;
      const int nlines = tb->nlines;
      for (int base = warp; base < nlines; base += NW * ELL_BATCH) {
        for (int r0 = 0; r0 < Rtot; r0 += 32) {
          const int rr = r0 + lane;
          float vv[ELL_BATCH];
          int ii[ELL_BATCH];
#pragma unroll
          for (int u = 0; u < ELL_BATCH; ++u) {
            const int line = base + u * NW;
            vv[u] = 0.f; ii[u] = 0;
            if (line < nlines && rr < Rtot) {
              const int e = tb->line_e[line], t = tb->line_t[line];
              const int g = tb->row_g[rr], n = tb->row_n[rr];
              if (t < tb->emax[g][e]) {
                const int64_t off = (((int64_t)(gs + g) * E1 + e) * N + t) * N + n;
                vv[u] = __ldg(p.ell_val + off);
                ii[u] = tb->nbase[g] + __ldg(p.ell_idx + off);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < ELL_BATCH; ++u) {
            const int line = base + u * NW;
            if (line < nlines && rr < Rtot) {
              Ev[(size_t)line * RMAX + rr] = vv[u];
              Ei[(size_t)line * RMAX + rr] = (uint8_t)ii[u];
            }
          }
        }
      }
    }
    tm.lap(18);
    }  // layer == 0: tile state staged once, reused by every layer
    const int gs = tb->gs, Ztot = tb->Ztot;
    // this thread's (graph, k) row: filter coefficients of this layer into registers
#pragma unroll
    for (int i = 0; i < FR; ++i) fr[i] = 0.f;
    if (S > 0 && r < Ztot) {
      const int g = tb->z_g[r], k = tb->z_k[r];
      if (k < tb->gk[g]) {                       // rows beyond k_eff multiply zero rows of U
        const float* f = p.coeff + layer * p.coeff_stride + ((int64_t)(gs + g) * K + k) * S;
#pragma unroll
        for (int i = 0; i < FR; ++i)
          if (i < S) fr[i] = __ldg(f + i);
      }
    }
    tm.lap(0);
    tc05::cp_async_wait_all();
    tcg::producers_sync();
    tm.lap(1);
    if (S == 0) return;
    // ---- phase B: U_g = Q_g^T X_g in 4 x 4 register tiles over (graph,k) rows x columns -----
    for (int task = tid; task < (Ztot >> 2) * dv; task += tcg::PRODUCER_THREADS) {
      const int dq = task % dv, zq = task / dv;
      const int g = tb->z_g[4 * zq], k0 = tb->z_k[4 * zq];
      const int n_g = tb->gn[g], nb = tb->nbase[g];
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      const float* xs = Xs + (size_t)nb * XP + 4 * dq;
      const float* qs = Qs + (size_t)nb * K + k0;
#pragma unroll 4
      for (int nn = 0; nn < n_g; ++nn) {
        const float4 x4 = *reinterpret_cast<const float4*>(xs + (size_t)nn * XP);
        const float4 q4 = *reinterpret_cast<const float4*>(qs + (size_t)nn * K);
        const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(qv[i], xv[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(UZ + (size_t)(4 * zq + i) * XP + 4 * dq) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    tcg::producers_sync();
    tm.lap(2);
  }

  __device__ __forceinline__ void produce(int sub, int kb, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    const int j0 = kb * tcg::BK;
    const int c = j0 / Din, d0 = j0 - c * Din;
    if ((sub & 1) == 0) {
      // row = (graph, Ritz index): f[k, s] * U[row, d0:d0+32]
      if (r >= tb->Ztot) return;
      float f = 0.f;
      if (S <= FR) {
#pragma unroll
        for (int i = 0; i < FR; ++i) f = (i == c) ? fr[i] : f;
      } else {
        const int g = tb->z_g[r], k = tb->z_k[r];
        f = (k < tb->gk[g]) ? __ldg(p.coeff + (sub >> 1) * p.coeff_stride + ((int64_t)(tb->gs + g) * K + k) * S + c) : 0.f;
      }
      const float4* u4 = reinterpret_cast<const float4*>(UZ + (size_t)r * XP + d0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = u4[q];
        v[4 * q + 0] = f * t.x; v[4 * q + 1] = f * t.y; v[4 * q + 2] = f * t.z; v[4 * q + 3] = f * t.w;
      }
      return;
    }
    // edge type e = c: sparse row of L_e (ELL) times X[:, d0:d0+32]; row = (graph, node)
    if (r >= tb->Rtot) return;
    const int e = c;
    const float* xs = Xs + d0;
    const int ts = tb->cnt_e[e], tmax = tb->tmax_e[e];
    const float* ev = Ev + (size_t)tb->base_e[e] * RMAX + r;
    const uint8_t* ei = Ei + (size_t)tb->base_e[e] * RMAX + r;
#pragma unroll 2
    for (int t = 0; t < ts; ++t) {
      const float a = ev[t * RMAX];
      if (a == 0.f) continue;                        // zero fill up to the tile's longest row
      const int i = ei[t * RMAX];
      const float4* x4 = reinterpret_cast<const float4*>(xs + (size_t)i * XP);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 tt = x4[q];
        v[4 * q + 0] = fmaf(a, tt.x, v[4 * q + 0]);
        v[4 * q + 1] = fmaf(a, tt.y, v[4 * q + 1]);
        v[4 * q + 2] = fmaf(a, tt.z, v[4 * q + 2]);
        v[4 * q + 3] = fmaf(a, tt.w, v[4 * q + 3]);
      }
    }
    if (ts < tmax) {                               // lines that did not fit the staging budget
      const int g = tb->row_g[r], n = tb->row_n[r];
      const int my = tb->emax[g][e], nb = tb->nbase[g];
      const int64_t off0 = (((int64_t)(tb->gs + g) * E1 + e) * N) * N + n;
      for (int t = ts; t < tmax; ++t) {
        float a = 0.f;
        int i = 0;
        if (t < my) {
          a = __ldg(p.ell_val + off0 + (int64_t)t * N);
          i = nb + __ldg(p.ell_idx + off0 + (int64_t)t * N);
        }
        if (a == 0.f) continue;
        const float4* x4 = reinterpret_cast<const float4*>(xs + (size_t)i * XP);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 tt = x4[q];
          v[4 * q + 0] = fmaf(a, tt.x, v[4 * q + 0]);
          v[4 * q + 1] = fmaf(a, tt.y, v[4 * q + 1]);
          v[4 * q + 2] = fmaf(a, tt.z, v[4 * q + 2]);
          v[4 * q + 3] = fmaf(a, tt.w, v[4 * q + 3]);
        }
      }
    }
  }

  // After the last k-block of the edge step: (V Z)[row, :] for every real row in 4 x 4 register
  // tiles into the (now dead) X buffer; overlaps with the tensor core draining its queue.
  __device__ void pre_epilogue(int sub) {
    if ((sub & 1) == 0) return;
    tcg::producers_sync();              // every producer is done reading X
    if (ptm) ptm->lap(22);              // (profiling) wait for the slowest producer group
    if (S == 0) return;
    const int hv = H / 4;
    for (int task = tid; task < tb->nquads * hv; task += tcg::PRODUCER_THREADS) {
      const int hq = task % hv, q = task / hv;
      const int g = tb->q_g[q], n0 = tb->q_n0[q];
      const int n_g = tb->gn[g], k_g = tb->gk[g];
      const int r0 = tb->nbase[g] + n0;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      const float* z = UZ + (size_t)tb->kbase[g] * XP + 4 * hq;
      const float* q0 = Qs + (size_t)r0 * K;
      const int r1 = (n0 + 1 < n_g) ? 1 : 0, r2 = (n0 + 2 < n_g) ? 2 : 0, r3 = (n0 + 3 < n_g) ? 3 : 0;
      // four Ritz indices per trip: Q columns and Z rows in [k_eff, ceil4(k_eff)) are zero
      for (int k = 0; k < k_g; k += 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(q0 + k);
        const float4 a1 = *reinterpret_cast<const float4*>(q0 + r1 * K + k);
        const float4 a2 = *reinterpret_cast<const float4*>(q0 + r2 * K + k);
        const float4 a3 = *reinterpret_cast<const float4*>(q0 + r3 * K + k);
        const float av[4][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w},
                                {a2.x, a2.y, a2.z, a2.w}, {a3.x, a3.y, a3.z, a3.w}};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 z4 = *reinterpret_cast<const float4*>(z + (size_t)(k + kk) * XP);
          const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i][kk], zv[j], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n0 + i < n_g)
          *reinterpret_cast<float4*>(Xs + (size_t)(r0 + i) * XP + 4 * hq) =
              make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    tcg::producers_sync();
  }

  __device__ __forceinline__ void store(int sub, int col, float (&x)[tcg::EW]) {
    if (col >= H) return;
    if ((sub & 1) == 0) {
      // drain Z[row, col:col+32] to shared memory (overwrites U, which is dead by now)
      if (r < tb->Ztot) {
        float4* z4 = reinterpret_cast<float4*>(UZ + (size_t)r * XP + col);
#pragma unroll
        for (int q = 0; q < tcg::EW / 4; ++q)
          z4[q] = make_float4(x[4 * q + 0], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
      }
      return;
    }
    if (r >= tb->Rtot) return;
    float4* o4 = reinterpret_cast<float4*>(Xs + (size_t)r * XP + col);
    const bool relu = p.relu != 0;
    const float* bias = p.bias ? p.bias + (sub >> 1) * H : nullptr;
#pragma unroll
    for (int q = 0; q < tcg::EW / 4; ++q) {
      float y[4] = {x[4 * q + 0], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]};
      if (S > 0) {                                  // + (V Z)[row, col + 4q ..], from pre_epilogue()
        const float4 t = o4[q];
        y[0] += t.x; y[1] += t.y; y[2] += t.z; y[3] += t.w;
      }
      if (col + 4 * q + 3 < H) {
        if (bias) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col) + q);
          y[0] += b4.x; y[1] += b4.y; y[2] += b4.z; y[3] += b4.w;
        }
        if (relu) { y[0] = fmaxf(y[0], 0.f); y[1] = fmaxf(y[1], 0.f); y[2] = fmaxf(y[2], 0.f); y[3] = fmaxf(y[3], 0.f); }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = col + 4 * q + u;
          if (c < H) {
            if (bias) y[u] += __ldg(bias + c);
            if (relu) y[u] = fmaxf(y[u], 0.f);
          }
        }
      }
      o4[q] = make_float4(y[0], y[1], y[2], y[3]);   // finished row chunk = next layer's X row
    }
  }

  // After the last layer: coalesced write-back of the final state (real rows from shared
  // memory, one warp per 512-byte row, plus the constant rows act(b) of padded nodes when
  // requested) and / or the fused readout.  Between layers the state never leaves the SM.
  __device__ void post_epilogue(int sub) {
    if ((sub & 1) == 0 || (sub >> 1) != p.L - 1) return;
    tcg::producers_sync();              // every chunk of every row is in shared memory
    if (ptm) ptm->lap(19);
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = tcg::PRODUCER_THREADS / 32;
    const int hv = H / 4, gs = tb->gs, Rtot = tb->Rtot;
    const float* bias = p.bias ? p.bias + (p.L - 1) * H : nullptr;
    if (p.out) {
      for (int row = warp; row < Rtot; row += NW) {
        const float4* src = reinterpret_cast<const float4*>(Xs + (size_t)row * XP);
        float4* dst = reinterpret_cast<float4*>(
            p.out + ((int64_t)(gs + tb->row_g[row]) * N + tb->row_n[row]) * H);
        for (int q4 = lane; q4 < hv; q4 += 32) dst[q4] = src[q4];
      }
      if (p.write_pad) {
        const int npad = tb->ng * N - Rtot;
        for (int i = warp; i < npad; i += NW) {
          // i-th padded (graph, node) pair of the tile, found by walking the per-graph pad counts
          int g = 0, rem = i;
          while (rem >= N - tb->gn[g]) { rem -= N - tb->gn[g]; ++g; }
          float4* dst = reinterpret_cast<float4*>(p.out + ((int64_t)(gs + g) * N + tb->gn[g] + rem) * H);
          for (int q4 = lane; q4 < hv; q4 += 32) {
            float y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float t = bias ? __ldg(bias + 4 * q4 + u) : 0.f;
              y[u] = (p.relu != 0) ? fmaxf(t, 0.f) : t;
            }
            dst[q4] = make_float4(y[0], y[1], y[2], y[3]);
          }
        }
      }
    }
    if (p.score) readout(bias);
  }

  // Fused readout (model/lanczos_net.py:185-194): y = (W_out x + b_out) * sigmoid(w_att.x + b_att)
  // per node, masked mean over the nodes of each graph.  Rows of padded nodes are the constant
  // act(b_last); they count only where the mask says so (or when there is no mask).
  __device__ void readout(const float* bias_last) {
    const int P = p.P, P1 = p.P + 1, PQ = (P1 + 3) >> 2, HP = H + 4;
    float* Wr = UZ;                                  // [4 PQ][HP]  W_out rows, w_att, zero rows (Z is dead)
    float* Yr = Wr + (size_t)4 * PQ * HP;            // [RMAX + 1][P1]  per-row outputs; last = pad row
    float* cx = Yr + (size_t)(RMAX + 1) * P1 + ((4 - ((RMAX + 1) * P1 & 3)) & 3);   // [H], 16 B aligned
    for (int e = tid; e < 4 * PQ * H; e += tcg::PRODUCER_THREADS) {
      const int o = e / H, h = e - o * H;
      Wr[o * HP + h] = (o < P) ? __ldg(p.W_out + o * H + h) : (o == P ? __ldg(p.w_att + h) : 0.f);
    }
    for (int h = tid; h < H; h += tcg::PRODUCER_THREADS) {
      float t = bias_last ? __ldg(bias_last + h) : 0.f;
      cx[h] = (p.relu != 0) ? fmaxf(t, 0.f) : t;
    }
    uint8_t* mk = reinterpret_cast<uint8_t*>(cx + H);   // [ng][N] node masks of the tile's graphs
    for (int e = tid; e < tb->ng * N; e += tcg::PRODUCER_THREADS)
      mk[e] = p.mask ? __ldg(p.mask + (int64_t)tb->gs * N + e) : (uint8_t)1;
    tcg::producers_sync();
    if (ptm) ptm->lap(20);
    const int Rtot = tb->Rtot;
    // warp <-> (block of 32 rows, third of the outputs): lane = row, so the row loads are
    // conflict-free and every weight load is one broadcast wavefront
    {
      const int warp = tid >> 5, lane = tid & 31;
      const int rb = warp & 3, og = warp >> 2;
      const int per = (P1 + tcg::NGROUPS - 1) / tcg::NGROUPS;
      const int o_end = min(P1, (og + 1) * per);
      const int row = rb * 32 + lane;
      const float4* x4 = reinterpret_cast<const float4*>(Xs + (size_t)(row < Rtot ? row : 0) * XP);
      for (int o0 = og * per; o0 < o_end; o0 += 6) {
        const int cnt = min(6, o_end - o0);
        const float4* w4 = reinterpret_cast<const float4*>(Wr + (size_t)o0 * HP);
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int h4 = 0; h4 < H / 4; ++h4) {
          const float4 x = x4[h4];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            if (j < cnt) {
              const float4 wv = w4[j * (HP / 4) + h4];
              acc[j] = fmaf(x.x, wv.x, acc[j]); acc[j] = fmaf(x.y, wv.y, acc[j]);
              acc[j] = fmaf(x.z, wv.z, acc[j]); acc[j] = fmaf(x.w, wv.w, acc[j]);
            }
          }
        }
        if (row < Rtot) {
#pragma unroll
          for (int j = 0; j < 6; ++j)
            if (j < cnt) {
              const int o = o0 + j;
              const float v = acc[j] + ((o < P) ? __ldg(p.b_out + o) : __ldg(p.b_att));
              Yr[row * P1 + o] = (o < P) ? v : 1.f / (1.f + expf(-v));     // column P: the gate
            }
        }
      }
      // the constant padded-node row: one warp per output, lanes stride the H features
      for (int o = warp; o < P1; o += tcg::PRODUCER_THREADS / 32) {
        float acc = 0.f;
        for (int h = lane; h < H; h += 32) acc = fmaf(cx[h], Wr[o * HP + h], acc);
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
        if (lane == 0) {
          const float v = acc + ((o < P) ? __ldg(p.b_out + o) : __ldg(p.b_att));
          Yr[RMAX * P1 + o] = (o < P) ? v : 1.f / (1.f + expf(-v));
        }
      }
    }
    tcg::producers_sync();
    if (ptm) ptm->lap(21);
    for (int e = tid; e < tb->ng * P; e += tcg::PRODUCER_THREADS) {
      const int g = e / P, o = e - g * P;
      const int nb = tb->nbase[g], n_g = tb->gn[g];
      const uint8_t* m = mk + g * N;
      float acc = 0.f;
      int cnt = 0;
      for (int n = 0; n < N; ++n) {
        if (m[n] == 0) continue;
        const float* y = Yr + (n < n_g ? nb + n : RMAX) * P1;
        acc += y[P] * y[o];
        ++cnt;
      }
      p.score[(int64_t)(tb->gs + g) * P + o] = acc / (float)cnt;
    }
  }
};

}  // namespace

namespace lnb {
// tile table + compact Ritz row list from the extents (shared by lnb_graph_prepare and
// lnb_graph_prepare_sparse); tiles = [4*B + 2] ints: B + 2 table entries followed by 3*B scratch
void launch_tile_assign(cudaStream_t s, const int32_t* gext, int B, int K, int32_t* tiles,
                        int32_t* rowmap, int32_t* nrows) {
  const size_t tbytes = (size_t)6 * (B + 1) * sizeof(int32_t);
  const int tsm = tbytes <= 200 * 1024;
  if (tsm && tbytes > 40 * 1024)
    cudaFuncSetAttribute(tile_assign_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tbytes);
  if (tsm)
    tile_assign_kernel<true><<<1, 1024, tbytes, s>>>(gext, B, K, tiles, tiles + B + 2, rowmap, nrows);
  else
    tile_assign_kernel<false><<<1, 1024, 0, s>>>(gext, B, K, tiles, tiles + B + 2, rowmap, nrows);
}
}  // namespace lnb

extern "C" {

// profiling aid: register (or clear with NULL) a device buffer of SMs x 32 uint64 phase timers
int lnb_debug_set_prof(unsigned long long* buf) {
  cudaError_t e = cudaMemcpyToSymbol(tcg::g_prof, &buf, sizeof(buf));
  if (e != cudaSuccess) { lnb::set_err("debug_set_prof: %s", cudaGetErrorString(e)); return (int)e; }
  lnb::set_prof_buffer(buf);
  return LNB_OK;
}

int lnb_graph_prepare(lnb_stream_t stream, const float* L, const float* Q, int B, int N, int E1,
                      int K, float* ell_val, uint8_t* ell_idx, int32_t* ell_max, int32_t* gext,
                      int32_t* tiles /* [4*B + 2]: B + 2 tile table followed by 3*B scratch */,
                      int32_t* rowmap, int32_t* nrows, int flags) {
  LNB_REQUIRE(L && Q && ell_val && ell_idx && ell_max && gext && tiles, "graph_prepare: null pointer");
  LNB_REQUIRE((rowmap == nullptr) == (nrows == nullptr), "graph_prepare: rowmap and nrows go together");
  LNB_REQUIRE(B >= 0 && N >= 1 && N <= 255 && E1 >= 1 && E1 <= EMAX && K >= 1,
              "graph_prepare: bad dims B=%d N=%d E1=%d K=%d", B, N, E1, K);
  if (B == 0) return LNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t lbytes = (size_t)N * N * E1 * sizeof(float);
  const int stage = lbytes <= 64 * 1024;
  if (stage) {
    if (lbytes > 40 * 1024)
      cudaFuncSetAttribute(graph_prepare_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lbytes);
    graph_prepare_kernel<true><<<B, 256, lbytes, s>>>(L, Q, N, E1, K, ell_val, ell_idx, ell_max, gext,
                                                      flags & 1);
  } else {
    graph_prepare_kernel<false><<<B, 256, 0, s>>>(L, Q, N, E1, K, ell_val, ell_idx, ell_max, gext,
                                                  flags & 1);
  }
  lnb::launch_tile_assign(s, gext, B, K, tiles, rowmap, nrows);
  lnb::count_launch(2);
  return lnb::finish_launch("graph_prepare");
}

static int launch_stack(lnb_stream_t stream, const lnb_spectral_stack& d, const char* who) {
  LNB_REQUIRE((d.X || (d.node_ids && d.emb_table)) && d.Q && d.ell_val && d.ell_idx && d.ell_max &&
                  d.gext && d.tiles && d.W_hi && d.W_lo && (d.out_state || d.score) &&
                  (d.coeff || d.S == 0),
              "%s: null pointer", who);
  LNB_REQUIRE(d.B >= 0 && d.N >= 1 && d.E1 >= 1 && d.K >= 1 && d.S >= 0 && d.H >= 1 &&
                  d.num_layers >= 1 && d.num_layers <= SpectralPolicy::LMAX,
              "%s: bad dims", who);
  LNB_REQUIRE(!d.score || (d.W_out && d.b_out && d.w_att && d.b_att && d.P >= 1 && d.P <= 48),
              "%s: readout needs W_out, b_out, w_att, b_att and 1 <= P <= 48", who);
  int dmax = 0;
  bool ok = d.N <= RMAX && d.K <= KMAX && d.K % 4 == 0 && d.H % 4 == 0 && d.H <= tcg::BN && d.E1 <= EMAX;
  for (int l = 0; l < d.num_layers; ++l) {
    ok = ok && d.Din[l] % 32 == 0 && d.Din[l] >= 32 && (l == 0 || d.Din[l] == d.H);
    dmax = d.Din[l] > dmax ? d.Din[l] : dmax;
  }
  ok = ok && (d.S + d.E1) * dmax <= d.Kw;
  if (!ok) {
    lnb::set_err("%s: unsupported shape N=%d K=%d H=%d E1=%d (needs N<=128, Din%%32==0, inner "
                 "layers Din==H, K%%4==0, K<=%d, H%%4==0, H<=128, E1<=%d)", who, d.N, d.K, d.H, d.E1,
                 KMAX, EMAX);
    return LNB_ERR_UNSUPPORTED;
  }
  if ((d.bias && (reinterpret_cast<uintptr_t>(d.bias) & 15)) || (reinterpret_cast<uintptr_t>(d.Q) & 15) ||
      (d.X && (reinterpret_cast<uintptr_t>(d.X) & 15)) || (d.emb_table && (reinterpret_cast<uintptr_t>(d.emb_table) & 15))) {
    lnb::set_err("%s: X / emb_table / Q / bias must be 16-byte aligned", who);
    return LNB_ERR_ARG;
  }
  if (d.B == 0) return LNB_OK;
  size_t smem = tcg::core_smem(SpectralPolicy::kStagesB) + 1024 + SpectralPolicy::smem_fixed(dmax, d.K, d.H);
  if (smem > 227 * 1024) {
    lnb::set_err("%s: tile state (Din=%d, K=%d, H=%d) needs %zu B of shared memory", who, dmax, d.K,
                 d.H, smem);
    return LNB_ERR_UNSUPPORTED;
  }
  int lb = (int)((227 * 1024 - smem) / SpectralPolicy::ell_line_bytes());
  if (lb > 255) lb = 255;
  smem += (size_t)lb * SpectralPolicy::ell_line_bytes();
  CUtensorMap map_hi, map_lo;
  int rc = tcg::make_weight_map(&map_hi, d.W_hi, d.num_layers * d.H, d.Kw, who);
  if (rc != LNB_OK) return rc;
  rc = tcg::make_weight_map(&map_lo, d.W_lo, d.num_layers * d.H, d.Kw, who);
  if (rc != LNB_OK) return rc;
  auto kern = tcg::tc_gemm_kernel<SpectralPolicy>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SpectralPolicy::Params p{};
  p.X = d.X; p.node_ids = d.node_ids; p.emb = d.emb_table; p.Q = d.Q;
  p.coeff = d.coeff; p.coeff_stride = d.coeff_layer_stride;
  p.ell_val = d.ell_val; p.ell_idx = d.ell_idx; p.ell_max = d.ell_max; p.gext = d.gext; p.tiles = d.tiles;
  p.bias = d.bias; p.out = d.out_state;
  p.W_out = d.W_out; p.b_out = d.b_out; p.w_att = d.w_att; p.b_att = d.b_att; p.mask = d.mask;
  p.score = d.score; p.P = d.P; p.emb_rows = d.emb_rows; p.L = d.num_layers;
  for (int l = 0; l < d.num_layers; ++l) p.Din[l] = d.Din[l];
  p.B = d.B; p.N = d.N; p.E1 = d.E1; p.K = d.K; p.S = d.S; p.H = d.H; p.relu = d.relu;
  p.LB = lb; p.write_pad = d.write_pad; p.dbg = tcg::debug_flags();
  // the tile count lives in device memory (no host sync): one persistent CTA per SM, bounded by
  // the worst case of one graph per tile
  const int grid = d.B < tcg::sm_count() ? d.B : tcg::sm_count();
  kern<<<grid, tcg::THREADS, smem, (cudaStream_t)stream>>>(map_hi, map_lo, p);
  lnb::count_launch();
  return lnb::finish_launch(who);
}

int lnb_spectral_stack_forward(lnb_stream_t stream, const lnb_spectral_stack* desc) {
  LNB_REQUIRE(desc, "spectral_stack_forward: null descriptor");
  return launch_stack(stream, *desc, "spectral_stack_forward");
}

int lnb_spectral_conv_fused(lnb_stream_t stream, const float* X, const float* Q, const float* coeff,
                            const float* ell_val, const uint8_t* ell_idx, const int32_t* ell_max,
                            const int32_t* gext, const int32_t* tiles, const float* W_hi,
                            const float* W_lo, const float* bias, int B, int N, int Din, int E1,
                            int K, int S, int H, int relu, int write_pad, float* out) {
  lnb_spectral_stack d{};
  d.X = X; d.Q = Q; d.coeff = coeff; d.coeff_layer_stride = 0;
  d.ell_val = ell_val; d.ell_idx = ell_idx; d.ell_max = ell_max; d.gext = gext; d.tiles = tiles;
  d.W_hi = W_hi; d.W_lo = W_lo; d.Kw = (S + E1) * Din; d.bias = bias;
  d.Din[0] = Din; d.num_layers = 1; d.out_state = out; d.write_pad = write_pad;
  d.B = B; d.N = N; d.E1 = E1; d.K = K; d.S = S; d.H = H; d.relu = relu;
  return launch_stack(stream, d, "spectral_conv_fused");
}

}  // extern "C"
