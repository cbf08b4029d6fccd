// Ritz-value filter MLPs of ALL layers in one persistent tcgen05 kernel
// (reference: model/lanczos_net.py:47-58 the per-layer Sequential, :109-113 its application to
// the B*K rows of Ritz-value powers).
//
// An item is (128-row tile, layer): four Linear stages S -> Hd -> Hd -> Hd -> S, each one
// accumulator lifetime ("step") of 3xTF32 MMAs with the A operand in tensor memory.  The
// activations NEVER leave tensor memory / registers: the epilogue of stage s reads the accumulator
// (tcgen05.ld), applies bias + ReLU, splits into tf32 hi / lo and writes the result straight
// into the A ring (tcgen05.st) as the k-blocks of stage s+1 -- output chunk cc of stage s IS
// k-block cc of stage s+1 for the same thread (row <-> TMEM lane in both).
//
// Every CTA walks a contiguous range of items (layer-major order) two at a time: A and B are
// consecutive row tiles of the SAME layer with the steps interleaved A0 B0 A1 B1 A2 B2 A3 B3.
// While the CUDA cores turn A.s into the operand of A.(s+1) the tensor core runs B.s, so the MMA
// pipe only idles while an accumulator is drained to registers; and every W tile is fetched from
// L2 once per pair (A.s and B.s multiply by the same weights) -- at full MMA rate a single item
// would ask L2 for more than its ~42 B/clk/SM share.
// Only rows (graph, k) with k < k_eff(graph) are evaluated (rowmap): zero-padded Ritz pairs
// multiply zero Ritz vectors downstream (exact).
//
// Warps: 0-11 workers (group g = warp / 4 owns k-block G when G % 3 == g, G = CTA-global k-block
// count; lane quarter = warp % 4), 12 TMA (W tiles, 6-slot shared-memory ring: the 4 k-blocks of
// a stage stay until B has used them while the next stage prefetches), 13 MMA issue.
// TMEM: [0,128) D_main, [128,256) D_corr, [256,512) A ring of 4 slots x (32 hi + 32 lo).
#include "tc_gemm.cuh"

namespace {

namespace chain {

constexpr int NSLOT = 4;                 // A ring slots in tensor memory
constexpr int NSTB = 6;                  // W ring slots in shared memory (one stage = 4 + prefetch)
constexpr int NGRP = 3;
constexpr int WORKER_WARPS = 4 * NGRP;
constexpr int TMA_WARP = WORKER_WARPS, MMA_WARP = WORKER_WARPS + 1;
constexpr int THREADS = (WORKER_WARPS + 2) * 32;
constexpr int COL_MAIN = 0, COL_CORR = 128, COL_A = 256;
constexpr int SMEM_BYTES = NSTB * tcg::STAGE_B_BYTES + 256;

struct Params {
  const float* table;     // [Rall, S]  powers of the Ritz values
  const int32_t* rowmap;  // [Rall]     compact list of rows to evaluate (nullptr: all rows)
  const int32_t* nrows;   // [1]        number of valid entries in rowmap (nullptr: Rall)
  const float* bias_all;  // [L * (3*Hd + S)]
  float* coeff;           // [L, Rall, S]
  int Rall, L, S, Hd;
};

// The step sequence every role of the CTA walks in the same order.
struct Walk {
  int rows, ntile, i0, i1, n;            // valid rows, row tiles, item range [i0, i1), k-blocks per hidden stage
  __device__ Walk(const Params& p, int cta, int ncta) {
    rows = p.nrows ? __ldg(p.nrows) : p.Rall;
    ntile = (rows + tcg::BM - 1) / tcg::BM;
    const long long items = (long long)ntile * p.L;
    i0 = (int)(items * cta / ncta);
    i1 = (int)(items * (cta + 1) / ncta);
    n = p.Hd / tcg::BK;
  }
  // item i = (layer i / ntile, tile i % ntile); items i, i + 1 pair up when they share the layer
  __device__ __forceinline__ bool paired(int i) const { return i + 1 < i1 && (i + 1) / ntile == i / ntile; }
};

__device__ __forceinline__ int w_row0(const Params& p, int layer, int stage) {
  return layer * (3 * p.Hd + p.S) + stage * p.Hd;
}

__global__ void __launch_bounds__(THREADS, 1)
mlp_chain_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                 const __grid_constant__ CUtensorMap map_hi32, const __grid_constant__ CUtensorMap map_lo32,
                 const Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* Bst = smem_raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bst + NSTB * tcg::STAGE_B_BYTES);
  uint64_t* b_full = bars;                 // [NSTB]
  uint64_t* b_empty = b_full + NSTB;       // [NSTB]
  uint64_t* a_full = b_empty + NSTB;       // [NSLOT]  4 warp arrivals
  uint64_t* a_empty = a_full + NSLOT;      // [NSLOT]  tcgen05.commit
  uint64_t* acc_full = a_empty + NSLOT;    // tcgen05.commit
  uint64_t* acc_empty = acc_full + 1;      // 12 warp arrivals
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Walk w(p, blockIdx.x, gridDim.x);
  const int n = w.n;
  const int ksteps0 = (p.S + 7) >> 3;      // tf32 k-steps of the S-wide first stage

  if (warp == TMA_WARP && lane == 0) {
    tc05::tma_prefetch_desc(&map_hi);
    tc05::tma_prefetch_desc(&map_lo);
    tc05::tma_prefetch_desc(&map_hi32);
    tc05::tma_prefetch_desc(&map_lo32);
  }
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < NSTB; ++i) { tc05::mbar_init(&b_full[i], 1); tc05::mbar_init(&b_empty[i], 1); }
      for (int i = 0; i < NSLOT; ++i) { tc05::mbar_init(&a_full[i], 4); tc05::mbar_init(&a_empty[i], 1); }
      tc05::mbar_init(acc_full, 1);
      tc05::mbar_init(acc_empty, WORKER_WARPS);
      tc05::fence_barrier_init();
    }
    __syncwarp();
    tc05::tmem_alloc(tmem_holder, tcg::TMEM_COLS);
  }
  tc05::fence_before_thread_sync();
  __syncthreads();
  tc05::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < WORKER_WARPS) {
    // ================================ workers ============================================
    const int grp = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;                              // row of the tile <-> TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);

    // write 32 fp32 values of this row as k-block G of the A ring (tf32 hi | lo)
    auto emit = [&](int G, float (&y)[32]) {
      const int slot = G % NSLOT;
      const uint32_t a_addr = lane_addr + COL_A + slot * 64;
      tc05::mbar_wait(&a_empty[slot], (((uint32_t)G / NSLOT) & 1u) ^ 1u);
      tc05::fence_after_thread_sync();
      uint32_t part[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) part[j] = tc05::tf32_rna_bits(y[j]);
      tc05::tmem_st_32x32(a_addr, part);
#pragma unroll
      for (int j = 0; j < 32; ++j) part[j] = tc05::tf32_rna_bits(y[j] - __uint_as_float(part[j]));
      tc05::tmem_st_32x32(a_addr + 32, part);
      tc05::tmem_wait_st();
      tc05::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) tc05::mbar_arrive(&a_full[slot]);
    };
    // accumulator columns [32 cc, 32 cc + 32) of this row: main part, then + correction
    auto ld_main = [&](int cc, uint32_t (&u)[32]) {
      tc05::tmem_ld_32x32(lane_addr + COL_MAIN + cc * 32, u);
    };
    auto add_corr = [&](int cc, const uint32_t (&u)[32], float (&x)[32]) {
      uint32_t vc[32];
      tc05::tmem_ld_32x32(lane_addr + COL_CORR + cc * 32, vc);
      tc05::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(u[j]) + __uint_as_float(vc[j]);
    };
    auto bias_relu = [&](const float* bias, float (&x)[32]) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j] + __ldg(bias + j), 0.f);
    };
    auto release_acc = [&]() {
      tc05::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) tc05::mbar_arrive(acc_empty);
    };

    uint32_t nstep = 0;
    int Gp = 0;                                               // k-block count at the pair start
    for (int it = w.i0; it < w.i1;) {
      const int item[2] = {it, it + 1};
      const bool vb = w.paired(it);
      it += vb ? 2 : 1;
      // first k-block of step (X, s) in the interleaved order A0 B0 A1 B1 ... (or A0 A1 A2 A3)
      auto base = [&](int X, int s) {
        return vb ? Gp + (s == 0 ? X : 2 + 2 * n * (s - 1) + X * n) : Gp + (s == 0 ? 0 : 1 + n * (s - 1));
      };
      int src[2] = {-1, -1};
      const int layer = item[0] / w.ntile;
#pragma unroll
      for (int X = 0; X < 2; ++X) {
        if (X == 1 && !vb) break;
        const int i = (item[X] % w.ntile) * tcg::BM + r;
        if (i < w.rows) src[X] = p.rowmap ? __ldg(p.rowmap + i) : i;
      }
      // stage-0 operands: the S powers of this row's Ritz value
#pragma unroll
      for (int X = 0; X < 2; ++X) {
        if (X == 1 && !vb) break;
        const int G = base(X, 0);
        if (G % NGRP == grp) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = (src[X] >= 0 && j < p.S) ? __ldg(p.table + (int64_t)src[X] * p.S + j) : 0.f;
          emit(G, v);
        }
      }
      const int nst = vb ? 8 : 4;
      for (int j = 0; j < nst; ++j, ++nstep) {
        const int X = vb ? (j & 1) : 0, s = vb ? (j >> 1) : j;
        const float* bias = p.bias_all + w_row0(p, layer, s);
        tc05::mbar_wait(acc_full, nstep & 1u);
        tc05::fence_after_thread_sync();
        if (s < 3) {
          // chunk cc of this stage becomes k-block nb + cc of the next one; its owner drains it
          const int nb = base(X, s + 1);
          const int c0 = (grp + NGRP - nb % NGRP) % NGRP, c1 = c0 + NGRP;
          if (c1 < n) {                                       // two chunks: c0 and c0 + 3
            uint32_t ua[32], ub[32];
            float xa[32], xb[32];
            ld_main(c0, ua);
            ld_main(c1, ub);
            add_corr(c0, ua, xa);                             // its wait::ld covers the loads above
            add_corr(c1, ub, xb);
            release_acc();                                    // the next step may overwrite D now
            bias_relu(bias + c0 * 32, xa);
            emit(nb + c0, xa);
            bias_relu(bias + c1 * 32, xb);
            emit(nb + c1, xb);
          } else if (c0 < n) {
            uint32_t ua[32];
            float xa[32];
            ld_main(c0, ua);
            add_corr(c0, ua, xa);
            release_acc();
            bias_relu(bias + c0 * 32, xa);
            emit(nb + c0, xa);
          } else {
            release_acc();
          }
        } else {
          const bool mine = (int)(nstep % NGRP) == grp;      // rotate the output work over groups
          float x[32];
          if (mine) { uint32_t u[32]; ld_main(0, u); add_corr(0, u, x); }
          release_acc();
          if (mine && src[X] >= 0) {
            float* dst = p.coeff + ((int64_t)layer * p.Rall + src[X]) * p.S;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < p.S) dst[c] = x[c] + __ldg(bias + c);
          }
        }
      }
      Gp += vb ? 2 + 6 * n : 1 + 3 * n;
    }
  } else if (warp == TMA_WARP) {
    // ================================ TMA producer (W tiles) ==============================
    // one load per (stage, k-block) of an item or same-layer pair
    uint32_t Wg = 0;
    for (int it = w.i0; it < w.i1;) {
      const int layer = it / w.ntile;
      it += w.paired(it) ? 2 : 1;
      for (int st_ = 0; st_ < 4; ++st_) {
        const int row0 = w_row0(p, layer, st_);
        const int nkb = st_ == 0 ? 1 : n;
        const bool small = st_ == 3;                          // S <= 32 output rows: 32-row boxes
        for (int kb = 0; kb < nkb; ++kb, ++Wg) {
          const uint32_t st = Wg % NSTB;
          tc05::mbar_wait(&b_empty[st], ((Wg / NSTB) & 1u) ^ 1u);
          if (tc05::elect_one()) {
            uint8_t* dst = Bst + st * tcg::STAGE_B_BYTES;
            tc05::mbar_arrive_expect_tx(&b_full[st], small ? 2 * 32 * 128 : tcg::STAGE_B_BYTES);
            tc05::tma_load_2d(dst, small ? &map_hi32 : &map_hi, &b_full[st], kb * tcg::BK, row0);
            tc05::tma_load_2d(dst + tcg::TILE_B_BYTES, small ? &map_lo32 : &map_lo, &b_full[st], kb * tcg::BK, row0);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ MMA issuer ==========================================
    constexpr uint32_t idesc256 = tc05::umma_idesc_tf32(tcg::BM, 2 * tcg::BN);
    constexpr uint32_t idesc128 = tc05::umma_idesc_tf32(tcg::BM, tcg::BN);
    constexpr uint32_t idesc32 = tc05::umma_idesc_tf32(tcg::BM, 32);
    const uint32_t d_main = tmem_base + COL_MAIN, d_corr = tmem_base + COL_CORR;
    uint32_t G = 0, Wp = 0, nstep = 0;
    for (int it = w.i0; it < w.i1;) {
      const bool vb = w.paired(it);
      it += vb ? 2 : 1;
      const int nst = vb ? 8 : 4;
      for (int j = 0; j < nst; ++j, ++nstep) {
        const int X = vb ? (j & 1) : 0, s = vb ? (j >> 1) : j;
        const int nkb = s == 0 ? 1 : n;
        const int ksteps = s == 0 ? ksteps0 : tcg::BK / 8;
        const bool last_use = !vb || X == 1;                  // B (or a single item) frees the W slot
        tc05::mbar_wait(acc_empty, (nstep & 1u) ^ 1u);       // previous accumulator drained
        tc05::fence_after_thread_sync();
        for (int kb = 0; kb < nkb; ++kb, ++G) {
          const uint32_t Wg = Wp + (s == 0 ? 0 : 1 + n * (s - 1)) + kb;
          const uint32_t st = Wg % NSTB, slot = G % NSLOT;
          tc05::mbar_wait(&b_full[st], (Wg / NSTB) & 1u);
          tc05::mbar_wait(&a_full[slot], (G / NSLOT) & 1u);
          tc05::fence_after_thread_sync();
          if (tc05::elect_one()) {
            const uint32_t a_hi = tmem_base + COL_A + slot * 64, a_lo = a_hi + 32;
            const uint32_t b_addr = tc05::smem_u32(Bst + st * tcg::STAGE_B_BYTES);
            const uint64_t dB = tc05::umma_desc_kmajor_sw128(b_addr);
            if (s < 3) {
              // [D_main | D_corr] += A_hi [W_hi ; W_lo]^T (N = 256),  D_corr += A_lo W_hi^T
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                tc05::umma_tf32_ts(d_main, a_hi + 8 * k, dB + 2 * k, idesc256, acc);
                tc05::umma_tf32_ts(d_corr, a_lo + 8 * k, dB + 2 * k, idesc128, 1u);
              }
            } else {
              // S <= 32 outputs: three N = 32 products
              const uint64_t dBl = tc05::umma_desc_kmajor_sw128(b_addr + tcg::TILE_B_BYTES);
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                tc05::umma_tf32_ts(d_main, a_hi + 8 * k, dB + 2 * k, idesc32, acc);
                tc05::umma_tf32_ts(d_corr, a_lo + 8 * k, dB + 2 * k, idesc32, acc);
                tc05::umma_tf32_ts(d_corr, a_hi + 8 * k, dBl + 2 * k, idesc32, 1u);
              }
            }
            tc05::umma_commit(&a_empty[slot]);
            if (last_use) tc05::umma_commit(&b_empty[st]);
            if (kb == nkb - 1) tc05::umma_commit(acc_full);
          }
          __syncwarp();
        }
      }
      Wp += 1 + 3 * n;
    }
  }

  tc05::fence_before_thread_sync();
  __syncthreads();
  if (warp == MMA_WARP) {
    __syncwarp();
    tc05::tmem_dealloc(tmem_base, tcg::TMEM_COLS);
  }
}

}  // namespace chain

// Compact list of the (graph, k) rows whose Ritz vector is not identically zero.
__global__ void __launch_bounds__(1024)
ritz_rowmap_kernel(const int32_t* __restrict__ gext, int B, int K, int32_t* __restrict__ rowmap,
                   int32_t* __restrict__ nrows) {
  __shared__ int warp_sums[32];
  __shared__ int running;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) running = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    int k = (b < B) ? min(gext[b * 2 + 1], K) : 0;
    int incl = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sums[lane] = wi - w;          // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int base = running + warp_sums[warp] + incl - k;
    for (int i = 0; i < k; ++i) rowmap[base + i] = b * K + i;
    __syncthreads();
    if (tid == 1023) running = base + k;
    __syncthreads();
  }
  if (tid == 0) nrows[0] = running;
}

}  // namespace

extern "C" {

int lnb_ritz_rowmap(lnb_stream_t stream, const int32_t* gext, int B, int K, int32_t* rowmap,
                    int32_t* nrows) {
  LNB_REQUIRE(gext && rowmap && nrows, "ritz_rowmap: null pointer");
  LNB_REQUIRE(B >= 0 && K >= 1, "ritz_rowmap: bad dims");
  ritz_rowmap_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(gext, B, K, rowmap, nrows);
  lnb::count_launch();
  return lnb::finish_launch("ritz_rowmap");
}

int lnb_ritz_filter_mlp(lnb_stream_t stream, const float* table, const int32_t* rowmap,
                        const int32_t* nrows, const float* W_hi, const float* W_lo,
                        const float* bias_all, int Rall, int L, int S, int Hd, float* coeff) {
  LNB_REQUIRE(table && W_hi && W_lo && bias_all && coeff, "ritz_filter_mlp: null pointer");
  LNB_REQUIRE((rowmap == nullptr) == (nrows == nullptr), "ritz_filter_mlp: rowmap and nrows go together");
  LNB_REQUIRE(Rall >= 0 && L >= 1 && S >= 1 && Hd >= 1, "ritz_filter_mlp: bad dims");
  if (S > 32 || Hd % 32 != 0 || Hd > tcg::BN) {
    lnb::set_err("ritz_filter_mlp: unsupported shape S=%d hidden=%d (needs S<=32, hidden%%32==0, hidden<=128)", S, Hd);
    return LNB_ERR_UNSUPPORTED;
  }
  if (Rall == 0) return LNB_OK;
  const int wrows = L * (3 * Hd + S);
  CUtensorMap map_hi, map_lo;
  int rc = tcg::make_weight_map(&map_hi, W_hi, wrows, Hd, "ritz_filter_mlp");
  if (rc != LNB_OK) return rc;
  rc = tcg::make_weight_map(&map_lo, W_lo, wrows, Hd, "ritz_filter_mlp");
  if (rc != LNB_OK) return rc;
  CUtensorMap map_hi32, map_lo32;                       // 32-row boxes for the S-row last stage
  rc = tcg::make_weight_map(&map_hi32, W_hi, wrows, Hd, "ritz_filter_mlp", 32);
  if (rc != LNB_OK) return rc;
  rc = tcg::make_weight_map(&map_lo32, W_lo, wrows, Hd, "ritz_filter_mlp", 32);
  if (rc != LNB_OK) return rc;
  const size_t smem = chain::SMEM_BYTES + 1024;
  cudaFuncSetAttribute(chain::mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  chain::Params p{table, rowmap, nrows, bias_all, coeff, Rall, L, S, Hd};
  const int items = lnb::ceil_div(Rall, tcg::BM) * L;
  const int grid = items < tcg::sm_count() ? items : tcg::sm_count();
  chain::mlp_chain_kernel<<<grid, chain::THREADS, smem, (cudaStream_t)stream>>>(map_hi, map_lo, map_hi32, map_lo32, p);
  lnb::count_launch();
  return lnb::finish_launch("ritz_filter_mlp");
}

}  // extern "C"
