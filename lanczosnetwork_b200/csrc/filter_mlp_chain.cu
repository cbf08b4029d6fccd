// Ritz-value filter MLPs of ALL layers in one persistent tcgen05 kernel
// (reference: model/lanczos_net.py:47-58 the per-layer Sequential, :109-113 its application to
// the B*K rows of Ritz-value powers).
//
// An item is (128-row tile, layer): four Linear stages S -> Hd -> Hd -> Hd -> S.
//   * The S-wide first stage (S <= 8 inputs) is evaluated by the CUDA cores in plain fp32 straight
//     into the A operand of the second stage -- a tensor-core step for K = 8 would cost a full
//     accumulator hand-over for 2 % of the flops.  (S > 8 falls back to an MMA stage.)
//   * The other stages are accumulator lifetimes ("steps") of 3xTF32 MMAs with the A operand in
//     tensor memory.  Activations NEVER leave tensor memory / registers: the epilogue of stage s
//     reads the accumulator (tcgen05.ld), applies bias + ReLU, splits into tf32 hi / lo and writes
//     the result straight into the A ring (tcgen05.st) as the k-blocks of stage s+1 -- output
//     chunk cc of stage s IS k-block cc of stage s+1 for the same thread (row <-> TMEM lane).
//
// Every CTA walks a contiguous range of items (layer-major order) two at a time: A and B are
// consecutive row tiles of the SAME layer with the steps interleaved A1 B1 A2 B2 A3 B3.  While
// the CUDA cores turn A.s into the operand of A.(s+1) the tensor core runs B.s, so the MMA pipe
// only idles while an accumulator is drained to registers; and every W tile is fetched from L2
// once per pair (A.s and B.s multiply by the same weights) -- at full MMA rate a single item
// would ask L2 for more than its ~42 B/clk/SM share.  The last k-block of a stage can only be
// written once the other tile's step has released its ring slot, i.e. when that step's
// accumulator is ready: its owner drains that accumulator FIRST (so the next step starts on the
// three k-blocks already in the ring) and writes the parked k-block afterwards.
//
// Warps: 0-11 workers (group g = warp / 4 owns chunk / k-block cc when cc % 3 == g; lane quarter
// = warp % 4), 12 TMA (W tiles, 6-slot shared-memory ring: the 4 k-blocks of a stage stay until B
// has used them while the next stage prefetches), 13 MMA issue.
// TMEM: [0,128) D_main, [128,256) D_corr, [256,512) A ring of 4 slots x (32 hi + 32 lo).
#include "tc_gemm.cuh"

namespace {

namespace chain {

constexpr int NSLOT = 4;                 // A ring slots in tensor memory
constexpr int NSTB = 6;                  // W ring slots in shared memory (one stage = 4 + prefetch)
constexpr int NGRP = 3;
constexpr int WORKER_WARPS = 4 * NGRP;
constexpr int WORKER_THREADS = 32 * WORKER_WARPS;
constexpr int TMA_WARP = WORKER_WARPS, MMA_WARP = WORKER_WARPS + 1;
constexpr int THREADS = (WORKER_WARPS + 2) * 32;
constexpr int COL_MAIN = 0, COL_CORR = 128, COL_A = 256;
constexpr int S0MAX = 8;                 // first-stage widths the CUDA cores handle
#ifndef LNB_CHAIN_MMA0
#define LNB_CHAIN_MMA0 0
#endif
constexpr bool FORCE_MMA0 = LNB_CHAIN_MMA0 != 0;
constexpr int OFF_BARS = NSTB * tcg::STAGE_B_BYTES;
constexpr int OFF_PARK = OFF_BARS + 256;                     // [32][128] fp32: group 0's parked chunk
constexpr int OFF_W1 = OFF_PARK + 32 * 128 * 4;              // [128][S0MAX] first-stage weights + [128] bias
constexpr int SMEM_BYTES = OFF_W1 + (128 * S0MAX + 128) * 4;

struct Params {
  const float* table;     // [Rall, S]  powers of the Ritz values
  const int32_t* rowmap;  // [Rall]     compact list of rows to evaluate (nullptr: all rows)
  const int32_t* nrows;   // [1]        number of valid entries in rowmap (nullptr: Rall)
  const float* W_hi;      // [L * (3*Hd + S), Hd] split weights (the first stage reads hi + lo)
  const float* W_lo;
  const float* bias_all;  // [L * (3*Hd + S)]
  float* coeff;           // [L, Rall, S]
  int Rall, L, S, Hd;
  unsigned long long* prof;   // profiling aid: MMA-warp wait cycles per CTA (16 slots), or nullptr
};

// The step sequence every role of the CTA walks in the same order.
struct Walk {
  int rows, ntile, i0, i1, n;            // valid rows, row tiles, item range [i0, i1), k-blocks per hidden stage
  bool mma0;                             // first stage on the tensor core (S > S0MAX)
  __device__ Walk(const Params& p, int cta, int ncta) {
    rows = p.nrows ? __ldg(p.nrows) : p.Rall;
    ntile = (rows + tcg::BM - 1) / tcg::BM;
    const long long items = (long long)ntile * p.L;
    i0 = (int)(items * cta / ncta);
    i1 = (int)(items * (cta + 1) / ncta);
    n = p.Hd / tcg::BK;
    mma0 = p.S > S0MAX || FORCE_MMA0;
  }
  // item i = (layer i / ntile, tile i % ntile); items i, i + 1 pair up when they share the layer
  __device__ __forceinline__ bool paired(int i) const { return i + 1 < i1 && (i + 1) / ntile == i / ntile; }
  __device__ __forceinline__ int s_first() const { return mma0 ? 0 : 1; }
  __device__ __forceinline__ int nkb(int s) const { return s == 0 ? 1 : n; }
  // k-blocks before stage s within one item
  __device__ __forceinline__ int kb_before(int s) const { return mma0 ? (s == 0 ? 0 : 1 + n * (s - 1)) : n * (s - 1); }
  __device__ __forceinline__ int kb_item() const { return mma0 ? 1 + 3 * n : 3 * n; }
  // first k-block (relative to the pair start) of step (X, s) in the order A.s B.s A.(s+1) ...
  __device__ __forceinline__ int base(bool vb, int X, int s) const {
    return vb ? 2 * kb_before(s) + X * nkb(s) : kb_before(s);
  }
};

__device__ __forceinline__ int w_row0(const Params& p, int layer, int stage) {
  return layer * (3 * p.Hd + p.S) + stage * p.Hd;
}
__device__ __forceinline__ void workers_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(WORKER_THREADS) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
mlp_chain_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                 const __grid_constant__ CUtensorMap map_hi32, const __grid_constant__ CUtensorMap map_lo32,
                 const Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* Bst = smem_raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bst + OFF_BARS);
  uint64_t* b_full = bars;                 // [NSTB]
  uint64_t* b_empty = b_full + NSTB;       // [NSTB]
  uint64_t* a_full = b_empty + NSTB;       // [NSLOT]  4 warp arrivals
  uint64_t* a_empty = a_full + NSLOT;      // [NSLOT]  tcgen05.commit
  uint64_t* acc_full = a_empty + NSLOT;    // tcgen05.commit
  uint64_t* acc_empty = acc_full + 1;      // 12 warp arrivals
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 1);
  float* park = reinterpret_cast<float*>(Bst + OFF_PARK);
  float* W1s = reinterpret_cast<float*>(Bst + OFF_W1);       // [Hd][S0MAX]
  float* b1s = W1s + 128 * S0MAX;                            // [Hd]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Walk w(p, blockIdx.x, gridDim.x);
  const int n = w.n;
  const int ksteps0 = (p.S + 7) >> 3;      // tf32 k-steps of an S-wide MMA first stage

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
    // group 0's last chunk of a paired step waits in shared memory until the next drain is done
    int parked_G = -1;
    auto park_put = [&](int G, const float (&y)[32]) {
#pragma unroll
      for (int j = 0; j < 32; ++j) park[j * 128 + r] = y[j];
      parked_G = G;
    };
    auto park_flush = [&]() {
      if (parked_G < 0) return;
      float y[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = park[j * 128 + r];
      emit(parked_G, y);
      parked_G = -1;
    };

    // profiling aid (warp 0): [16] wait acc_full, [17] drain, [18] parked flush, [19] first emit,
    // [20] second chunk, [21] first-stage compute, [22] output stage
    long long wt[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool wprof = p.prof != nullptr && warp == 0;
    long long wt0 = wprof ? clock64() : 0;
    auto wlap = [&](int i) { if (wprof) { const long long t = clock64(); wt[i] += t - wt0; wt0 = t; } };
    uint32_t nstep = 0;
    int Gp = 0;                                               // k-block count at the pair start
    int cur_layer = -1;
    for (int it = w.i0; it < w.i1;) {
      const int item[2] = {it, it + 1};
      const bool vb = w.paired(it);
      it += vb ? 2 : 1;
      int src[2] = {-1, -1};
      const int layer = item[0] / w.ntile;
#pragma unroll
      for (int X = 0; X < 2; ++X) {
        if (X == 1 && !vb) break;
        const int i = (item[X] % w.ntile) * tcg::BM + r;
        if (i < w.rows) src[X] = p.rowmap ? __ldg(p.rowmap + i) : i;
      }
      if (!w.mma0) {
        // ---- first stage on the CUDA cores, written as the k-blocks of stage 1 -------------
        if (layer != cur_layer) {                             // stage this layer's W1 (hi + lo), b1
          workers_sync();
          const int row0 = w_row0(p, layer, 0);
          for (int e = tid; e < p.Hd * S0MAX; e += WORKER_THREADS) {
            const int c = e / S0MAX, i = e - c * S0MAX;
            const int64_t o = (int64_t)(row0 + c) * p.Hd + i;
            W1s[e] = (i < p.S) ? __ldg(p.W_hi + o) + __ldg(p.W_lo + o) : 0.f;
          }
          for (int c = tid; c < p.Hd; c += WORKER_THREADS) b1s[c] = __ldg(p.bias_all + row0 + c);
          workers_sync();
          cur_layer = layer;
        }
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          if (X == 1 && !vb) break;
          float tv[S0MAX];
#pragma unroll
          for (int i = 0; i < S0MAX; ++i)
            tv[i] = (src[X] >= 0 && i < p.S) ? __ldg(p.table + (int64_t)src[X] * p.S + i) : 0.f;
          const int nb = Gp + w.base(vb, X, 1);
          for (int cc = grp; cc < n; cc += NGRP) {
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float4* wr = reinterpret_cast<const float4*>(W1s + (cc * 32 + j) * S0MAX);
              const float4 wa = wr[0], wb = wr[1];
              float a = b1s[cc * 32 + j];
              a = fmaf(tv[0], wa.x, a); a = fmaf(tv[1], wa.y, a); a = fmaf(tv[2], wa.z, a); a = fmaf(tv[3], wa.w, a);
              a = fmaf(tv[4], wb.x, a); a = fmaf(tv[5], wb.y, a); a = fmaf(tv[6], wb.z, a); a = fmaf(tv[7], wb.w, a);
              y[j] = fmaxf(a, 0.f);
            }
            // the last chunk of B's stage needs A's step to make room first: park it
            if (vb && X == 1 && cc == n - 1 && cc >= NGRP) park_put(nb + cc, y); else emit(nb + cc, y);
          }
        }
      } else {
        // ---- S > 8: the first stage is an MMA step; its operand = the S powers ---------------
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          if (X == 1 && !vb) break;
          const int G = Gp + w.base(vb, X, 0);
          if (grp == 0) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = (src[X] >= 0 && j < p.S) ? __ldg(p.table + (int64_t)src[X] * p.S + j) : 0.f;
            emit(G, v);
          }
        }
      }
      wlap(5);
      const int nst = (vb ? 2 : 1) * (4 - w.s_first());
      for (int j = 0; j < nst; ++j, ++nstep) {
        const int X = vb ? (j & 1) : 0, s = w.s_first() + (vb ? (j >> 1) : j);
        const float* bias = p.bias_all + w_row0(p, layer, s);
        tc05::mbar_wait(acc_full, nstep & 1u);
        tc05::fence_after_thread_sync();
        wlap(0);
        if (s < 3) {
          // chunk cc of this stage becomes k-block nb + cc of the next one; group cc % 3 drains it
          const int nb = Gp + w.base(vb, X, s + 1);
          const int c0 = grp, c1 = grp + NGRP;
          if (c1 < n) {                                       // two chunks
            uint32_t ua[32], ub[32];
            float xa[32], xb[32];
            ld_main(c0, ua);
            ld_main(c1, ub);
            add_corr(c0, ua, xa);                             // its wait::ld covers the loads above
            add_corr(c1, ub, xb);
            release_acc();                                    // the next step may overwrite D now
            wlap(1);
            park_flush();                                     // k-block parked by the previous step
            wlap(2);
            bias_relu(bias + c0 * 32, xa);
            emit(nb + c0, xa);
            wlap(3);
            bias_relu(bias + c1 * 32, xb);
            if (vb && c1 == n - 1) park_put(nb + c1, xb); else emit(nb + c1, xb);
            wlap(4);
          } else if (c0 < n) {
            uint32_t ua[32];
            float xa[32];
            ld_main(c0, ua);
            add_corr(c0, ua, xa);
            release_acc();
            park_flush();
            bias_relu(bias + c0 * 32, xa);
            emit(nb + c0, xa);
          } else {
            release_acc();
            park_flush();
          }
        } else {
          const bool mine = (int)(nstep % NGRP) == grp;      // rotate the output work over groups
          float x[32];
          if (mine) { uint32_t u[32]; ld_main(0, u); add_corr(0, u, x); }
          release_acc();
          park_flush();
          if (mine && src[X] >= 0) {
            float* dst = p.coeff + ((int64_t)layer * p.Rall + src[X]) * p.S;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < p.S) dst[c] = x[c] + __ldg(bias + c);
          }
          wlap(6);
        }
      }
      Gp += (vb ? 2 : 1) * w.kb_item();
    }
    if (wprof && lane == 0)
      for (int i = 0; i < 7; ++i) p.prof[blockIdx.x * 32 + 16 + i] = (unsigned long long)wt[i];
  } else if (warp == TMA_WARP) {
    // ================================ TMA producer (W tiles) ==============================
    // one load per (stage, k-block) of an item or same-layer pair
    uint32_t Wg = 0;
    for (int it = w.i0; it < w.i1;) {
      const int layer = it / w.ntile;
      it += w.paired(it) ? 2 : 1;
      for (int st_ = w.s_first(); st_ < 4; ++st_) {
        const int row0 = w_row0(p, layer, st_);
        const int nkb = w.nkb(st_);
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
    // profiling aid: cycles the issuing warp waits for [0] a drained accumulator, [1] W tiles,
    // [2] A k-blocks, [3] issues; [4] whole loop
    long long pt[5] = {0, 0, 0, 0, 0};
    long long pa[4] = {0, 0, 0, 0}, pw[4] = {0, 0, 0, 0};   // hidden stages: A / W wait by k-block index
    const bool prof = p.prof != nullptr;
    const long long pstart = prof ? clock64() : 0;
    for (int it = w.i0; it < w.i1;) {
      const bool vb = w.paired(it);
      it += vb ? 2 : 1;
      const int nst = (vb ? 2 : 1) * (4 - w.s_first());
      for (int j = 0; j < nst; ++j, ++nstep) {
        const int X = vb ? (j & 1) : 0, s = w.s_first() + (vb ? (j >> 1) : j);
        const int nkb = w.nkb(s);
        const int ksteps = s == 0 ? ksteps0 : tcg::BK / 8;
        const bool last_use = !vb || X == 1;                  // B (or a single item) frees the W slot
        long long t0 = prof ? clock64() : 0, t1;
        tc05::mbar_wait(acc_empty, (nstep & 1u) ^ 1u);       // previous accumulator drained
        tc05::fence_after_thread_sync();
        if (prof) { t1 = clock64(); pt[0] += t1 - t0; t0 = t1; }
        for (int kb = 0; kb < nkb; ++kb, ++G) {
          const uint32_t Wg = Wp + w.kb_before(s) + kb;
          const uint32_t st = Wg % NSTB, slot = G % NSLOT;
          tc05::mbar_wait2(&b_full[st], (Wg / NSTB) & 1u, &a_full[slot], (G / NSLOT) & 1u);
          tc05::fence_after_thread_sync();
          if (prof) { t1 = clock64(); pt[2] += t1 - t0; if (s < 3) pa[kb & 3] += t1 - t0; t0 = t1; }
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
          if (prof) { t1 = clock64(); pt[3] += t1 - t0; t0 = t1; }
        }
      }
      Wp += w.kb_item();
    }
    if (prof && lane == 0) {
      pt[4] = clock64() - pstart;
      for (int i = 0; i < 5; ++i) p.prof[blockIdx.x * 32 + i] = (unsigned long long)pt[i];
      for (int i = 0; i < 4; ++i) {
        p.prof[blockIdx.x * 32 + 5 + i] = (unsigned long long)pa[i];
        p.prof[blockIdx.x * 32 + 9 + i] = (unsigned long long)pw[i];
      }
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
  chain::Params p{table, rowmap, nrows, W_hi, W_lo, bias_all, coeff, Rall, L, S, Hd, lnb::prof_buffer()};
  const int items = lnb::ceil_div(Rall, tcg::BM) * L;
  const int grid = items < tcg::sm_count() ? items : tcg::sm_count();
  chain::mlp_chain_kernel<<<grid, chain::THREADS, smem, (cudaStream_t)stream>>>(map_hi, map_lo, map_hi32, map_lo32, p);
  lnb::count_launch();
  return lnb::finish_launch("ritz_filter_mlp");
}

}  // extern "C"
