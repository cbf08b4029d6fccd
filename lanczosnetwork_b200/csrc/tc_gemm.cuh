// Persistent tcgen05 3xTF32 GEMM skeleton (sm_100a):
//     D[128 x 128 tile] = sum_k A[rows, k] * W[n, k]        fp32-grade accuracy from TF32 MMAs
// The A operand of every 128x32 k-block is PRODUCED BY CUDA-CORE WARPS straight into tensor
// memory (tcgen05.st, row r <-> TMEM lane r, k <-> column), so a policy can either load rows
// from HBM (dense layer) or compute them on the fly (fused graph messages) -- the tensor-core
// side is identical.  W_hi / W_lo tiles arrive by TMA (SWIZZLE_128B) through an mbarrier ring.
//
// Accuracy: tensor-core accumulation truncates, so the 3 split products are kept in two
// accumulators -- D_main += A_hi*W_hi and D_corr += A_lo*W_hi + A_hi*W_lo -- and summed in fp32
// by the epilogue (the small terms no longer add truncation steps to the large accumulator).
//
// CTA = 14 warps, 1 CTA / SM, persistent over tiles:
//   warps 0-11      three producer groups of 4 warps (k-block kb -> group kb % 3, TMEM A stage
//                   = global k-block count % 4), then epilogue
//   warp 12         TMA producer (W tiles)
//   warp 13         TMEM allocation + single-thread tcgen05.mma issue / commit
// TMEM (512 columns): [0,128) D_main, [128,256) D_corr, [256,512) 4 A stages x (32 hi + 32 lo).
#pragma once
#include "common.cuh"
#include "tc05.cuh"
#include <stdlib.h>

namespace tcg {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int NGROUPS = 3;                         // producer groups == TMEM A stages (kb % 3)
constexpr int TILE_B_BYTES = BN * BK * 4;          // 16 KB per hi or lo tile
constexpr int STAGE_B_BYTES = 2 * TILE_B_BYTES;    // [W_hi tile | W_lo tile] = 256 rows x 128 B
constexpr int EW = 16;                             // epilogue unit: 16 accumulator columns
constexpr int TMEM_COLS = 512;
constexpr int COL_MAIN = 0, COL_CORR = 128, COL_A = 256;   // A stage s: COL_A + 64 s (hi), +32 (lo)
constexpr int PRODUCER_THREADS = 128 * NGROUPS;
constexpr int TMA_WARP = 4 * NGROUPS, MMA_WARP = 4 * NGROUPS + 1;
constexpr int THREADS = PRODUCER_THREADS + 64;
constexpr int MAX_B_STAGES = 6;
// shared memory of the skeleton: W ring (Policy::kStagesB stages) + barriers / TMEM holder
__host__ __device__ constexpr int core_smem(int stages_b) { return stages_b * STAGE_B_BYTES + 256; }

struct Core {
  uint8_t* Bst;        // [kStagesB][hi 16 KB | lo 16 KB]
  uint64_t* b_full;    // [kStagesB]  TMA arrive.expect_tx
  uint64_t* b_empty;   // [kStagesB]  tcgen05.commit
  uint64_t* a_full;    // [NGROUPS]   4 producer-warp arrivals
  uint64_t* a_empty;   // [NGROUPS]   tcgen05.commit
  uint64_t* acc_full;  // [1]
  uint64_t* acc_empty; // [1]
  uint32_t* tmem_holder;
};

__device__ __forceinline__ Core carve(uint8_t* base, int stages_b) {
  Core c;
  c.Bst = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + stages_b * STAGE_B_BYTES);
  c.b_full = bars;
  c.b_empty = c.b_full + MAX_B_STAGES;
  c.a_full = c.b_empty + MAX_B_STAGES;
  c.a_empty = c.a_full + NGROUPS;
  c.acc_full = c.a_empty + NGROUPS;
  c.acc_empty = c.acc_full + 1;
  c.tmem_holder = reinterpret_cast<uint32_t*>(c.acc_empty + 1);
  return c;
}

// Optional phase timers (profiling aid): when a buffer is registered with lnb_debug_set_prof,
// thread 0 of every CTA accumulates clock64() deltas per phase into prof[cta*32 + phase]
// (slots 8 / 9: k-loop / accumulator wait of odd sub-steps, 10: post_epilogue):
//   0 staging issue  1 staging wait  2 U = V^T X   (policy)   3 k-loop  4 pre_epilogue
//   5 wait for the accumulator  6 tcgen05.ld  7 epilogue store
__device__ unsigned long long* g_prof = nullptr;

struct PhaseTimer {          // thread 0 of the CTA only; no-op unless a buffer is registered
  unsigned long long* buf;
  long long t0;
  __device__ __forceinline__ void start(int cta, int tid) {
    buf = (tid == 0 && g_prof) ? g_prof + cta * 32 : nullptr;
    if (buf) t0 = clock64();
  }
  __device__ __forceinline__ void lap(int slot) {
    if (buf) { long long t = clock64(); atomicAdd(&buf[slot], (unsigned long long)(t - t0)); t0 = t; }
  }
};

__device__ __forceinline__ void producers_sync() {   // named barrier 1: all producer threads
  asm volatile("bar.sync 1, %0;" ::"n"(PRODUCER_THREADS) : "memory");
}

// Policy contract (all __device__).  A "step" is one accumulator lifetime: its k-blocks are
// produced / multiplied, then the epilogue hands the 128 x 128 result to the policy.
//   static constexpr int kStagesB                        depth of the W (shared memory) ring: 2 .. 6
//   struct Params;                                       kernel parameter block (by value)
//   static int  num_steps(const Params&, int cta, int ncta)       steps this CTA runs
//   static void decode(const Params&, int cta, int ncta, int it, int& m_tile, int& sub)
//   static int  num_kblocks(const Params&, int sub)
//   static void w_coords(const Params&, int sub, int kb, int& col0, int& row0)   TMA coords of W
//   Policy(const Params&, uint8_t* policy_smem, int tid)  constructed by producer threads only
//   void step_begin(int m_tile, int sub, int kb_first, PhaseTimer&)   may call producers_sync(); kb_first =
//                                                         this thread's first k-block of the step
//   void produce(int sub, int kb, float (&v)[32])         the 32 A values of this thread's row
//   void pre_epilogue(int sub)                            after the step's last produce()
//   void post_epilogue(int sub)                           after the step's last store()
//   void store(int sub, int col, float (&x)[EW])          accumulator columns [col, col+EW) of
//                                                         this thread's row (main + corr summed)
template <class Policy>
__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_hi,
               const __grid_constant__ CUtensorMap map_lo, const typename Policy::Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // 1024-byte alignment (SWIZZLE_128B) by OFFSETTING the __shared__ array -- keeps the
  // shared address space visible to the compiler (LDS/STS instead of generic LD/ST)
  const uint32_t pad = (1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* base = smem_raw + pad;
  constexpr int SB = Policy::kStagesB;
  Core c = carve(base, SB);
  uint8_t* policy_smem = base + core_smem(SB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int nsteps = Policy::num_steps(p, cta, ncta);
  unsigned long long prof_ns0 = 0;                  // whole-CTA wall time / cycles (slots 11, 12)
  long long prof_c0 = 0;
  if (tid == 0 && g_prof) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(prof_ns0));
    prof_c0 = clock64();
  }

  if (warp == TMA_WARP && lane == 0) {
    tc05::tma_prefetch_desc(&map_hi);
    tc05::tma_prefetch_desc(&map_lo);
  }
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < SB; ++s) { tc05::mbar_init(&c.b_full[s], 1); tc05::mbar_init(&c.b_empty[s], 1); }
      for (int s = 0; s < NGROUPS; ++s) { tc05::mbar_init(&c.a_full[s], 4); tc05::mbar_init(&c.a_empty[s], 1); }
      tc05::mbar_init(c.acc_full, 1);
      tc05::mbar_init(c.acc_empty, 4 * NGROUPS);       // one arrival per producer warp
      tc05::fence_barrier_init();
    }
    __syncwarp();
    tc05::tmem_alloc(c.tmem_holder, TMEM_COLS);
  }
  tc05::fence_before_thread_sync();
  __syncthreads();
  tc05::fence_after_thread_sync();
  const uint32_t tmem_base = *c.tmem_holder;

  if (warp < TMA_WARP) {
    // ================================ producers + epilogue ================================
    const int grp = warp >> 2;                       // producer group == pipeline stage
    const int wq = warp & 3;                         // TMEM lane quarter
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    const uint32_t a_hi = lane_addr + COL_A + grp * 64;
    Policy pol(p, policy_smem, tid);
    uint32_t use = 0;                                // uses of this group's stage so far
    uint32_t gk0 = 0;                                // global k-block count at step start
    for (int it = 0; it < nsteps; ++it) {
      int m_tile, sub;
      Policy::decode(p, cta, ncta, it, m_tile, sub);
      const int nkb = Policy::num_kblocks(p, sub);
      // stage (== group) of a k-block is its GLOBAL index % 3, like the TMA / MMA warps count it
      const int kb_first = (grp + NGROUPS - (int)(gk0 % NGROUPS)) % NGROUPS;
      gk0 += nkb;
      PhaseTimer tm;
      tm.start(cta, tid);
      pol.step_begin(m_tile, sub, kb_first, tm);
      for (int kb = kb_first; kb < nkb; kb += NGROUPS, ++use) {
        float v[32];
        if (!(p.dbg & 8)) pol.produce(sub, kb, v);
        tc05::mbar_wait(&c.a_empty[grp], (use & 1u) ^ 1u);
        tc05::fence_after_thread_sync();
        if (!(p.dbg & 1)) {
          uint32_t part[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) part[j] = tc05::tf32_rna_bits(v[j]);
          tc05::tmem_st_32x32(a_hi, part);
#pragma unroll
          for (int j = 0; j < 32; ++j) part[j] = tc05::tf32_rna_bits(v[j] - __uint_as_float(part[j]));
          tc05::tmem_st_32x32(a_hi + 32, part);
          tc05::tmem_wait_st();
        }
        tc05::fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) tc05::mbar_arrive(&c.a_full[grp]);
      }
      tm.lap((sub & 1) ? 8 : 3);
      pol.pre_epilogue(sub);
      tm.lap(4);
      // ---- epilogue: 16-column unit cc belongs to group cc % NGROUPS ----
      tc05::mbar_wait(c.acc_full, (uint32_t)it & 1u);
      tc05::fence_after_thread_sync();
      tm.lap((sub & 1) ? 9 : 5);
      {
        // 8 units over 3 groups (3 / 3 / 2); the loads of the next unit are in flight while the
        // current one is stored (tcgen05.wait::ld waits for everything issued so far)
        uint32_t vm[EW], vc[EW];
        tc05::tmem_ld_32x16(lane_addr + COL_MAIN + grp * EW, vm);
        tc05::tmem_ld_32x16(lane_addr + COL_CORR + grp * EW, vc);
        tc05::tmem_wait_ld();
#pragma unroll 1
        for (int cc = grp; cc < BN / EW; cc += NGROUPS) {
          const int col = cc * EW;
          float x[EW];
#pragma unroll
          for (int j = 0; j < EW; ++j) x[j] = __uint_as_float(vm[j]) + __uint_as_float(vc[j]);
          if (cc + NGROUPS < BN / EW) {
            tc05::tmem_ld_32x16(lane_addr + COL_MAIN + col + NGROUPS * EW, vm);
            tc05::tmem_ld_32x16(lane_addr + COL_CORR + col + NGROUPS * EW, vc);
          }
          tm.lap(6);
          pol.store(sub, col, x);
          tc05::tmem_wait_ld();
          tm.lap(7);
        }
      }
      tc05::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) tc05::mbar_arrive(c.acc_empty);    // accumulators are free again
      pol.post_epilogue(sub);
      tm.lap(10);
    }
  } else if (warp == TMA_WARP) {
    // ================================ TMA producer (W tiles) ==============================
    // the whole warp runs the loop (warp-uniform control flow); one elected lane issues
    uint32_t cnt = 0;
    for (int it = 0; it < nsteps; ++it) {
      int m_tile, sub;
      Policy::decode(p, cta, ncta, it, m_tile, sub);
      const int nkb = Policy::num_kblocks(p, sub);
      for (int kb = 0; kb < nkb; ++kb, ++cnt) {
        int col0, row0;
        Policy::w_coords(p, sub, kb, col0, row0);
        const uint32_t st = cnt % SB;
        tc05::mbar_wait(&c.b_empty[st], ((cnt / SB) & 1u) ^ 1u);
        if (tc05::elect_one()) {
          if (p.dbg & 4) {
            tc05::mbar_arrive(&c.b_full[st]);
          } else {
            uint8_t* dst = c.Bst + st * STAGE_B_BYTES;
            tc05::mbar_arrive_expect_tx(&c.b_full[st], STAGE_B_BYTES);
            tc05::tma_load_2d(dst, &map_hi, &c.b_full[st], col0, row0);
            tc05::tma_load_2d(dst + TILE_B_BYTES, &map_lo, &c.b_full[st], col0, row0);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ MMA issuer ==========================================
    // per k-step of 8:  [D_main | D_corr] += A_hi * [W_hi ; W_lo]^T   (one N = 256 MMA)
    //                    D_corr           += A_lo * W_hi^T            (one N = 128 MMA)
    constexpr uint32_t idesc256 = tc05::umma_idesc_tf32(BM, 2 * BN);
    constexpr uint32_t idesc128 = tc05::umma_idesc_tf32(BM, BN);
    uint32_t cnt = 0;
    const uint32_t d_main = tmem_base + COL_MAIN, d_corr = tmem_base + COL_CORR;
    for (int it = 0; it < nsteps; ++it) {
      int m_tile, sub;
      Policy::decode(p, cta, ncta, it, m_tile, sub);
      const int nkb = Policy::num_kblocks(p, sub);
      tc05::mbar_wait(c.acc_empty, ((uint32_t)it & 1u) ^ 1u);   // previous epilogue done
      tc05::fence_after_thread_sync();
      for (int kb = 0; kb < nkb; ++kb, ++cnt) {
        const uint32_t sb = cnt % SB, sa = cnt % NGROUPS;
        tc05::mbar_wait2(&c.b_full[sb], (cnt / SB) & 1u, &c.a_full[sa], (cnt / NGROUPS) & 1u);
        tc05::fence_after_thread_sync();
        if (tc05::elect_one()) {
          const uint32_t a_hi = tmem_base + COL_A + sa * 64, a_lo = a_hi + 32;
          const uint64_t dB = tc05::umma_desc_kmajor_sw128(tc05::smem_u32(c.Bst + sb * STAGE_B_BYTES));
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
              tc05::umma_tf32_ts(d_main, a_hi + 8 * k, dB + 2 * k, idesc256, acc);
              tc05::umma_tf32_ts(d_corr, a_lo + 8 * k, dB + 2 * k, idesc128, 1u);
            }
          }
          tc05::umma_commit(&c.a_empty[sa]);
          tc05::umma_commit(&c.b_empty[sb]);
          if (kb == nkb - 1) tc05::umma_commit(c.acc_full);
        }
        __syncwarp();
      }
    }
  }

  tc05::fence_before_thread_sync();
  __syncthreads();
  if (warp == MMA_WARP) {
    __syncwarp();
    tc05::tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (tid == 0 && g_prof) {
    unsigned long long ns1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns1));
    atomicAdd(&g_prof[cta * 32 + 11], ns1 - prof_ns0);
    atomicAdd(&g_prof[cta * 32 + 12], (unsigned long long)(clock64() - prof_c0));
    g_prof[cta * 32 + 13] = prof_ns0;               // CTA start (ns), for launch skew
  }
}

// Epilogue helper: bias / ReLU / bounds-checked store of EW accumulator columns of one row.
__device__ __forceinline__ void store_row_chunk(float* orow, int ncols, const float* bias, bool relu,
                                                int col, const float (&x)[EW]) {
  if (orow == nullptr || col >= ncols) return;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(orow + col) & 15) == 0;
#pragma unroll
  for (int j = 0; j < EW; j += 4) {
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = col + j + u;
      float y = x[j + u];
      if (n < ncols) {
        if (bias) y += __ldg(bias + n);
        if (relu) y = fmaxf(y, 0.f);
      }
      o[u] = y;
    }
    if (vec_ok && col + j + 3 < ncols) {
      *reinterpret_cast<float4*>(orow + col + j) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (col + j + u < ncols) orow[col + j + u] = o[u];
    }
  }
}

// ---- host helpers --------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;   // idempotent lookup; benign if two threads race
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Row-major [rows, cols] fp32 weight matrix; box = 32 columns (128 B) x 128 rows; 128 B swizzle.
inline int make_weight_map(CUtensorMap* map, const float* W, int rows, int cols, const char* who,
                           int box_rows = BN) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { lnb::set_err("%s: cuTensorMapEncodeTiled unavailable", who); return LNB_ERR_UNSUPPORTED; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(W), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lnb::set_err("%s: cuTensorMapEncodeTiled failed (CUresult %d)", who, (int)r);
    return LNB_ERR_ARG;
  }
  return LNB_OK;
}

// LNB_DBG=<bits>: pipeline experiments (1 skip TMEM stores, 2 skip MMA issue, 4 skip TMA loads,
// 8 skip produce()).  Results are wrong with any bit set; for profiling only.
inline int debug_flags() {
  const char* e = getenv("LNB_DBG");
  return e ? atoi(e) : 0;
}

inline int sm_count() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

}  // namespace tcg
