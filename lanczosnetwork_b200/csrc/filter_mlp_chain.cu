// Ritz-value filter MLPs of ALL layers in one persistent tcgen05 kernel
// (reference: model/lanczos_net.py:47-58 the per-layer Sequential, :109-113 its application to
// the B*K rows of Ritz-value powers).  The four Linear stages of one (row tile, layer) item run
// back to back as four accumulator lifetimes of the tc_gemm.cuh skeleton; the 128 x 128
// activations stay in shared memory between stages (the previous epilogue writes bias + ReLU
// output there, the next stage's producers read it as their A operand), so nothing but the
// final [rows, S] coefficients touches HBM.  Only rows (graph, k) with k < k_eff(graph) are
// evaluated: zero-padded Ritz pairs multiply zero Ritz vectors downstream (exact).
#include "tc_gemm.cuh"

namespace {

struct MlpChainPolicy {
  static constexpr int kStagesB = 3;
  struct Params {
    const float* table;     // [Rall, S]  powers of the Ritz values
    const int32_t* rowmap;  // [Rall]     compact list of rows to evaluate (nullptr: all rows)
    const int32_t* nrows;   // [1]        number of valid entries in rowmap (nullptr: Rall)
    const float* bias_all;  // [L * (3*Hd + S)]
    float* coeff;           // [L, Rall, S]
    int Rall, L, S, Hd;
    int dbg;
  };
  static __device__ __forceinline__ int rows(const Params& p) { return p.nrows ? __ldg(p.nrows) : p.Rall; }
  static __device__ __forceinline__ int num_steps(const Params& p, int cta, int ncta) {
    const int items = ((rows(p) + tcg::BM - 1) / tcg::BM) * p.L;
    return 4 * (items > cta ? (items - cta + ncta - 1) / ncta : 0);
  }
  static __device__ __forceinline__ void decode(const Params& p, int cta, int ncta, int it,
                                                int& m_tile, int& sub) {
    const int item = cta + (it >> 2) * ncta;
    m_tile = item / p.L;
    sub = (item % p.L) * 4 + (it & 3);          // (layer, stage)
  }
  static __device__ __forceinline__ int num_kblocks(const Params& p, int sub) {
    return (sub & 3) == 0 ? 1 : p.Hd / tcg::BK;
  }
  static __device__ __forceinline__ int row_off(const Params& p, int sub) {
    return (sub >> 2) * (3 * p.Hd + p.S) + (sub & 3) * p.Hd;
  }
  static __device__ __forceinline__ void w_coords(const Params& p, int sub, int kb, int& col0, int& row0) {
    col0 = kb * tcg::BK;
    row0 = row_off(p, sub);
  }

  const Params& p;
  const int r, AP;
  float* Act;               // [128][Hd + 4]
  int src;                  // dense row index b*K + k of this thread's row (or -1)

  __device__ MlpChainPolicy(const Params& p_, uint8_t* smem, int tid)
      : p(p_), r(tid & 127), AP(p_.Hd + 4), Act(reinterpret_cast<float*>(smem)), src(-1) {}
  static size_t smem_bytes(int Hd) { return (size_t)tcg::BM * (Hd + 4) * 4; }

  __device__ __forceinline__ void step_begin(int m_tile, int sub, int, tcg::PhaseTimer&) {
    tcg::producers_sync();      // the previous stage's activations are complete / fully consumed
    if ((sub & 3) == 0) {
      const int i = m_tile * tcg::BM + r;
      src = -1;
      if (i < rows(p)) src = p.rowmap ? __ldg(p.rowmap + i) : i;
    }
  }
  __device__ __forceinline__ void produce(int sub, int kb, float (&v)[32]) {
    if ((sub & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = (src >= 0 && j < p.S) ? __ldg(p.table + (int64_t)src * p.S + j) : 0.f;
      return;
    }
    const float4* a4 = reinterpret_cast<const float4*>(Act + (size_t)r * AP + kb * tcg::BK);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = a4[q];
      v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  }
  __device__ __forceinline__ void pre_epilogue(int) {}
  __device__ __forceinline__ void post_epilogue(int) {}
  __device__ __forceinline__ void store(int sub, int col, float (&x)[tcg::EW]) {
    const int stage = sub & 3;
    const float* bias = p.bias_all + row_off(p, sub);
    if (stage < 3) {
      if (col >= p.Hd) return;
      float4* o4 = reinterpret_cast<float4*>(Act + (size_t)r * AP + col);
#pragma unroll
      for (int q = 0; q < tcg::EW / 4; ++q) {
        float y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = col + 4 * q + u;
          y[u] = (c < p.Hd) ? fmaxf(x[4 * q + u] + __ldg(bias + c), 0.f) : 0.f;
        }
        o4[q] = make_float4(y[0], y[1], y[2], y[3]);
      }
      return;
    }
    if (src < 0 || col >= p.S) return;
    float* dst = p.coeff + ((int64_t)(sub >> 2) * p.Rall + src) * p.S;
#pragma unroll
    for (int j = 0; j < tcg::EW; ++j)
      if (col + j < p.S) dst[col + j] = x[j] + __ldg(bias + col + j);
  }
};

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
  const size_t smem = tcg::core_smem(MlpChainPolicy::kStagesB) + 1024 + MlpChainPolicy::smem_bytes(Hd);
  auto kern = tcg::tc_gemm_kernel<MlpChainPolicy>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  MlpChainPolicy::Params p{table, rowmap, nrows, bias_all, coeff, Rall, L, S, Hd, tcg::debug_flags()};
  const int items = lnb::ceil_div(Rall, tcg::BM) * L;
  const int grid = items < tcg::sm_count() ? items : tcg::sm_count();
  kern<<<grid, tcg::THREADS, smem, (cudaStream_t)stream>>>(map_hi, map_lo, p);
  lnb::count_launch();
  return lnb::finish_launch("ritz_filter_mlp");
}

}  // extern "C"
