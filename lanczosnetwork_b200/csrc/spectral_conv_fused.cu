// Fused spectral graph-convolution layer (reference: model/lanczos_net.py:157-182):
//     msg = [ V diag(f_s) V^T X  (s < S) ] ++ [ L_e X  (e <= E) ];   X' = ReLU(cat(msg) W^T + b)
// as ONE persistent tcgen05 kernel (skeleton: tc_gemm.cuh).  The [B*N, C*D] message matrix of
// the unfused path (204 MB / layer at B=1024) never exists: producer warps compute each
// 128 x 32 message tile on CUDA cores and store it straight into tensor memory, where the
// tensor core multiplies it by the TMA-staged weight tile.
//
// Tile = 128 rows = G graph slots of NS rows (NS = 32/64/128 >= N), so a warp never straddles
// graphs: the right-hand sides X_g / U_g = V_g^T X_g are read from shared memory as warp-wide
// broadcasts.  Edge-type operators are consumed through a per-forward ELL compression of the
// dense L[B,N,N,E+1] (lnb_graph_prepare): the QM8 operators are ~4 % dense and skipping exact
// zeros is exact, so the edge channels cost ~1/10 of the dense product and the kernel is bound
// by the tensor pipe, not by message production.  Long-scale channels use the factored form
// (V * f_s) (V^T X) with the Ritz extent k_eff (zero-padded pairs skipped, also exact).
#include "tc_gemm.cuh"

namespace {

constexpr int KMAX = 32;        // max Ritz pairs held per thread in the U = V^T X prologue

// --------------------------------------------------------------------------------------------
// Per-forward operator compression.
//   ell_val/ell_idx [B, E1, N(t), N(n)]: the t-th non-zero of row n of channel e, stored
//   t-major so a warp (consecutive n) reads consecutive addresses; zero-filled up to
//   ell_max[b,e] = max non-zeros of any row.  qext[b] = {n_eff, k_eff}: Q[b] is zero outside
//   its leading n_eff rows / k_eff columns.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
graph_prepare_kernel(const float* __restrict__ L, const float* __restrict__ Q, int N, int E1, int K,
                     float* __restrict__ ell_val, uint8_t* __restrict__ ell_idx,
                     int32_t* __restrict__ ell_max, int32_t* __restrict__ qext) {
  __shared__ int s_max[64];
  __shared__ int s_ext[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 64) s_max[tid] = 0;
  if (tid < 2) s_ext[tid] = 0;
  __syncthreads();
  const float* Lb = L + (int64_t)b * N * N * E1;
  const int pairs = N * E1;
  // phase 1: compact every (row n, channel e); pair index p = n*E1 + e keeps the E1 channels of
  // one row in adjacent threads -> their loads of L[b,n,i,:] coalesce.
  for (int p0 = 0; p0 < pairs; p0 += 256) {
    const int p = p0 + tid;
    int cnt = 0;
    if (p < pairs) {
      const int n = p / E1, e = p % E1;
      float* val = ell_val + ((int64_t)(b * E1 + e) * N) * N + n;
      uint8_t* idx = ell_idx + ((int64_t)(b * E1 + e) * N) * N + n;
      for (int i = 0; i < N; ++i) {
        float v = Lb[((int64_t)n * N + i) * E1 + e];
        if (v != 0.f) {
          val[(int64_t)cnt * N] = v;
          idx[(int64_t)cnt * N] = (uint8_t)i;
          ++cnt;
        }
      }
      atomicMax(&s_max[e], cnt);
    }
  }
  // Q extents
  const float* Qb = Q + (int64_t)b * N * K;
  int ne = 0, ke = 0;
  for (int i = tid; i < N * K; i += 256) {
    if (Qb[i] != 0.f) {
      ne = max(ne, i / K + 1);
      ke = max(ke, i % K + 1);
    }
  }
  if (ne) atomicMax(&s_ext[0], ne);
  if (ke) atomicMax(&s_ext[1], ke);
  __syncthreads();
  // phase 2: zero-fill the tail of every row up to the channel maximum (recount = cheap)
  for (int p0 = 0; p0 < pairs; p0 += 256) {
    const int p = p0 + tid;
    if (p < pairs) {
      const int n = p / E1, e = p % E1;
      int cnt = 0;
      for (int i = 0; i < N; ++i) cnt += (Lb[((int64_t)n * N + i) * E1 + e] != 0.f) ? 1 : 0;
      float* val = ell_val + ((int64_t)(b * E1 + e) * N) * N + n;
      uint8_t* idx = ell_idx + ((int64_t)(b * E1 + e) * N) * N + n;
      for (int t = cnt; t < s_max[e]; ++t) {
        val[(int64_t)t * N] = 0.f;
        idx[(int64_t)t * N] = 0;
      }
    }
  }
  if (tid < E1) ell_max[b * E1 + tid] = s_max[tid];
  if (tid < 2) qext[b * 2 + tid] = s_ext[tid];
}

// --------------------------------------------------------------------------------------------
struct SpectralPolicy {
  struct Params {
    const float* X;         // [B, N, Din]
    const float* Q;         // [B, N, K]
    const float* coeff;     // [B, K, S]
    const float* ell_val;   // [B, E1, N, N]
    const uint8_t* ell_idx; // [B, E1, N, N]
    const int32_t* ell_max; // [B, E1]
    const int32_t* qext;    // [B, 2]
    const float* bias;      // [H]
    float* out;             // [B, N, H]
    int B, N, Din, E1, K, S, H, relu, NS;
    int TCAP;               // ELL entries per (row, channel) staged in shared memory
  };
  static __device__ __forceinline__ int n_tiles(const Params& p) { return (p.H + tcg::BN - 1) / tcg::BN; }
  static __device__ __forceinline__ int num_tiles(const Params& p) {
    const int G = tcg::BM / p.NS;
    return ((p.B + G - 1) / G) * n_tiles(p);
  }
  static __device__ __forceinline__ int num_kblocks(const Params& p) {
    return (p.S + p.E1) * p.Din / tcg::BK;
  }

  const Params& p;
  const int tid, r, g, n;   // row in tile, graph slot, node
  const int G, XP, KP;      // slots per tile, padded X row stride, padded Q row stride
  float* Xs;                // [G][N][XP]
  float* Us;                // [G][K][Din]
  float* Qs;                // [G][N][KP]
  float* Fs;                // [G][K][S]
  int* Es;                  // [G][E1 + 2]  ell_max per channel, n_eff, k_eff
  float* Ev;                // [G][E1][TCAP][NS] staged ELL values
  uint8_t* Ei;              // [G][E1][TCAP][NS] staged ELL column indices
  int b;                    // graph of this row (or -1)
  bool valid;

  __device__ SpectralPolicy(const Params& p_, uint8_t* smem, int tid_)
      : p(p_), tid(tid_), r(tid_ & 127), g((tid_ & 127) / p_.NS), n((tid_ & 127) % p_.NS),
        G(tcg::BM / p_.NS), XP(p_.Din + 4), KP(p_.K | 1), b(-1), valid(false) {
    Xs = reinterpret_cast<float*>(smem);
    Us = Xs + (size_t)G * p.N * XP;
    Qs = Us + (size_t)G * p.K * p.Din;
    Fs = Qs + (size_t)G * p.N * KP;
    Es = reinterpret_cast<int*>(Fs + (size_t)G * p.K * p.S);
    Ev = reinterpret_cast<float*>(Es + (size_t)G * (p.E1 + 2));
    Ei = reinterpret_cast<uint8_t*>(Ev + (size_t)G * p.E1 * p.TCAP * p.NS);
  }

  static size_t smem_bytes(int N, int Din, int K, int S, int E1, int NS) {
    const int G = tcg::BM / NS;
    size_t fl = (size_t)G * N * (Din + 4) + (size_t)G * K * Din + (size_t)G * N * (K | 1) +
                (size_t)G * K * S;
    return fl * 4 + (size_t)G * (E1 + 2) * 4 + 16;
  }
  static size_t ell_stage_bytes(int E1, int NS, int tcap) {
    return (size_t)(tcg::BM / NS) * E1 * tcap * NS * 5;
  }

  __device__ void tile_begin(int m_tile, int /*n_tile*/) {
    const int b0 = m_tile * G;
    b = b0 + g;
    valid = (b < p.B) && (n < p.N);
    if (b >= p.B) b = -1;
    tcg::producers_sync();                       // previous tile's readers are done
    // ---- stage X, Q, filter coefficients, extents of the G graphs --------------------------
    const int dv = p.Din / 4;
    for (int e = tid; e < G * p.N * dv; e += tcg::PRODUCER_THREADS) {
      const int gg = e / (p.N * dv), rem = e % (p.N * dv);
      const int nn = rem / dv, q4 = rem % dv;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + gg < p.B)
        v = __ldg(reinterpret_cast<const float4*>(p.X + ((int64_t)(b0 + gg) * p.N + nn) * p.Din) + q4);
      *reinterpret_cast<float4*>(Xs + ((size_t)gg * p.N + nn) * XP + 4 * q4) = v;
    }
    for (int e = tid; e < G * p.N * p.K; e += tcg::PRODUCER_THREADS) {
      const int gg = e / (p.N * p.K), rem = e % (p.N * p.K);
      const int nn = rem / p.K, kk = rem % p.K;
      Qs[((size_t)gg * p.N + nn) * KP + kk] =
          (b0 + gg < p.B) ? __ldg(p.Q + ((int64_t)(b0 + gg) * p.N + nn) * p.K + kk) : 0.f;
    }
    for (int e = tid; e < G * p.K * p.S; e += tcg::PRODUCER_THREADS) {
      const int gg = e / (p.K * p.S);
      Fs[e] = (b0 + gg < p.B) ? __ldg(p.coeff + (int64_t)(b0 + gg) * p.K * p.S + e % (p.K * p.S)) : 0.f;
    }
    for (int e = tid; e < G * (p.E1 + 2); e += tcg::PRODUCER_THREADS) {
      const int gg = e / (p.E1 + 2), w = e % (p.E1 + 2);
      int v = 0;
      if (b0 + gg < p.B)
        v = (w < p.E1) ? p.ell_max[(b0 + gg) * p.E1 + w] : p.qext[(b0 + gg) * 2 + (w - p.E1)];
      Es[e] = v;
    }
    // ELL rows of the tile's graphs (first TCAP entries per row/channel), zero beyond the
    // channel maximum so the inner loop needs no per-lane guard
    for (int e = tid; e < G * p.E1 * p.TCAP * p.NS; e += tcg::PRODUCER_THREADS) {
      const int nn = e % p.NS, t = (e / p.NS) % p.TCAP;
      const int ch = (e / (p.NS * p.TCAP)) % p.E1, gg = e / (p.NS * p.TCAP * p.E1);
      float v = 0.f;
      int ix = 0;
      if (b0 + gg < p.B && nn < p.N && t < __ldg(p.ell_max + (b0 + gg) * p.E1 + ch)) {
        const int64_t off = (((int64_t)(b0 + gg) * p.E1 + ch) * p.N + t) * p.N + nn;
        v = __ldg(p.ell_val + off);
        ix = __ldg(p.ell_idx + off);
      }
      Ev[e] = v;
      Ei[e] = (uint8_t)ix;
    }
    tcg::producers_sync();
    // ---- U_g = Q_g^T X_g  (K x Din per graph): thread <-> (graph, column), all k in registers
    for (int pr = tid; pr < G * p.Din; pr += tcg::PRODUCER_THREADS) {
      const int gg = pr / p.Din, d = pr % p.Din;
      const int n_eff = Es[gg * (p.E1 + 2) + p.E1];
      float acc[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
      const float* xs = Xs + (size_t)gg * p.N * XP + d;
      const float* qs = Qs + (size_t)gg * p.N * KP;
      for (int nn = 0; nn < n_eff; ++nn) {
        const float x = xs[(size_t)nn * XP];
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < p.K) acc[k] = fmaf(qs[nn * KP + k], x, acc[k]);
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < p.K) Us[((size_t)gg * p.K + k) * p.Din + d] = acc[k];
    }
    tcg::producers_sync();
  }

  __device__ __forceinline__ void produce(int kb, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    const int j0 = kb * tcg::BK;
    const int c = j0 / p.Din, d0 = j0 % p.Din;
    if (b < 0) return;                            // warp-uniform: a warp never straddles graphs
    const int* es = Es + g * (p.E1 + 2);
    if (c < p.S) {
      // long scale s = c:  row n of (Q * f_s) times U_g[:, d0:d0+32]
      const int k_eff = es[p.E1 + 1];
      const float* qrow = Qs + ((size_t)g * p.N + (n < p.N ? n : 0)) * KP;
      const float* f = Fs + (size_t)g * p.K * p.S + c;
      const float* u = Us + (size_t)g * p.K * p.Din + d0;
#pragma unroll 2
      for (int i = 0; i < k_eff; ++i) {
        const float a = (n < p.N) ? qrow[i] * f[i * p.S] : 0.f;
        const float4* u4 = reinterpret_cast<const float4*>(u + (size_t)i * p.Din);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = u4[q];
          v[4 * q + 0] = fmaf(a, t.x, v[4 * q + 0]);
          v[4 * q + 1] = fmaf(a, t.y, v[4 * q + 1]);
          v[4 * q + 2] = fmaf(a, t.z, v[4 * q + 2]);
          v[4 * q + 3] = fmaf(a, t.w, v[4 * q + 3]);
        }
      }
    } else {
      // edge type e = c - S: sparse row of L_e (ELL) times X_g[:, d0:d0+32]
      const int e = c - p.S;
      const int tmax = es[e];
      const int nn = (n < p.N) ? n : 0;
      const float* val = p.ell_val + ((int64_t)(b * p.E1 + e) * p.N) * p.N + nn;
      const uint8_t* idx = p.ell_idx + ((int64_t)(b * p.E1 + e) * p.N) * p.N + nn;
      const float* xs = Xs + (size_t)g * p.N * XP + d0;
      const int ts = tmax < p.TCAP ? tmax : p.TCAP;
      const float* ev = Ev + ((size_t)(g * p.E1 + e) * p.TCAP) * p.NS + n;
      const uint8_t* ei = Ei + ((size_t)(g * p.E1 + e) * p.TCAP) * p.NS + n;
#pragma unroll 2
      for (int t = 0; t < ts; ++t) {
        const float a = ev[t * p.NS];
        const int i = ei[t * p.NS];
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
      for (int t = ts; t < tmax; ++t) {            // rows denser than the staged capacity
        const float a = (n < p.N) ? __ldg(val + (int64_t)t * p.N) : 0.f;
        const int i = __ldg(idx + (int64_t)t * p.N);
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
  __device__ __forceinline__ void tile_end() {}
  static __device__ __forceinline__ int w_row0(const Params&, int n_tile) { return n_tile * tcg::BN; }
  __device__ __forceinline__ float* out_ptr(int n_tile) const {
    return valid ? p.out + ((int64_t)b * p.N + n) * p.H + n_tile * tcg::BN : nullptr;
  }
  __device__ __forceinline__ int cols_valid(int n_tile) const {
    const int left = p.H - n_tile * tcg::BN;
    return left < tcg::BN ? left : tcg::BN;
  }
  __device__ __forceinline__ const float* bias_ptr(int n_tile) const {
    return p.bias ? p.bias + n_tile * tcg::BN : nullptr;
  }
  __device__ __forceinline__ bool relu() const { return p.relu != 0; }
};

}  // namespace

extern "C" {

int lnb_graph_prepare(lnb_stream_t stream, const float* L, const float* Q, int B, int N, int E1,
                      int K, float* ell_val, uint8_t* ell_idx, int32_t* ell_max, int32_t* qext) {
  LNB_REQUIRE(L && Q && ell_val && ell_idx && ell_max && qext, "graph_prepare: null pointer");
  LNB_REQUIRE(B >= 0 && N >= 1 && N <= 255 && E1 >= 1 && E1 <= 64 && K >= 1,
              "graph_prepare: bad dims B=%d N=%d E1=%d K=%d", B, N, E1, K);
  if (B == 0) return LNB_OK;
  graph_prepare_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(L, Q, N, E1, K, ell_val, ell_idx,
                                                            ell_max, qext);
  lnb::count_launch();
  return lnb::finish_launch("graph_prepare");
}

int lnb_spectral_conv_fused(lnb_stream_t stream, const float* X, const float* Q, const float* coeff,
                            const float* ell_val, const uint8_t* ell_idx, const int32_t* ell_max,
                            const int32_t* qext, const float* W_hi, const float* W_lo,
                            const float* bias, int B, int N, int Din, int E1, int K, int S, int H,
                            int relu, float* out) {
  LNB_REQUIRE(X && Q && coeff && ell_val && ell_idx && ell_max && qext && W_hi && W_lo && out,
              "spectral_conv_fused: null pointer");
  LNB_REQUIRE(B >= 0 && N >= 1 && Din >= 1 && E1 >= 1 && K >= 1 && S >= 0 && H >= 1,
              "spectral_conv_fused: bad dims");
  if (N > 128 || Din % 32 != 0 || K > KMAX || H % 4 != 0) {
    lnb::set_err("spectral_conv_fused: unsupported shape N=%d Din=%d K=%d H=%d "
                 "(needs N<=128, Din%%32==0, K<=%d, H%%4==0)", N, Din, K, H, KMAX);
    return LNB_ERR_UNSUPPORTED;
  }
  if (B == 0) return LNB_OK;
  const int NS = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  size_t smem = tcg::CORE_SMEM + 1024 + SpectralPolicy::smem_bytes(N, Din, K, S, E1, NS);
  int tcap = 0;
  if (smem <= 227 * 1024) {
    while (tcap < 8 && smem + SpectralPolicy::ell_stage_bytes(E1, NS, tcap + 1) <= 227 * 1024) ++tcap;
    smem += SpectralPolicy::ell_stage_bytes(E1, NS, tcap);
  }
  if (smem > 227 * 1024) {
    lnb::set_err("spectral_conv_fused: tile state (N=%d, Din=%d, K=%d) needs %zu B of shared memory",
                 N, Din, K, smem);
    return LNB_ERR_UNSUPPORTED;
  }
  const int Kw = (S + E1) * Din;
  CUtensorMap map_hi, map_lo;
  int rc = tcg::make_weight_map(&map_hi, W_hi, H, Kw, "spectral_conv_fused");
  if (rc != LNB_OK) return rc;
  rc = tcg::make_weight_map(&map_lo, W_lo, H, Kw, "spectral_conv_fused");
  if (rc != LNB_OK) return rc;
  auto kern = tcg::tc_gemm_kernel<SpectralPolicy>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SpectralPolicy::Params p{X, Q, coeff, ell_val, ell_idx, ell_max, qext, bias, out,
                           B, N, Din, E1, K, S, H, relu, NS, tcap};
  const int G = tcg::BM / NS;
  const int tiles = lnb::ceil_div(B, G) * lnb::ceil_div(H, tcg::BN);
  const int grid = tiles < tcg::sm_count() ? tiles : tcg::sm_count();
  kern<<<grid, tcg::THREADS, smem, (cudaStream_t)stream>>>(map_hi, map_lo, p);
  lnb::count_launch();
  return lnb::finish_launch("spectral_conv_fused");
}

}  // extern "C"
