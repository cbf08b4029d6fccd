// Dense layer on the 5th-generation tensor cores (tcgen05, sm_100a):
//     C[M,N] = act( A[M,K] @ W[N,K]^T + bias )          fp32 in, fp32 out, 3xTF32
// Row-loading policy of the persistent skeleton in tc_gemm.cuh: each producer thread owns one
// row of the 128-row tile, streams 32 consecutive K-values per k-block (8 x LDG.128, one of
// its k-blocks ahead) and hands them to the skeleton, which splits them into tf32 hi/lo and
// stores them straight into tensor memory.
// Replaces the nn.Linear call sites of the reference (model/lanczos_net.py:112,181,186;
// model/ada_lanczos_net.py:54-63 via :274).
#include "tc_gemm.cuh"

namespace {

struct RowLoadPolicy {
  static constexpr int kStagesB = 3;
  // groups > 1: block-diagonal ("grouped") layer -- column block g of A [M, groups*K] times
  // W_g [N, K] (stacked [groups*N, K]) into column block g of C [M, groups*N].
  struct Params {
    const float* A;
    const float* bias;
    float* C;
    int M, N, K, relu, groups;
    int dbg;                // debug experiment flags (LNB_DBG), 0 in production
  };
  static __device__ __forceinline__ int tiles_per_group(const Params& p) { return (p.N + tcg::BN - 1) / tcg::BN; }
  static __device__ __forceinline__ int n_tiles(const Params& p) { return p.groups * tiles_per_group(p); }
  static __device__ __forceinline__ int num_tiles(const Params& p) {
    return ((p.M + tcg::BM - 1) / tcg::BM) * n_tiles(p);
  }
  static __device__ __forceinline__ int num_steps(const Params& p, int cta, int ncta) {
    const int t = num_tiles(p);
    return t > cta ? (t - cta + ncta - 1) / ncta : 0;
  }
  static __device__ __forceinline__ void decode(const Params& p, int cta, int ncta, int it,
                                                int& m_tile, int& sub) {
    const int tile = cta + it * ncta, nt = n_tiles(p);
    m_tile = tile / nt;
    sub = tile % nt;
  }
  static __device__ __forceinline__ int num_kblocks(const Params& p, int) { return (p.K + tcg::BK - 1) / tcg::BK; }
  static __device__ __forceinline__ int w_row0(const Params& p, int sub) {
    const int tpg = tiles_per_group(p);
    return (sub / tpg) * p.N + (sub % tpg) * tcg::BN;
  }
  static __device__ __forceinline__ void w_coords(const Params& p, int sub, int kb, int& col0, int& row0) {
    col0 = kb * tcg::BK;
    row0 = w_row0(p, sub);
  }

  const Params& p;
  const int r, grp;
  const int lda, ldc;
  const float* arow;
  bool row_ok;
  int row;
  float cur[32];

  __device__ RowLoadPolicy(const Params& p_, uint8_t*, int tid)
      : p(p_), r(tid & 127), grp(tid >> 7), lda(p_.groups * p_.K), ldc(p_.groups * p_.N),
        arow(nullptr), row_ok(false), row(0) {}

  __device__ __forceinline__ void load(int kb, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kb * tcg::BK + 4 * j;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok && k < p.K) t = __ldg(reinterpret_cast<const float4*>(arow + k));
      v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  }
  __device__ __forceinline__ void step_begin(int m_tile, int sub, int kb_first, tcg::PhaseTimer&) {
    row = m_tile * tcg::BM + r;
    row_ok = row < p.M;
    arow = p.A + (int64_t)(row_ok ? row : 0) * lda + (sub / tiles_per_group(p)) * p.K;
    load(kb_first, cur);
  }
  __device__ __forceinline__ void produce(int, int kb, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = cur[j];
    load(kb + tcg::NGROUPS, cur);   // this group's next k-block (zeros past K)
  }
  __device__ __forceinline__ void pre_epilogue(int) {}
  __device__ __forceinline__ void post_epilogue(int) {}
  __device__ __forceinline__ void store(int sub, int col, const float (&x)[tcg::EW]) {
    const int left = p.N - (sub % tiles_per_group(p)) * tcg::BN;
    const int w0 = w_row0(p, sub);
    tcg::store_row_chunk(row_ok ? p.C + (int64_t)row * ldc + w0 : nullptr,
                         left < tcg::BN ? left : tcg::BN, p.bias ? p.bias + w0 : nullptr,
                         p.relu != 0, col, x);
  }
};

constexpr size_t SMEM_BYTES = tcg::core_smem(RowLoadPolicy::kStagesB) + 1024;

}  // namespace

static int launch_linear(lnb_stream_t stream, const float* A, const float* W_hi, const float* W_lo,
                         const float* bias, int M, int N, int K, int groups, int relu, float* C,
                         const char* who) {
  LNB_REQUIRE(A && W_hi && W_lo && C, "%s: null pointer", who);
  LNB_REQUIRE(M >= 0 && N >= 1 && K >= 1 && groups >= 1, "%s: bad dims M=%d N=%d K=%d groups=%d",
              who, M, N, K, groups);
  LNB_REQUIRE(K % 4 == 0, "%s: K=%d must be a multiple of 4 (16-byte rows)", who, K);
  LNB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W_hi & 15) == 0 && ((uintptr_t)W_lo & 15) == 0,
              "%s: A / W must be 16-byte aligned", who);
  if (M == 0) return LNB_OK;
  CUtensorMap map_hi, map_lo;
  int rc = tcg::make_weight_map(&map_hi, W_hi, groups * N, K, who);
  if (rc != LNB_OK) return rc;
  rc = tcg::make_weight_map(&map_lo, W_lo, groups * N, K, who);
  if (rc != LNB_OK) return rc;
  auto kern = tcg::tc_gemm_kernel<RowLoadPolicy>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  RowLoadPolicy::Params p{A, bias, C, M, N, K, relu, groups, tcg::debug_flags()};
  const int tiles = lnb::ceil_div(M, tcg::BM) * lnb::ceil_div(N, tcg::BN) * groups;
  const int grid = tiles < tcg::sm_count() ? tiles : tcg::sm_count();
  kern<<<grid, tcg::THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(map_hi, map_lo, p);
  lnb::count_launch();
  return lnb::finish_launch(who);
}

extern "C" {

int lnb_linear_tf32x3(lnb_stream_t stream, const float* A, const float* W_hi, const float* W_lo,
                      const float* bias, int M, int N, int K, int relu, float* C) {
  return launch_linear(stream, A, W_hi, W_lo, bias, M, N, K, 1, relu, C, "linear_tf32x3");
}

int lnb_linear_tf32x3_grouped(lnb_stream_t stream, const float* A, const float* W_hi,
                              const float* W_lo, const float* bias, int M, int groups, int N, int K,
                              int relu, float* C) {
  return launch_linear(stream, A, W_hi, W_lo, bias, M, N, K, groups, relu, C,
                       "linear_tf32x3_grouped");
}

}  // extern "C"
