// Dense layer on the 5th-generation tensor cores (tcgen05, sm_100a):
//     C[M,N] = act( A[M,K] @ W[N,K]^T + bias )          fp32 in, fp32 out
// computed as three TF32 products per k-step (A_lo*W_hi + A_hi*W_lo + A_hi*W_hi) accumulated
// in fp32 in tensor memory, which recovers fp32-grade accuracy from 10-bit-mantissa operands.
//
// Design (one 128x128 output tile per CTA, two CTAs resident per SM):
//   warps 0-3  A producers, then epilogue.  Thread r owns tile row r == TMEM lane r: it loads
//              32 consecutive K-values of its row (8 x LDG.128, one k-block ahead), splits them
//              into tf32 hi/lo in registers and writes them with tcgen05.st straight into the
//              A-operand region of tensor memory -- no shared-memory staging or swizzle for A.
//   warp 4     TMA producer: W_hi / W_lo 128x32 tiles (SWIZZLE_128B) -> shared memory ring.
//   warp 5     allocates TMEM; one elected lane issues tcgen05.mma (A from TMEM, B from smem
//              descriptors) and tcgen05.commit to recycle the A / B stages.
// Replaces the nn.Linear call sites of the reference (model/lanczos_net.py:112,181,186;
// model/ada_lanczos_net.py:54-63 via :274).
#include "common.cuh"
#include "tc05.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int NB_STAGES = 3;   // shared-memory ring for W tiles
constexpr int NA_STAGES = 2;   // tensor-memory ring for the A operand
constexpr int TILE_B_BYTES = BN * BK * 4;          // 16 KB
constexpr int TMEM_COLS = 256;                     // 128 accumulator + 2 x (32 hi + 32 lo)
constexpr int A_COL0 = 128;
constexpr int THREADS = 192;
constexpr size_t SMEM_BYTES = 2 * NB_STAGES * TILE_B_BYTES + 256 + 1024;  // tiles + barriers + align

__device__ __forceinline__ void load_a_block(const float* __restrict__ arow, bool row_ok, int kb,
                                             int K, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int k = kb * BK + 4 * j;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok && k < K) t = __ldg(reinterpret_cast<const float4*>(arow + k));
    v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}

__global__ void __launch_bounds__(THREADS, 2)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap map_hi,
                     const __grid_constant__ CUtensorMap map_lo, const float* __restrict__ A,
                     const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K,
                     int relu) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* Bhi = base;
  uint8_t* Blo = base + NB_STAGES * TILE_B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * NB_STAGES * TILE_B_BYTES);
  uint64_t* b_full = bars;                     // [NB_STAGES]
  uint64_t* b_empty = b_full + NB_STAGES;      // [NB_STAGES]
  uint64_t* a_full = b_empty + NB_STAGES;      // [NA_STAGES]
  uint64_t* a_empty = a_full + NA_STAGES;      // [NA_STAGES]
  uint64_t* accum_full = a_empty + NA_STAGES;  // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nkb = (K + BK - 1) / BK;

  if (warp == 4 && lane == 0) {
    tc05::tma_prefetch_desc(&map_hi);
    tc05::tma_prefetch_desc(&map_lo);
  }
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < NB_STAGES; ++s) { tc05::mbar_init(&b_full[s], 1); tc05::mbar_init(&b_empty[s], 1); }
      for (int s = 0; s < NA_STAGES; ++s) { tc05::mbar_init(&a_full[s], 128); tc05::mbar_init(&a_empty[s], 1); }
      tc05::mbar_init(accum_full, 1);
      tc05::fence_barrier_init();
    }
    __syncwarp();
    tc05::tmem_alloc(tmem_holder, TMEM_COLS);
  }
  tc05::fence_before_thread_sync();
  __syncthreads();
  tc05::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < 4) {
    // ------------------------------ A producer -------------------------------------------
    const int row = m0 + warp * 32 + lane;
    const bool row_ok = row < M;
    const float* arow = A + (int64_t)(row_ok ? row : 0) * K;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float cur[32];
    load_a_block(arow, row_ok, 0, K, cur);
    for (int kb = 0; kb < nkb; ++kb) {
      float nxt[32];
      if (kb + 1 < nkb) load_a_block(arow, row_ok, kb + 1, K, nxt);
      const int sa = kb % NA_STAGES;
      const uint32_t pha = (uint32_t)(kb / NA_STAGES) & 1u;
      tc05::mbar_wait(&a_empty[sa], pha ^ 1u);
      tc05::fence_after_thread_sync();
      uint32_t part[32];
      const uint32_t a_hi = lane_addr + A_COL0 + sa * 64;
#pragma unroll
      for (int j = 0; j < 32; ++j) part[j] = tc05::tf32_rna_bits(cur[j]);
      tc05::tmem_st_32x32(a_hi, part);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        part[j] = tc05::tf32_rna_bits(cur[j] - __uint_as_float(part[j]));
      tc05::tmem_st_32x32(a_hi + 32, part);
      tc05::tmem_wait_st();
      tc05::fence_before_thread_sync();
      tc05::mbar_arrive(&a_full[sa]);
      if (kb + 1 < nkb) {
#pragma unroll
        for (int j = 0; j < 32; ++j) cur[j] = nxt[j];
      }
    }
    // ------------------------------ epilogue ---------------------------------------------
    tc05::mbar_wait(accum_full, 0u);
    tc05::fence_after_thread_sync();
    float* crow = C + (int64_t)(row_ok ? row : 0) * N;
    const bool vec_ok = (N % 4) == 0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tc05::tmem_ld_32x32(lane_addr + c * 32, v);
      tc05::tmem_wait_ld();
      const int nb = n0 + c * 32;
      if (row_ok && nb < N) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int n = nb + j + u;
            float x = __uint_as_float(v[j + u]);
            if (n < N) {
              if (bias) x += __ldg(bias + n);
              if (relu) x = fmaxf(x, 0.f);
            }
            o[u] = x;
          }
          if (vec_ok && nb + j + 3 < N) {
            *reinterpret_cast<float4*>(crow + nb + j) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (nb + j + u < N) crow[nb + j + u] = o[u];
          }
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------ TMA producer (W tiles) -------------------------------
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int sb = kb % NB_STAGES;
        const uint32_t phb = (uint32_t)(kb / NB_STAGES) & 1u;
        tc05::mbar_wait(&b_empty[sb], phb ^ 1u);
        tc05::mbar_arrive_expect_tx(&b_full[sb], 2 * TILE_B_BYTES);
        tc05::tma_load_2d(Bhi + sb * TILE_B_BYTES, &map_hi, &b_full[sb], kb * BK, n0);
        tc05::tma_load_2d(Blo + sb * TILE_B_BYTES, &map_lo, &b_full[sb], kb * BK, n0);
      }
    }
  } else {
    // ------------------------------ MMA issuer -------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = tc05::umma_idesc_tf32(BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int sb = kb % NB_STAGES, sa = kb % NA_STAGES;
        const uint32_t phb = (uint32_t)(kb / NB_STAGES) & 1u;
        const uint32_t pha = (uint32_t)(kb / NA_STAGES) & 1u;
        tc05::mbar_wait(&b_full[sb], phb);
        tc05::mbar_wait(&a_full[sa], pha);
        tc05::fence_after_thread_sync();
        const uint32_t a_hi = tmem_base + A_COL0 + sa * 64;
        const uint32_t a_lo = a_hi + 32;
        const uint64_t dhi = tc05::umma_desc_kmajor_sw128(tc05::smem_u32(Bhi + sb * TILE_B_BYTES));
        const uint64_t dlo = tc05::umma_desc_kmajor_sw128(tc05::smem_u32(Blo + sb * TILE_B_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          // 8 tf32 = 32 bytes per k-step: +2 in the descriptor's 16-byte address units
          tc05::umma_tf32_ts(tmem_base, a_lo + 8 * k, dhi + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc05::umma_tf32_ts(tmem_base, a_hi + 8 * k, dlo + 2 * k, idesc, 1u);
          tc05::umma_tf32_ts(tmem_base, a_hi + 8 * k, dhi + 2 * k, idesc, 1u);
        }
        tc05::umma_commit(&a_empty[sa]);
        tc05::umma_commit(&b_empty[sb]);
      }
      tc05::umma_commit(accum_full);
    }
  }

  tc05::fence_before_thread_sync();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc05::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;   // idempotent lookup; benign if two threads race
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Row-major [rows, cols] fp32 matrix, box = 32 columns (128 B) x 128 rows, 128-byte swizzle.
int make_weight_map(CUtensorMap* map, const float* W, int rows, int cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { lnb::set_err("linear_tf32x3: cuTensorMapEncodeTiled unavailable"); return LNB_ERR_UNSUPPORTED; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(W), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lnb::set_err("linear_tf32x3: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return LNB_ERR_ARG;
  }
  return LNB_OK;
}

}  // namespace

extern "C" int lnb_linear_tf32x3(lnb_stream_t stream, const float* A, const float* W_hi,
                                 const float* W_lo, const float* bias, int M, int N, int K,
                                 int relu, float* C) {
  LNB_REQUIRE(A && W_hi && W_lo && C, "linear_tf32x3: null pointer");
  LNB_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear_tf32x3: bad dims M=%d N=%d K=%d", M, N, K);
  LNB_REQUIRE(K % 4 == 0, "linear_tf32x3: K=%d must be a multiple of 4 (16-byte rows)", K);
  LNB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W_hi & 15) == 0 && ((uintptr_t)W_lo & 15) == 0 &&
                  ((uintptr_t)C & 15) == 0,
              "linear_tf32x3: operands must be 16-byte aligned");
  if (M == 0) return LNB_OK;
  CUtensorMap map_hi, map_lo;
  int rc = make_weight_map(&map_hi, W_hi, N, K);
  if (rc != LNB_OK) return rc;
  rc = make_weight_map(&map_lo, W_lo, N, K);
  if (rc != LNB_OK) return rc;
  cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)SMEM_BYTES);   // per device; cheap
  dim3 grid(lnb::ceil_div(M, BM), lnb::ceil_div(N, BN));
  LNB_REQUIRE(grid.y <= 65535, "linear_tf32x3: N too large");
  linear_tf32x3_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(map_hi, map_lo, A, bias, C,
                                                                           M, N, K, relu);
  lnb::count_launch();
  return lnb::finish_launch("linear_tf32x3");
}
