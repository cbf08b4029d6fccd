// Generic strided batched fp32 GEMM on CUDA cores (FFMA), arbitrary shapes and strides.
//   C[b,z] = act( alpha * (A[b,z] * kscale[b,z]) @ B[b,z] + beta * D[b,z] + bias )
// This is the general-shape path behind the torch.bmm call sites of the reference
// (model/lanczos_net.py:117,121,167,174,178; ada_lanczos_net.py:270,281,284,331,338,342):
// channel-innermost operator slices L[:, :, :, e] are consumed in place through their element
// stride (no strided-slice copies), Q^T through swapped strides.  The big dense layers go
// through the tcgen05 kernel in linear_tf32x3.cu instead.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, PADM = 4;

__global__ void __launch_bounds__(256)
batched_gemm_kernel(lnb_gemm_desc d, int a_k_contig, int b_n_contig) {
  __shared__ __align__(16) float As[TK][TM + PADM];
  __shared__ __align__(16) float Bs[TK][TN + PADM];

  const int bz = blockIdx.x;
  const int b = bz / d.nz, z = bz % d.nz;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.z * TN;
  const float* __restrict__ A = d.A + b * d.a_sb + z * d.a_sz;
  const float* __restrict__ Bm = d.B + b * d.b_sb + z * d.b_sz;
  const float* __restrict__ ks = d.kscale ? d.kscale + b * d.s_sb + z * d.s_sz : nullptr;
  float* __restrict__ C = d.C + b * d.c_sb + z * d.c_sz;
  const float* __restrict__ Dadd = d.addend ? d.addend + b * d.d_sb + z * d.d_sz : nullptr;
  const float alpha = d.alpha == 0.f ? 1.f : d.alpha;

  const int t = threadIdx.x;
  const int ty = t / 16, tx = t % 16;  // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + 256 * i;
      int m, k;
      if (a_k_contig) { k = e % TK; m = e / TK; } else { m = e % TM; k = e / TM; }
      float v = 0.f;
      if (m0 + m < d.M && k0 + k < d.K) {
        v = A[(int64_t)(m0 + m) * d.a_sm + (int64_t)(k0 + k) * d.a_sk];
        if (ks) v *= ks[(int64_t)(k0 + k) * d.s_sk];
      }
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + 256 * i;
      int n, k;
      if (b_n_contig) { n = e % TN; k = e / TN; } else { k = e % TK; n = e / TK; }
      float v = 0.f;
      if (n0 + n < d.N && k0 + k < d.K)
        v = Bm[(int64_t)(k0 + k) * d.b_sk + (int64_t)(n0 + n) * d.b_sn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= d.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= d.N) continue;
      float v = alpha * acc[i][j];
      if (Dadd) v = fmaf(d.beta, Dadd[(int64_t)m * d.d_sm + (int64_t)n * d.d_sn], v);
      if (d.bias) v += d.bias[(int64_t)z * d.bias_sz + n];
      if (d.relu) v = fmaxf(v, 0.f);
      C[(int64_t)m * d.c_sm + (int64_t)n * d.c_sn] = v;
    }
  }
}

}  // namespace

extern "C" int lnb_batched_gemm(lnb_stream_t stream, const lnb_gemm_desc* desc) {
  LNB_REQUIRE(desc, "batched_gemm: null descriptor");
  const lnb_gemm_desc& d = *desc;
  LNB_REQUIRE(d.A && d.B && d.C, "batched_gemm: null matrix pointer");
  LNB_REQUIRE(d.batch >= 0 && d.nz >= 1 && d.M >= 0 && d.N >= 0 && d.K >= 0,
              "batched_gemm: bad dims batch=%d nz=%d M=%d N=%d K=%d", d.batch, d.nz, d.M, d.N,
              d.K);
  if (d.batch == 0 || d.M == 0 || d.N == 0) return LNB_OK;
  int64_t gz = (int64_t)d.batch * d.nz;
  LNB_REQUIRE(gz <= 2147483647LL, "batched_gemm: batch*nz=%lld too large", (long long)gz);
  dim3 grid((unsigned)gz, lnb::ceil_div(d.M, TM), lnb::ceil_div(d.N, TN));
  LNB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "batched_gemm: M or N too large for the grid");
  batched_gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, d.a_sk == 1 ? 1 : 0,
                                                               d.b_sn == 1 ? 1 : 0);
  lnb::count_launch();
  return lnb::finish_launch("batched_gemm");
}
