// unsorted_segment_sum forward / backward.
// Replaces operators/src/cuda/segment_reduction.cu:39-95 (reference).  Semantics are the
// *intended* ones: output is [B, num_segments, dim2] with batch stride num_segments*dim2
// (the reference kernel hard-codes dim1*dim2, segment_reduction.cu:48, which only agrees
// when num_segments == dim1).  Out-of-range ids are ignored instead of writing out of bounds.
#include "common.cuh"

namespace {

// One thread per 4 consecutive features when dim2 % 4 == 0 (vector path), else scalar.
template <int VEC>
__global__ void segsum_fwd_kernel(const float* __restrict__ data,
                                  const int64_t* __restrict__ ids, int64_t total, int dim1,
                                  int dim2v, int num_segments, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % dim2v);
    int64_t row = i / dim2v;  // b*dim1 + c
    int64_t b = row / dim1;
    int64_t seg = ids[row];
    if (seg < 0 || seg >= num_segments) continue;
    float* dst = out + ((b * num_segments + seg) * dim2v + x) * VEC;
    if (VEC == 4) {
      // one 16-byte reduction (red.global.add.v4.f32, sm_90+) instead of four scalar atomics;
      // dst is 16-byte aligned on this path (dim2 % 4 == 0 and an aligned output base)
      const float4 v = reinterpret_cast<const float4*>(data)[i];
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w) : "memory");
    } else {
      atomicAdd(dst, data[i]);
    }
  }
}

template <int VEC>
__global__ void segsum_bwd_kernel(const float* __restrict__ gout,
                                  const int64_t* __restrict__ ids, int64_t total, int dim1,
                                  int dim2v, int num_segments, float* __restrict__ gdata) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % dim2v);
    int64_t row = i / dim2v;
    int64_t b = row / dim1;
    int64_t seg = ids[row];
    bool ok = seg >= 0 && seg < num_segments;
    if (VEC == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v = reinterpret_cast<const float4*>(gout)[(b * num_segments + seg) * dim2v + x];
      reinterpret_cast<float4*>(gdata)[i] = v;
    } else {
      gdata[i] = ok ? gout[(b * num_segments + seg) * dim2v + x] : 0.f;
    }
  }
}

int grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = 148 * 16;  // 148 SMs x 16 resident 256-thread blocks, grid-stride beyond
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

int lnb_unsorted_segment_sum_forward(lnb_stream_t stream, const float* data,
                                     const int64_t* segment_ids, const int* data_shape,
                                     int num_segments, float* output) {
  LNB_REQUIRE(data_shape, "segment_sum_forward: null shape");
  int B = data_shape[0], d1 = data_shape[1], d2 = data_shape[2];
  if ((int64_t)B * d1 * d2 == 0) return LNB_OK;   // empty input: nothing to add
  LNB_REQUIRE(data && segment_ids && output, "segment_sum_forward: null pointer");
  LNB_REQUIRE(B >= 0 && d1 >= 0 && d2 >= 0 && num_segments > 0,
              "segment_sum_forward: bad shape [%d,%d,%d] S=%d", B, d1, d2, num_segments);
  int64_t n = (int64_t)B * d1 * d2;
  if (n == 0) return LNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (d2 % 4 == 0 && aligned16(data) && aligned16(output)) {
    int64_t t = n / 4;
    segsum_fwd_kernel<4><<<grid_for(t), 256, 0, s>>>(data, segment_ids, t, d1, d2 / 4,
                                                      num_segments, output);
  } else {
    segsum_fwd_kernel<1><<<grid_for(n), 256, 0, s>>>(data, segment_ids, n, d1, d2, num_segments,
                                                      output);
  }
  lnb::count_launch();
  return lnb::finish_launch("segment_sum_forward");
}

int lnb_unsorted_segment_sum_backward(lnb_stream_t stream, const float* grad_output,
                                      const int64_t* segment_ids, const int* data_shape,
                                      int num_segments, float* grad_data) {
  LNB_REQUIRE(data_shape, "segment_sum_backward: null shape");
  int B = data_shape[0], d1 = data_shape[1], d2 = data_shape[2];
  if ((int64_t)B * d1 * d2 == 0) return LNB_OK;
  LNB_REQUIRE(grad_output && segment_ids && grad_data, "segment_sum_backward: null pointer");
  LNB_REQUIRE(B >= 0 && d1 >= 0 && d2 >= 0 && num_segments > 0,
              "segment_sum_backward: bad shape [%d,%d,%d] S=%d", B, d1, d2, num_segments);
  int64_t n = (int64_t)B * d1 * d2;
  if (n == 0) return LNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (d2 % 4 == 0 && aligned16(grad_output) && aligned16(grad_data)) {
    int64_t t = n / 4;
    segsum_bwd_kernel<4><<<grid_for(t), 256, 0, s>>>(grad_output, segment_ids, t, d1, d2 / 4,
                                                      num_segments, grad_data);
  } else {
    segsum_bwd_kernel<1><<<grid_for(n), 256, 0, s>>>(grad_output, segment_ids, n, d1, d2,
                                                      num_segments, grad_data);
  }
  lnb::count_launch();
  return lnb::finish_launch("segment_sum_backward");
}

}  // extern "C"
