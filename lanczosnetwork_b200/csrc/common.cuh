// Shared helpers for liblanczosnet_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/lanczosnet_b200.h"

namespace lnb {

// thread-local error text + launch counter (no other global mutable state)
char* err_buf();
void set_err(const char* fmt, ...);
void count_launch(int n = 1);
unsigned long long* prof_buffer();          // profiling aid (lnb_debug_set_prof), nullptr = off
void set_prof_buffer(unsigned long long* p);

void launch_tile_assign(cudaStream_t s, const int32_t* gext, int B, int K, int32_t* tiles,
                        int32_t* rowmap, int32_t* nrows);   // spectral_conv_fused.cu

inline int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_err("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return LNB_OK;
}

#define LNB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      lnb::set_err(__VA_ARGS__);          \
      return LNB_ERR_ARG;                 \
    }                                     \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace lnb
