// Shared helpers for libmmt (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmt.h"

namespace mmt {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// check the launch that was just issued; returns MMT_OK / MMT_ECUDA
int check_launch(const char* what);

#define MMT_REQUIRE(cond, msg)                       \
  do {                                               \
    if (!(cond)) {                                   \
      mmt::set_error("%s: %s", __func__, msg);       \
      return MMT_EARG;                               \
    }                                                \
  } while (0)

#define MMT_ALIGNED(p)                                                  \
  do {                                                                  \
    if ((p) != nullptr && !mmt::aligned16(p)) {                         \
      mmt::set_error("%s: pointer %s not 16-byte aligned", __func__, #p); \
      return MMT_EALIGN;                                                \
    }                                                                   \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// streaming (evict-first) 128-bit store: outputs written once and not re-read by this kernel
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

}  // namespace mmt
