// Shared helpers for libmmt (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmt.h"

namespace mmt {

// SM count of the CURRENT device (queried once per device; B200: 148 = 2 dies x 74 SMs)
int num_sms();

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Opt a kernel into `bytes` of dynamic shared memory on the current device.  The attribute is per device and per
// function: `done` is the caller's bit mask of devices already opted in (one static mask per kernel).  Thread-safe;
// returns MMT_OK / MMT_ECUDA.
using DeviceMask = std::atomic<unsigned long long>;
int opt_in_smem(const void* func, int bytes, DeviceMask* done);

// 16 words of host-mapped pinned memory (same address on every device under UVA) that a trapping kernel fills in
// before it dies (tc_common.cuh: trap_report); nullptr if the allocation failed.  Never freed (process lifetime).
uint32_t* trap_record();

// value of an integer environment variable read ONCE per process (diagnostic switches; never on the launch path)
int env_int_once(const char* name, std::atomic<int>* cache);   // *cache: INT_MIN until read

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
