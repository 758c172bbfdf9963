// micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM for nw worker warps (32x32b shapes x4 / x8 / x16 / x32)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multimodaltraj_2_b200/csrc/tc_common.cuh"
using namespace mmt;
template <int X> __device__ __forceinline__ void ld(uint32_t addr, uint32_t* r);
template <> __device__ __forceinline__ void ld<4>(uint32_t addr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
template <> __device__ __forceinline__ void ld<8>(uint32_t addr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory");
}
template <> __device__ __forceinline__ void ld<16>(uint32_t addr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr) : "memory");
}
template <int X, int ST>
__global__ void __launch_bounds__(544, 1) k(int nw, int iters, long long* out) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 16) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp < nw) {
    const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t acc = 0;
    uint32_t r[16] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ST) {
          if (X == 8) tmem_st8(t_row + j * 8, r);
          else tmem_st4(t_row + j * 4, r);
        } else {
          ld<X>(t_row + (j * X) % 64, r);
        }
      }
      if (ST) tmem_wait_st(); else tmem_wait_ld();
      acc += r[0];
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    if (acc == 0x1234567) out[1] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}
template <int X, int ST> void run(long long* d) {
  const int iters = 2000;
  for (int nw : {1, 4, 8, 16}) {
    k<X, ST><<<148, 544>>>(nw, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)nw * iters * 4 * X * 32 * 4;
    printf("%s x%-2d warps %2d: %.1f clk per instruction per warp, %.1f B/clk/SM\n", ST ? "st" : "ld", X, nw, (double)h / (iters * 4), bytes / h);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<4, 0>(d); run<8, 0>(d); run<16, 0>(d); run<8, 1>(d); run<4, 1>(d);
  return 0;
}
