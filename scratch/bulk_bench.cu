// micro-benchmark: latency of cp.async.bulk global(L2-resident)->shared with mbarrier completion, per size,
// 1 CTA vs 148 CTAs, ring depth 1 (pure latency) and depth 5 (throughput with the same ring as rollout_tc).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multimodaltraj_2_b200/csrc/tc_common.cuh"
using namespace mmt;
__global__ void __launch_bounds__(64, 1) k(const uint8_t* src, int bytes, int depth, int n, int same, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 5 * 12288;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 5; ++s) mbar_init(bar + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (same ? 0 : (size_t)blockIdx.x * 20 * 12288);
    // warm: touch
    long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t ph[5] = {0, 0, 0, 0, 0};
    while (done < n) {
      while (issued < n && issued - done < depth) {
        const int s = issued % 5;
        mbar_arrive_expect_tx(bar + 8 * s, bytes);
        bulk_g2s(sbase + s * 12288, base + (size_t)(issued % 20) * 12288, bytes, bar + 8 * s);
        ++issued;
      }
      const int s = done % 5;
      mbar_wait(bar + 8 * s, ph[s]);
      ph[s] ^= 1;
      ++done;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}
int main() {
  uint8_t* src; cudaMalloc(&src, 148 * 20 * 12288); cudaMemset(src, 1, 148 * 20 * 12288);
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n = 400;
  for (int grid : {1, 148})
    for (int same : {1, 0})
      for (int depth : {1, 5})
        for (int bytes : {768, 3072, 12288}) {
          k<<<grid, 64, 64 * 1024>>>(src, bytes, depth, n, same, d);
          k<<<grid, 64, 64 * 1024>>>(src, bytes, depth, n, same, d);
          cudaDeviceSynchronize();
          long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          printf("grid %3d %s depth %d bytes %5d: %.0f clk per copy (%.1f B/clk/SM)\n", grid, same ? "same-addr" : "own-addr ", depth, bytes, (double)h / n, (double)bytes * n / h);
        }
  return 0;
}
