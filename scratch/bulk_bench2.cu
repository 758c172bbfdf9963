// micro-benchmark 2: where does the fixed ~360 clk per cp.async.bulk go?  (a) time of the issuing instructions alone,
// (b) one thread per stage (5 lanes of one warp / 5 warps) instead of one thread for the whole ring.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multimodaltraj_2_b200/csrc/tc_common.cuh"
using namespace mmt;
// mode 0: single thread, ring of 5; mode 1: lanes 0..4 of warp 0, one stage each; mode 2: warps 0..4 lane 0, one stage each
__global__ void __launch_bounds__(160, 1) k(const uint8_t* src, int bytes, int mode, int n, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 5 * 12288;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 5; ++s) mbar_init(bar + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  long long t_issue = 0, t_wait = 0;
  if (mode == 0) {
    if (threadIdx.x == 0) {
      int issued = 0, done = 0;
      uint32_t ph[5] = {0, 0, 0, 0, 0};
      while (done < n) {
        while (issued < n && issued - done < 5) {
          const int s = issued % 5;
          long long a = clock64();
          mbar_arrive_expect_tx(bar + 8 * s, bytes);
          bulk_g2s(sbase + s * 12288, src + (size_t)(issued % 20) * 12288, bytes, bar + 8 * s);
          t_issue += clock64() - a;
          ++issued;
        }
        const int s = done % 5;
        long long a = clock64();
        mbar_wait(bar + 8 * s, ph[s]);
        t_wait += clock64() - a;
        ph[s] ^= 1;
        ++done;
      }
    }
  } else {
    const bool mine = mode == 1 ? (warp == 0 && lane < 5) : (lane == 0 && warp < 5);
    const int s = mode == 1 ? lane : warp;
    if (mine) {
      uint32_t ph = 0;
      for (int i = s; i < n; i += 5) {
        long long a = clock64();
        mbar_arrive_expect_tx(bar + 8 * s, bytes);
        bulk_g2s(sbase + s * 12288, src + (size_t)(i % 20) * 12288, bytes, bar + 8 * s);
        long long b = clock64();
        mbar_wait(bar + 8 * s, ph);
        t_issue += b - a;
        t_wait += clock64() - b;
        ph ^= 1;
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; out[2] = t_wait; }
}
int main() {
  uint8_t* src; cudaMalloc(&src, 20 * 12288); cudaMemset(src, 1, 20 * 12288);
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n = 400;
  for (int mode : {0, 1, 2})
    for (int bytes : {768, 12288}) {
      k<<<148, 160, 64 * 1024>>>(src, bytes, mode, n, d);
      k<<<148, 160, 64 * 1024>>>(src, bytes, mode, n, d);
      cudaDeviceSynchronize();
      long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
      printf("mode %d bytes %5d: %.0f clk per copy (%.1f B/clk/SM); thread0: issue %.0f clk/copy-issued, wait %.0f clk/wait\n", mode, bytes, (double)h[0] / n,
             (double)bytes * n / h[0], (double)h[1] / (mode ? n / 5 : n), (double)h[2] / (mode ? n / 5 : n));
    }
  return 0;
}
