// micro-benchmark 4: MMA issue cost of (a) one divergent thread (if lane == 0) vs (b) the whole warp convergent with
// elect.sync inside the asm and warp-uniform operands.  TS form, N = 96, commit per chunk.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multimodaltraj_2_b200/csrc/tc_common.cuh"
using namespace mmt;
__device__ __forceinline__ void umma_ts_elect(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(384, 1) k(int passes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = sbase + 160 * 1024, tslot = bar + 64;
  for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { for (int s = 0; s < 8; ++s) mbar_init(bar + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 8) tmem_alloc(tslot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 160 * 1024 + 64);
  constexpr uint32_t idesc = make_idesc_bf16(128, 96);
  if (warp == 9) {
    if (MODE == 0) {
      if ((tid & 31) == 0) {
        const long long t0 = clock64();
        for (int p = 0; p < passes; ++p) {
#pragma unroll
          for (int kc = 0; kc < 5; ++kc) {
            const uint64_t db = make_desc_sw128(sbase + 80 * 1024 + kc * 12288);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16_ts(tmem, tmem + 320 + kc * 32 + ks * 8, db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
            umma_commit(bar + 8 * 4);
          }
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        if (blockIdx.x == 0) out[0] = clock64() - t0;
      }
    } else {
      tmem = __shfl_sync(0xffffffffu, tmem, 0);
      const long long t0 = clock64();
      for (int p = 0; p < passes; ++p) {
#pragma unroll
        for (int kc = 0; kc < 5; ++kc) {
          const uint64_t db = make_desc_sw128(sbase + 80 * 1024 + kc * 12288);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts_elect(tmem, tmem + 320 + kc * 32 + ks * 8, db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
          commit_elect(bar + 8 * 4);
        }
      }
      commit_elect(bar);
      mbar_wait(bar, 0);
      if (blockIdx.x == 0 && (tid & 31) == 0) out[0] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}
template <int MODE> void run(long long* d) {
  auto kern = k<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 162 * 1024);
  const int passes = 20;
  kern<<<148, 384, 162 * 1024>>>(passes, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("mode %d (%s): %.1f clk per MMA\n", MODE, MODE ? "convergent warp + elect.sync" : "single divergent thread", (double)h / (passes * 20));
}
int main() {
  long long* d; cudaMalloc(&d, 1024);
  run<0>(d); run<1>(d);
  return 0;
}
