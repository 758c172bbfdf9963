// micro-benchmark for the next kernel design: gate-GEMM MMA rate with the weight stream running, one CTA per SM
// (tcgen05 cta_group::1, M128 N96 K16, a 12 KB weight chunk per 4 MMAs per SM) against a CTA PAIR (cta_group::2,
// M256 N96 K16: every SM holds 48 of the 96 weight columns, 6 KB per chunk per SM).  A operand in tensor memory (TS
// form) as in rollout_tc.cu; the issuer does not wait for the stages (data is irrelevant for the timing), the
// producers are throttled by the per-chunk commits exactly as in the kernel.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../multimodaltraj_2_b200/csrc/tc_common.cuh"
using namespace mmt;

constexpr int NSTAGE = 4, STAGE = 12288;
constexpr int SM_BAR = NSTAGE * STAGE;   // W_FULL[4] W_EMPTY[4] PEER_FULL[4] DONE, tmem slot
constexpr int SM_TOTAL = SM_BAR + 160;

__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int PAIR>
__global__ void __launch_bounds__(128, 1) k(const uint8_t* w, int passes, int stream, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t W_FULL = sbase + SM_BAR, W_EMPTY = W_FULL + 8 * NSTAGE, PEER_FULL = W_EMPTY + 8 * NSTAGE,
                 DONE = PEER_FULL + 8 * NSTAGE, tslot = DONE + 16;
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
  constexpr int bytes = PAIR ? STAGE / 2 : STAGE;
  for (int i = tid; i < NSTAGE * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < 3 * NSTAGE + 1; ++s) mbar_init(W_FULL + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(tslot, 512);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * (3 * NSTAGE) + 16);   // tslot
  const uint32_t total = (uint32_t)passes * 5;   // chunks (4 MMAs each)

  if ((warp == 1 || warp == 2) && lane == 0 && stream) {
    // producers: chunk `it` -> stage it % 4, refilled as soon as the MMAs that read it have completed
    for (uint32_t it = warp - 1; it < total; it += 2) {
      const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
      if (stream == 2) mbar_wait_spin(W_EMPTY + 8 * s, ph ^ 1u); else mbar_wait(W_EMPTY + 8 * s, ph ^ 1u);
      mbar_arrive_expect_tx(W_FULL + 8 * s, bytes);
      bulk_g2s(sbase + s * STAGE, w + (size_t)(it % 20) * STAGE + rank * bytes, bytes, W_FULL + 8 * s);
    }
  } else if (PAIR && warp == 3 && lane == 0 && rank == 1 && stream) {
    // the peer tells the leader when its half of a stage has landed (the leader cannot wait on a remote mbarrier)
    for (uint32_t it = 0; it < total; ++it) {
      const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
      if (stream == 2) mbar_wait_spin(W_FULL + 8 * s, ph); else mbar_wait(W_FULL + 8 * s, ph);
      uint32_t ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(PEER_FULL + 8 * s), "r"(0));
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
    }
  } else if (warp == 0 && rank == 0) {
    // issuer: the whole warp runs convergently, one elected lane issues
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, 96);
    const long long t0 = clock64();
    uint32_t it = 0;
    for (int p = 0; p < passes; ++p) {
      const uint32_t d = tmem + 160 + (p & 1) * 96;
#pragma unroll
      for (int kc = 0; kc < 5; ++kc, ++it) {
        const uint32_t s = it % NSTAGE;
        if (stream) {
          if (stream == 2) {   // spinning waits (no suspend hint): lowest wake-up latency
            mbar_wait_spin(W_FULL + 8 * s, (it / NSTAGE) & 1u);
            if (PAIR) mbar_wait_spin(PEER_FULL + 8 * s, (it / NSTAGE) & 1u);
          } else {
            mbar_wait(W_FULL + 8 * s, (it / NSTAGE) & 1u);
            if (PAIR) mbar_wait(PEER_FULL + 8 * s, (it / NSTAGE) & 1u);
          }
          __syncwarp();
        }
        const uint64_t db = make_desc_sw128(sbase + s * STAGE);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = (kc | ks) ? 1u : 0u;
          if (PAIR) {
            asm volatile(
                "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
                "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d),
                "r"(tmem + kc * 32 + ks * 8), "l"(db + (uint64_t)(ks * 2)), "r"(idesc), "r"(acc), "r"(0)
                : "memory");
          } else {
            umma_bf16_ts_elect(d, tmem + kc * 32 + ks * 8, db + (uint64_t)(ks * 2), idesc, acc);
          }
        }
        if (PAIR) {
          asm volatile(
              "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\telect.sync _|e, 0xffffffff;\n\t"
              "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(
                  W_EMPTY + 8 * s)
              : "memory");
        } else {
          umma_commit_elect(W_EMPTY + 8 * s);
        }
      }
    }
    if (PAIR) {
      asm volatile(
          "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
          "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(DONE)
          : "memory");
    } else {
      umma_commit_elect(DONE);
    }
    mbar_wait(DONE, 0);
    __syncwarp();
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();
  if (warp == 3) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else tmem_dealloc(tmem, 512);
  }
}

template <int PAIR> void run(const uint8_t* w, long long* d, int stream) {
  auto kern = k<PAIR>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
  const int passes = 200;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = SM_TOTAL;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, w, passes, stream, d);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s (pair %d stream %d)\n", cudaGetErrorString(e), PAIR, stream); exit(1); }
  }
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s, weight stream %s: %.1f clk per MMA, %.0f clk per pass of 20 (nominal 48 / 960)\n",
         PAIR ? "CTA pair   cta_group::2 M256 N96 K16 (6 KB per chunk and SM) " : "single CTA cta_group::1 M128 N96 K16 (12 KB per chunk and SM)",
         stream == 2 ? "on, spinning waits" : stream ? "on " : "off", (double)h / (passes * 20), (double)h / passes);
}
int main(int argc, char** argv) {   // pair_bench <pair 0|1> <stream 0|1>: one configuration per process (a protocol bug must not take the others down)
  const int pair = argc > 1 ? atoi(argv[1]) : 0, stream = argc > 2 ? atoi(argv[2]) : 0;
  uint8_t* w; cudaMalloc(&w, 20 * STAGE); cudaMemset(w, 0, 20 * STAGE);
  long long* d; cudaMalloc(&d, 64);
  if (pair) run<1>(w, d, stream); else run<0>(w, d, stream);
  fflush(stdout);
  return 0;
}
