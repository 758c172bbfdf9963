// PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, bulk copy, TMEM, UMMA descriptors.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace mmt {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of returning after a few cycles -- a bare try_wait loop spins ~every 26 clk and steals issue
// slots from the warps that share its scheduler (ncu: 10 % of all samples in the producers' loop).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// A trap names itself first: the trap record is a 16-word block of host-mapped pinned memory (capi.cu:
// trap_record(); reported by mmt_last_trap and appended to the error text of the next failing launch check), written
// with a system-scope fence before __trap() kills the context.  word 0 = site code (kernel << 8 | wait), 1 = CTA,
// 2 = thread, 3 = barrier shared address, 4 = awaited parity, 5 = extra.
// rec == nullptr (a compile-time constant in rollout_tc_kernel's production instantiation) reduces this to a bare trap:
// measured on a B200 (profiles/r02_trap_styles.txt), carrying the record pointer into that kernel's wait loops costs
// 5 % (2.058 vs 1.957 ms; a pointer in a __device__ global 2-8 %: the kernel sits at its 96-register cap), so there only
// the diagnostics instantiation (MMT_RO_FLAGS != 0) names its site.
static __device__ __forceinline__ void trap_report(uint32_t* rec, uint32_t site, uint32_t bar, uint32_t parity, uint32_t extra = 0) {
  if (rec) {
    volatile uint32_t* r = rec;
    r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = bar; r[4] = parity; r[5] = extra;
    r[0] = site;
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void require_smem_alignment(const void* smem, uint32_t* rec, uint32_t kernel_id) {
  if ((smem_u32(smem) & 1023u) != 0u) trap_report(rec, (kernel_id << 8) | 0xFFu, smem_u32(smem), 0);
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.  One iteration of the hinted
// try_wait measured ~40 clk while the barrier is pending (ncu: 15 k iterations per warp in 0.33 ms of waiting), so
// 2^26 iterations are >= 1.4 s -- three orders of magnitude above any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* rec = nullptr, uint32_t site = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) trap_report(rec, site, bar, parity);
  }
}
// spinning wait (no suspend hint): lowest wake-up latency, costs issue slots -- for single-thread roles on a latency-critical ring
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity, uint32_t* rec = nullptr, uint32_t site = 0) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (++spins > (1u << 28)) trap_report(rec, site, bar, parity);
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// A operand in tensor memory (TS form): row i of A in TMEM lane i, two consecutive-k bf16 per 32-bit column
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent issue: the WHOLE warp executes these with warp-uniform operands and one elected lane issues.
// ptxas then keeps the operands in uniform registers; issuing from a divergent `if (lane == 0)` region costs an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop per MMA (57-68 vs 49 clk per M128 x N96 MMA: scratch/mma_bench4.cu).
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, const uint32_t r[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float v[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 lanes x 256 bit: thread t of the warp holds (lane t/4, columns 2(t%4), 2(t%4)+1) in v[0..1] and (lane t/4 + 8, same
// columns) in v[2..3] (cute SM100_TMEM_LOAD_16dp256b1x); the lane field of the address selects lanes 0-15 or 16-31 of
// the warp's quarter.  One thread then owns 4 rows x 2 columns instead of 1 row x 8 columns: per-column constants
// cost a quarter of the shared-memory load instructions.
__device__ __forceinline__ void tmem_ld_16x256b(uint32_t addr, float v[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_16x256b(uint32_t addr, const uint32_t r[4]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row, k) inside a [rows x 64] bf16 K-major SWIZZLE_128B block
__host__ __device__ __forceinline__ uint32_t sw128_off(int row, int k) {
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)k & 7u) << 1);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&b);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// fp16 operands (MMT_PREC_F16: the same kernels with 10 instead of 7 stored mantissa bits; kind::f16 takes either format)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float f16_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
__device__ __forceinline__ float f16_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
// instruction descriptor with A = B = f16 (format code 0), D = f32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool F16>
__host__ __device__ constexpr uint32_t make_idesc_op(int M, int N) {
  return F16 ? make_idesc_f16(M, N) : make_idesc_bf16(M, N);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_op2(float lo, float hi) {
  if constexpr (F16) return pack_f16x2(lo, hi); else return pack_bf16x2(lo, hi);
}
template <bool F16>
__device__ __forceinline__ float op_lo(uint32_t w) {
  if constexpr (F16) return f16_lo(w); else return bf16_lo(w);
}
template <bool F16>
__device__ __forceinline__ float op_hi(uint32_t w) {
  if constexpr (F16) return f16_hi(w); else return bf16_hi(w);
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 issue two fp32 lanes per instruction)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)), "l"(*reinterpret_cast<uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 tanh2(float2 a) { return make_float2(tanh_fast(a.x), tanh_fast(a.y)); }

}  // namespace mmt
