// The whole obs T -> pred P recurrence of the bf16 path as ONE persistent kernel (sm_100a):
// pairwise kernel + adjacency + masked softmax -> graph aggregation (tcgen05) -> gsk_lstm_cell gate GEMM
// (tcgen05) -> gate update -> head -> next position, for all T+P-1 steps, with the recurrent state
// ON CHIP.  ("tensor pipe only if the 20-step recurrence is fused with on-chip state".)
//
// One CTA per SM owns a 128-row tile (128/N whole scenes) for all its steps:
//   tensor memory  A operand of the gate GEMM [e | h | mh] as bf16 pairs (160 columns; the MMAs take A
//                  from TMEM, so only the weights cross shared memory), 2 gate accumulators (96 columns:
//                  i|j|o of 32 units), mc accumulator (128), mh accumulator (128, aliases accumulator 1)
//   shared memory  H, C   bf16 [agent][unit] SWIZZLE_128B images = MN-major B operand (agents along K) of
//                         the aggregation GEMMs  mh = att x h,  mc = att x c
//                  ATT    block-diagonal un-normalised attention, K-major A operand of the aggregation
//                  W ring gate weights streamed from L2 by cp.async.bulk, 8 stages of 12 KB (one k-chunk)
//                  CF     fp32 cell state c, [unit/4][row][4]: conflict-free 128-bit access by (row, unit quad);
//                         keeping it out of the register file is what lets 16 worker warps fit (4 per scheduler:
//                         with 8 the workers issued one instruction per ~7 clk, latency-bound: profiles/)
// Only positions / vislets are read from HBM and only the 5 head parameters per predicted step are
// written: ~0.5 KB per agent-trajectory instead of ~30 KB per agent for the per-step kernels.
//
// Warp roles: warps 0-15 workers (attention build, e / mh operand build, gate epilogue, head): warp w owns TMEM
// lane quarter w % 4 (rows) and column slice w / 4 (8 of the 32 units of a pass); warps 16-17 stream weight
// stages; warps 18-19 issue the tcgen05.mma (gate passes alternate between them).
// Measured design inputs (scratch/mma_bench.cu, bulk_bench2.cu, mufu_bench2.cu on B200): an SS-form
// M128 x N96 MMA re-reads 4 KB of A from shared memory per 48-clk MMA and, together with the workers'
// own shared-memory traffic, ran at 83-90 clk in situ; one thread managing a bulk-copy ring sustains one
// copy per ~360 clk; MUFU.TANH / EX2 issue 16 lanes/clk/SM.
#include <cuda_bf16.h>
#include <limits.h>

#include <type_traits>

#include "mmt_common.cuh"
#include "tc_common.cuh"

#ifndef RO_AGG256
#define RO_AGG256 1   // 0: two N = 128 aggregation MMAs per k-step (mh, then mc) as in the first fused version
#endif

namespace mmt {

constexpr int RO_U = 128;
constexpr int RO_UN = 32;                       // units per gate pass
constexpr int RO_NP = RO_U / RO_UN;             // 4 passes
constexpr int RO_N = 3 * RO_UN;                 // 96 accumulator columns per pass
constexpr int RO_NKC = 5;                       // k-chunks of 64: e | h0 h1 | mh0 mh1
constexpr int RO_NCH = RO_NP * RO_NKC;          // 20 weight chunks per step
constexpr int RO_BLK = 128 * 128;               // one [128 rows x 128 B] block
constexpr int RO_CHUNK_BYTES = RO_N * 64 * 2;   // 12288: one (pass, k-chunk) of the packed weights
constexpr int RO_STAGE_BYTES = RO_CHUNK_BYTES;
constexpr int RO_NSTAGE = 4;
constexpr int RO_WORKERS = 512;
constexpr int RO_WWARPS = RO_WORKERS / 32;      // 16: first helper warp
constexpr int RO_NPROD = 2;                     // weight-stage producer warps (one issuing thread each)
constexpr int RO_NISSUE = 2;                    // MMA-issuing warps: passes p % 2 == j (one thread tops out at ~68 clk/MMA)
constexpr int RO_THREADS = RO_WORKERS + 32 * (RO_NPROD + RO_NISSUE);

constexpr int RS_H = 0;                         // 2 blocks (units 0-63 | 64-127)
constexpr int RS_C = RS_H + 2 * RO_BLK;         // 2 blocks
constexpr int RS_ATT = RS_C + 2 * RO_BLK;       // 2 blocks (agents 0-63 | 64-127 along K)
constexpr int RS_CF = RS_ATT + 2 * RO_BLK;      // fp32 c: 64 KB
constexpr int RS_W = RS_CF + 128 * 128 * 4;
constexpr int RS_BAR = RS_W + RO_NSTAGE * RO_STAGE_BYTES;
constexpr int RS_TMEM = RS_BAR + 256;
constexpr int RS_BIAS = RS_TMEM + 16;                  // b[384], w_If, w_It, w_Of, w_Ot [4][128]
constexpr int RS_WE = RS_BIAS + (384 + 512) * 4;       // W_e[4][64], b_e[64]
constexpr int RS_WHT = RS_WE + (256 + 64) * 4;         // head weights transposed: W_hT[5][2U] (unit pairs feed FFMA2)
constexpr int RS_PX = RS_WHT + 5 * 256 * 4;            // float[128] current x (invalid agents: far away)
constexpr int RS_PY = RS_PX + 512;                     // float[128] current y
constexpr int RS_NEXT = RS_PY + 512;                   // float2[128] predicted next positions
constexpr int RS_SUM = RS_NEXT + 1024;                 // float[4][128] attention row sums
constexpr int RS_OBS = RS_SUM + 2048;                  // float4[128]: next observed frame (pos.xy, vislet.xy), cp.async
constexpr int RS_TOTAL = RS_OBS + 2048;
static_assert(RS_TOTAL + 1024 <= 227 * 1024, "shared memory budget");

// tensor-memory columns
constexpr uint32_t RT_A = 0;          // A operand: e (32 columns) | h (64) | mh (64), two bf16 per column
constexpr uint32_t RT_A_H = 32, RT_A_MH = 96;
constexpr uint32_t RT_ACC0 = 160, RT_ACC1 = 256;
constexpr uint32_t RT_MH = 256;       // mh accumulator (fp32, 128 columns): aliases accumulator 1 + 32 spare columns
constexpr uint32_t RT_MC = 384;       // mc accumulator (fp32, 128 columns)
constexpr uint32_t RT_HEAD = 352;     // 4 column slices x 8: head partial sums of a row, exchanged through TMEM
                                      // (the 32 columns of the mh accumulator beyond accumulator 1: free after the conversion)
// F16 = the operand format of every MMA of the kernel: bf16 (MMT_PREC_BF16) or fp16 (MMT_PREC_F16: three more mantissa
// bits in the A operand, the state images, the attention numerators and the packed weights; same instructions, same speed)
template <bool F16> constexpr uint32_t kIdescGate = make_idesc_op<F16>(128, RO_N);
template <bool F16> constexpr uint32_t kIdescAggMN = make_idesc_op<F16>(128, 128) | (1u << 16);   // B operand MN-major
template <bool F16> constexpr uint32_t kIdescAgg256 = make_idesc_op<F16>(128, 256) | (1u << 16);  // [h | c] in one MMA

struct RoArgs {
  const float* pos;      // [R, F, 2]
  const float* vis;      // [R, T, 2]
  const uint8_t* valid;  // [R]
  const float *W_e, *b_e, *b, *w_If, *w_It, *w_Of, *w_Ot, *W_h, *b_h;
  const uint8_t* Wp;     // packed bf16 gate weights (mmt_pack_gate_weights_bf16)
  float* params;         // [R, P, 5]
  int R, N, T, P, F, num_tiles;
  float r2, neg_inv_log2e;
  int flags;  // diagnostics: 1 = no weight streaming (timing experiments only), 16 = time the W_FULL waits
  long long* dbg;        // optional [64 steps][32] clock64 stamps of CTA 0: [0,16) worker thread 0, [16,32) MMA thread
  uint32_t* trap;        // host-mapped trap record (tc_common.cuh: trap_report); may be null
};
// trap sites of this kernel (kernel id 1): which bounded wait expired
enum : uint32_t {
  RT_W_EMPTY = 0x101, RT_W_FULL = 0x102, RT_ACC_EMPTY = 0x103, RT_ACC_EMPTY_P0 = 0x104, RT_ACC1_EMPTY_P0 = 0x105,
  RT_ATT_READY = 0x106, RT_E_READY = 0x107, RT_MH_READY_I0 = 0x108, RT_MH_READY_I1 = 0x109, RT_P0_ISSUED = 0x10A,
  RT_P1_ISSUED = 0x10B, RT_P2_ISSUED = 0x10C, RT_AGG_FULL = 0x10D, RT_ACC_FULL = 0x10E
};

// MN-major SWIZZLE_128B operand: 64 MN-elements (128 B) contiguous, 8 k-rows of 128 B per atom;
// LBO = byte distance between 64-element MN chunks, SBO = byte distance between 8-row k groups
// (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units).
__device__ __forceinline__ uint64_t make_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// k-chunk of the packed weights / A operand (0 = e, 1-2 = h, 3-4 = mh) consumed at position kcn of a pass
__device__ __forceinline__ constexpr int ro_perm(int kcn) { return kcn == 0 ? 1 : kcn == 1 ? 2 : kcn == 2 ? 0 : kcn; }

// tanh on [-1.05, 1.05] on the FMA pipe: x P3(x^2), minimax, max abs error 1.1e-4 (tanh.approx.f32: ~5e-4).  The new cell
// state is a convex combination of the old one (|c| <= 1 from the zero start) and tanh(j), so |c'| <= 1 up to rounding: the
// one tanh per unit whose argument is bounded leaves the MUFU pipe (4 -> 3 per unit on observed steps, 5 -> 4 on emitting ones).
__device__ __forceinline__ float2 tanh_unit2(float2 x) {
  const float2 u = fmul2(x, x);
  float2 p = ffma2(make_float2(-0.0254361462f, -0.0254361462f), u, make_float2(0.117456769f, 0.117456769f));
  p = ffma2(p, u, make_float2(-0.330260619f, -0.330260619f));
  p = ffma2(p, u, make_float2(0.999905849f, 0.999905849f));
  return fmul2(x, p);
}
__device__ __forceinline__ float2 half_tanh_unit2(float2 x) {   // tanh(x) / 2: the 1/2 folded into the coefficients
  const float2 u = fmul2(x, x);
  float2 p = ffma2(make_float2(-0.0127180731f, -0.0127180731f), u, make_float2(0.0587283845f, 0.0587283845f));
  p = ffma2(p, u, make_float2(-0.1651303095f, -0.1651303095f));
  p = ffma2(p, u, make_float2(0.4999529245f, 0.4999529245f));
  return fmul2(x, p);
}

// wait executed by a whole (convergent) warp: reconverge before the next elect.sync
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, uint32_t* trap, uint32_t site) {
  mbar_wait(bar, parity, trap, site);
  __syncwarp();
}

// One mbarrier arrival per warp: 512 per-thread arrivals on one barrier word serialise in the shared-memory
// atomic unit (the issuer saw a barrier complete ~1000 clk after the typical worker had arrived).  Every lane has
// issued its own fences; __syncwarp orders the lanes' writes before lane 0's releasing arrive.
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// DIAG = true compiles the timeline stamps and the timing-experiment flags in (scratch/ro_timeline.py); the production
// instantiation carries none of it: a few extra instructions per chunk in the MMA-issuing warps cost 2.5 % of the kernel.
template <bool DIAG, bool F16>
__global__ void __launch_bounds__(RO_THREADS, 1) rollout_tc_kernel(RoArgs a) {
  // 1024-byte alignment (SWIZZLE_128B atoms) requested from the toolchain instead of fixed up at run time: the base is
  // then a link-time constant and every barrier address / UMMA descriptor derived from it is uniform
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* const smem = smem_dyn;
  uint32_t* const trap = DIAG ? a.trap : nullptr;   // production: bare traps (see tc_common.cuh: trap_report)
  require_smem_alignment(smem, trap, 1);
  if (DIAG && (a.flags & 1024) && blockIdx.x == 1 && threadIdx.x == 33) trap_report(trap, 0x1EE, 0xABCD, 1);   // self-test of the trap record
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + RS_BAR;
  const uint32_t W_FULL = bar0, W_EMPTY = bar0 + 8 * RO_NSTAGE, ACC_FULL = bar0 + 16 * RO_NSTAGE,
                 ACC_EMPTY = ACC_FULL + 16, ATT_READY = ACC_EMPTY + 16, E_READY = ATT_READY + 8,
                 MH_READY = E_READY + 8, AGG_FULL = MH_READY + 8, P0_ISSUED = AGG_FULL + 8, P1_ISSUED = P0_ISSUED + 8,
                 P2_ISSUED = P1_ISSUED + 8;
  float* s_bias = reinterpret_cast<float*>(smem + RS_BIAS);
  float* s_we = reinterpret_cast<float*>(smem + RS_WE);
  float* s_wht = reinterpret_cast<float*>(smem + RS_WHT);
  float* s_px = reinterpret_cast<float*>(smem + RS_PX);
  float* s_py = reinterpret_cast<float*>(smem + RS_PY);
  float2* s_next = reinterpret_cast<float2*>(smem + RS_NEXT);
  float* s_sum = reinterpret_cast<float*>(smem + RS_SUM);
  float4* s_obs = reinterpret_cast<float4*>(smem + RS_OBS);
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + RS_TMEM);
  const int nsteps = a.T + a.P - 1;

  if (tid == 0) {
    for (int s = 0; s < RO_NSTAGE; ++s) {
      mbar_init(W_FULL + 8 * s, 1);
      mbar_init(W_EMPTY + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(ACC_FULL + 8 * b, 1);
      mbar_init(ACC_EMPTY + 8 * b, RO_WWARPS);
    }
    mbar_init(ATT_READY, RO_WWARPS);
    mbar_init(E_READY, RO_WWARPS);
    mbar_init(MH_READY, RO_WWARPS);
    mbar_init(AGG_FULL, 1);
    mbar_init(P0_ISSUED, 1);
    mbar_init(P1_ISSUED, 1);
    mbar_init(P2_ISSUED, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == RO_WWARPS) tmem_alloc(sbase + RS_TMEM, 512);
  // sigmoid(z) = 0.5 tanh(z/2) + 0.5: the 1/2 is folded into the packed i/o weight columns, biases and peepholes
  for (int i = tid; i < 384; i += RO_THREADS) s_bias[i] = (i >= 128 && i < 256) ? a.b[i] : 0.5f * a.b[i];
  for (int i = tid; i < 128; i += RO_THREADS) {
    s_bias[384 + i] = 0.5f * a.w_If[i];
    s_bias[512 + i] = 0.5f * a.w_It[i];
    s_bias[640 + i] = 0.5f * a.w_Of[i];
    s_bias[768 + i] = 0.5f * a.w_Ot[i];
  }
  for (int i = tid; i < 256; i += RO_THREADS) s_we[i] = a.W_e[i];
  for (int i = tid; i < 64; i += RO_THREADS) s_we[256 + i] = a.b_e[i];
  for (int i = tid; i < 256 * 5; i += RO_THREADS) s_wht[(i % 5) * 256 + i / 5] = a.W_h[i];
  // attention operand: entries outside a row's own scene stay zero for the whole kernel
  for (int i = tid; i < 2 * RO_BLK / 16; i += RO_THREADS)
    reinterpret_cast<uint4*>(smem + RS_ATT)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp >= RO_WWARPS && warp < RO_WWARPS + RO_NPROD) {
    // =============================== weight-stage producers ===============================
    // One thread managing a whole ring sustains only one cp.async.bulk per ~360 clk (issue + mbarrier round
    // trip serialise in that thread: scratch/bulk_bench2.cu); RO_NPROD threads in different warps take the
    // stages round-robin instead.
    if (lane == 0) {
      int my_tiles = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) ++my_tiles;
      const uint32_t total = (DIAG && (a.flags & 1)) ? 0u : (uint32_t)my_tiles * nsteps * RO_NCH;
      for (uint32_t it = warp - RO_WWARPS; it < total; it += RO_NPROD) {
        const uint32_t s = it % RO_NSTAGE, ph = (it / RO_NSTAGE) & 1u, j = it % RO_NCH;
        if (DIAG && (a.flags & 64)) mbar_wait_spin(W_EMPTY + 8 * s, ph ^ 1u, trap, RT_W_EMPTY); else mbar_wait(W_EMPTY + 8 * s, ph ^ 1u, trap, RT_W_EMPTY);
        if (DIAG && (a.flags & 512) && j == 6 && ((it / RO_NCH) % 5u) == 2u) {   // fault injection: a late weight chunk (tests/test_gpu_parity.py)
          const long long t0 = clock64();
          while (clock64() - t0 < 6000) {}
        }
        mbar_arrive_expect_tx(W_FULL + 8 * s, RO_STAGE_BYTES);
        bulk_g2s(sbase + RS_W + s * RO_STAGE_BYTES, a.Wp + (size_t)(j - j % RO_NKC + ro_perm(j % RO_NKC)) * RO_STAGE_BYTES,
                 RO_STAGE_BYTES, W_FULL + 8 * s);
      }
    }
  } else if (warp >= RO_WWARPS + RO_NPROD) {
    // =============================== MMA issuers ===============================
    // Issuer 0: aggregation + gate passes 0, 2; issuer 1: gate passes 1, 3.  Two issuing threads reach the
    // nominal 48 clk per M128 x N96 MMA where one tops out at ~68 (scratch/mma_bench3.cu).  The whole warp runs the
    // code convergently; one elected lane issues each tcgen05 instruction.  Which issuer a warp is, the pass and the
    // chunk are compile-time constants of two separate instantiations, so that every MMA operand (descriptors, tensor-
    // memory addresses, barrier addresses) is warp-uniform to the compiler and lives in uniform registers: with the
    // issuer index taken from threadIdx at run time each MMA cost ~10 extra instructions (R2UR.BROADCAST per operand).
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    auto issuer = [&](auto me_tag) {
      constexpr int me = decltype(me_tag)::value;
      uint32_t sc = 0;
      const uint64_t d_att0 = make_desc_sw128(sbase + RS_ATT), d_att1 = make_desc_sw128(sbase + RS_ATT + RO_BLK);
      const uint64_t d_h = make_desc_sw128_mn(sbase + RS_H, RO_BLK, 1024);
      [[maybe_unused]] const uint64_t d_c = make_desc_sw128_mn(sbase + RS_C, RO_BLK, 1024);
      const uint64_t d_w = make_desc_sw128(sbase + RS_W);
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x)
        for (int t = 0; t < nsteps; ++t, ++sc) {
          const uint32_t par = sc & 1u;
          long long* dbg = (DIAG && a.dbg && blockIdx.x == 0 && lane == 0 && sc < 64) ? a.dbg + sc * 32 + 16 : nullptr;
          const bool timed = DIAG && dbg && (a.flags & 16);
          long long wwait = 0;
          const uint32_t pc0 = sc * RO_NP;
          // One (pass, k-chunk) of the gate GEMM: wait for its weight stage, 4 MMAs (A = 32 TMEM columns), release the
          // stage.  j = 5 p + kcn is the chunk's position in the step; chunks are consumed in the order h0 h1 | e | mh0 mh1
          // (ro_perm): the h part only needs the previous step.  20 chunks per step over 4 stages: the stage is j % 4
          // and the parity of its use (5 sc + j / 4) & 1.
          auto gate_chunk = [&](uint32_t d_tmem, auto j_tag) {
            constexpr int j = decltype(j_tag)::value, kcn = j % RO_NKC;
            constexpr uint32_t s = j % RO_NSTAGE;
            static_assert(RO_NCH % RO_NSTAGE == 0, "stage of a chunk must not depend on the step");
            if (!(DIAG && (a.flags & 1))) {
              const long long w0 = timed ? clock64() : 0;
              // no tcgen05.fence here: the stage was written by the async proxy (bulk copy -> mbarrier), not by another
              // thread's tcgen05 operation; a fence per chunk drained the MMA pipeline (pass 2650 -> see profiles/)
              if (!(DIAG && (a.flags & 128))) mbar_wait_warp(W_FULL + 8 * s, (sc + (j / RO_NSTAGE)) & 1u, trap, RT_W_FULL);   // 128: timing experiment
              if (timed) wwait += clock64() - w0;
            }
            const uint64_t db = d_w + (uint64_t)((s * RO_STAGE_BYTES) >> 4);
            const uint32_t acol = tmem_u + RT_A + (uint32_t)ro_perm(kcn) * 32u;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16_ts_elect(d_tmem, acol + ks * 8, db + (uint64_t)(ks * 2), kIdescGate<F16>, (kcn | ks) ? 1u : 0u);
            if (!(DIAG && (a.flags & 1)) || (a.flags & 256)) umma_commit_elect(W_EMPTY + 8 * s);
          };
          auto gate_pass = [&](auto p_tag) {   // passes 1-3: all five chunks back to back
            constexpr int p = decltype(p_tag)::value;
            const uint32_t pc = pc0 + p;       // global pass counter: accumulator p & 1, use number pc >> 1
            constexpr uint32_t b = p & 1u;
            mbar_wait_warp(ACC_EMPTY + 8 * b, ((pc >> 1) & 1u) ^ 1u, trap, RT_ACC_EMPTY);
            // The weight ring is ONE ring shared by the two issuers, and an mbarrier parity wait is only meaningful for a
            // waiter at most one phase ahead: chunk j of this pass re-uses the stage of chunk j - 4 of the previous pass,
            // which the OTHER issuer consumes.  Were that chunk still in flight when this issuer reached its wait, the
            // parity of the phase before it would read as "complete", the MMAs would run on a half-written stage and
            // the extra W_EMPTY arrival would derail the producers (seen as an expired bounded wait = CUDA 719 once the
            // weight stream ran > ~1000 clk late).  So a pass starts only after the previous pass has been issued whole.
#ifndef RO_NO_PASS_ORDER   // (defined only by scratch/ro_race.py's build of the pre-fix protocol)
            if constexpr (p == 2) mbar_wait_warp(P1_ISSUED, par, trap, RT_P1_ISSUED);
            if constexpr (p == 3) mbar_wait_warp(P2_ISSUED, par, trap, RT_P2_ISSUED);
#endif
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + (b ? RT_ACC1 : RT_ACC0);
            if (DIAG && dbg) dbg[3 + 2 * p] = clock64();
            gate_chunk(d_tmem, std::integral_constant<int, p * RO_NKC + 0>{});
            gate_chunk(d_tmem, std::integral_constant<int, p * RO_NKC + 1>{});
            gate_chunk(d_tmem, std::integral_constant<int, p * RO_NKC + 2>{});
            gate_chunk(d_tmem, std::integral_constant<int, p * RO_NKC + 3>{});
            gate_chunk(d_tmem, std::integral_constant<int, p * RO_NKC + 4>{});
            umma_commit_elect(ACC_FULL + 8 * b);
            if constexpr (p == 1 || p == 2) {
              if (lane == 0) mbar_arrive(p == 1 ? P1_ISSUED : P2_ISSUED);
              __syncwarp();
            }
            if (DIAG && dbg) dbg[4 + 2 * p] = clock64();
          };
          if constexpr (me == 0) {
            // ---- pass 0 (accumulator 0, free since the previous step's pass-2 epilogue).  Its h chunks go first, while
            //      the workers still build the attention: h' of the previous step is in tensor memory once the workers
            //      have released accumulator 1 after its last pass (on a tile's first step: once E_READY has published
            //      the zeroed state).
            const uint32_t d0 = tmem_u + RT_ACC0;
            mbar_wait_warp(ACC_EMPTY, ((pc0 >> 1) & 1u) ^ 1u, trap, RT_ACC_EMPTY_P0);
            if (t > 0) {
              mbar_wait_warp(ACC_EMPTY + 8, (((pc0 + 1) >> 1) & 1u) ^ 1u, trap, RT_ACC1_EMPTY_P0);
              tc_fence_after();
              if (DIAG && dbg) dbg[3] = clock64();
              gate_chunk(d0, std::integral_constant<int, 0>{});
              gate_chunk(d0, std::integral_constant<int, 1>{});
            }
            mbar_wait_warp(ATT_READY, par, trap, RT_ATT_READY);
            tc_fence_after();
            if (DIAG && dbg) dbg[0] = clock64();
#if RO_AGG256
            // ---- aggregation: [mh | mc] = att x [h | c] as ONE N = 256 MMA per k-step (the H and C images are four
            //      contiguous 64-unit blocks, the two accumulators 256 contiguous columns): the attention operand
            //      crosses shared memory once instead of twice
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16_elect(tmem_u + RT_MH, (ks < 4 ? d_att0 : d_att1) + (uint64_t)((ks & 3) * 2), d_h + (uint64_t)(ks * 128),
                        kIdescAgg256<F16>, ks ? 1u : 0u);
            umma_commit_elect(AGG_FULL);
#else
            // ---- aggregation: mh = att x h (committed first: its conversion is on the critical path), mc = att x c
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16_elect(tmem_u + RT_MH, (ks < 4 ? d_att0 : d_att1) + (uint64_t)((ks & 3) * 2), d_h + (uint64_t)(ks * 128),
                        kIdescAggMN<F16>, ks ? 1u : 0u);
            umma_commit_elect(AGG_FULL);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16_elect(tmem_u + RT_MC, (ks < 4 ? d_att0 : d_att1) + (uint64_t)((ks & 3) * 2), d_c + (uint64_t)(ks * 128),
                        kIdescAggMN<F16>, ks ? 1u : 0u);
#endif
            if (DIAG && dbg) dbg[1] = clock64();
            mbar_wait_warp(E_READY, par, trap, RT_E_READY);
            tc_fence_after();
            if (DIAG && dbg) dbg[2] = clock64();
            if (t == 0) {
              if (DIAG && dbg) dbg[3] = clock64();
              gate_chunk(d0, std::integral_constant<int, 0>{});
              gate_chunk(d0, std::integral_constant<int, 1>{});
            }
            gate_chunk(d0, std::integral_constant<int, 2>{});
            mbar_wait_warp(MH_READY, par, trap, RT_MH_READY_I0);   // the mh chunks wait for the conversion
            tc_fence_after();
            if (DIAG && dbg) dbg[12] = clock64();
            gate_chunk(d0, std::integral_constant<int, 3>{});
            gate_chunk(d0, std::integral_constant<int, 4>{});
            umma_commit_elect(ACC_FULL);
            if (DIAG && dbg) dbg[4] = clock64();
            // pass 1 queues behind pass 0 in the tensor pipe: interleaved, both ran at half rate and accumulator 0,
            // which the workers wait for first, completed ~1200 clk later
            if (lane == 0) mbar_arrive(P0_ISSUED);
            __syncwarp();
            gate_pass(std::integral_constant<int, 2>{});
          } else {
            mbar_wait_warp(MH_READY, par, trap, RT_MH_READY_I1);   // accumulator 1 aliases the mh accumulator: wait for its conversion
            mbar_wait_warp(P0_ISSUED, par, trap, RT_P0_ISSUED);
            tc_fence_after();
            gate_pass(std::integral_constant<int, 1>{});
            gate_pass(std::integral_constant<int, 3>{});
          }
          if (DIAG && dbg) dbg[13 + me] = wwait;
        }
    };
    if (warp == RO_WWARPS + RO_NPROD) issuer(std::integral_constant<int, 0>{}); else issuer(std::integral_constant<int, 1>{});
  } else {
    // =============================== workers ===============================
    const int q = warp & 3, cs = warp >> 2;
    const int r = q * 32 + lane;           // this thread's row: TMEM lane, attention row, epilogue row
    const int N = a.N;
    const int sb = (r / N) * N;            // first row of this row's scene inside the tile
    const int nch = N >> 3;                // 8-column attention chunks per row; this thread takes chunks cs, cs + 4, ...
    // the diagonal entry (j == r) of the attention row is masked out of the packed bf16 words of its 8-column chunk
    const int jdiag8 = (r - sb) & ~7;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* const cf_row = smem + RS_CF + r * 16;     // + (u >> 2) * 2048: 4 fp32 c of units u .. u+3
    uint32_t sc = 0;

    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int row0 = tile * 128;
      const int gr = row0 + r;
      const bool rok = gr < a.R;
      const bool v = rok && a.valid[gr] != 0;
      worker_sync();   // every worker has finished the previous tile before the reset
      // zero the recurrent state: h, c (bf16 images + fp32 c) in shared memory and the h columns of the TMEM A operand
      {
        uint4* hz = reinterpret_cast<uint4*>(smem + RS_H);   // H | C contiguous: 64 KB
        uint4* cz = reinterpret_cast<uint4*>(smem + RS_CF);  // 64 KB
#pragma unroll
        for (int k = 0; k < 4 * RO_BLK / 16 / RO_WORKERS; ++k) {
          hz[tid + RO_WORKERS * k] = make_uint4(0, 0, 0, 0);
          cz[tid + RO_WORKERS * k] = make_uint4(0, 0, 0, 0);
        }
        const uint32_t z8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tmem_st8(t_row + RT_A_H + cs * 16, z8);
        tmem_st8(t_row + RT_A_H + cs * 16 + 8, z8);
      }
      // observed frames arrive through shared memory: (pos, vislet) of frame t+1 is fetched with cp.async during
      // step t by the slice-0 / slice-1 thread of each row (no registers held across the step)
      float2 prevp = make_float2(0.f, 0.f), visv = make_float2(0.f, 0.f);
      if (cs < 2) {
        float2 f0 = make_float2(0.f, 0.f);
        if (rok) f0 = cs == 0 ? __ldg(reinterpret_cast<const float2*>(a.pos) + (size_t)gr * a.F)
                              : __ldg(reinterpret_cast<const float2*>(a.vis) + (size_t)gr * a.T);
        reinterpret_cast<float2*>(s_obs + r)[cs] = f0;
      }
      worker_sync();

      for (int t = 0; t < nsteps; ++t, ++sc) {
        const uint32_t par = sc & 1u;
        const bool emit = t >= a.T - 1;
        long long* dbg = (DIAG && a.dbg && blockIdx.x == 0 && tid == ((a.flags >> 8) & 15) * 32 && sc < 64) ? a.dbg + sc * 32 : nullptr;
        if (DIAG && dbg) dbg[0] = clock64();
        // ---- (a) current position and the cell input x = [cur - prev | vislet] of row r
        float2 cur;
        if (t < a.T) {
          const float4 ob = s_obs[r];
          cur = make_float2(ob.x, ob.y);
          visv = make_float2(ob.z, ob.w);
        } else {
          cur = s_next[r];
        }
        float4 xv = make_float4(0.f, 0.f, visv.x, visv.y);
        if (t > 0) {
          xv.x = __fsub_rn(cur.x, prevp.x);
          xv.y = __fsub_rn(cur.y, prevp.y);
        }
        prevp = cur;
        if (!v) xv = make_float4(0.f, 0.f, 0.f, 0.f);   // whatever an invalid slot holds (NaN included) stays out of the state
        if (cs == 0) {   // invalid agents sit far away: d2 = inf fails d2 < r2 for every partner
          s_px[r] = v ? cur.x : 3.0e18f;
          s_py[r] = v ? cur.y : 3.0e18f;
        }
        worker_sync();   // positions (and, on a tile's first step, the zeroed state) visible; s_obs consumed
        if (t + 1 < a.T && rok && cs < 2) {   // prefetch the next observed frame
          const float* src = cs == 0 ? a.pos + ((size_t)gr * a.F + (t + 1)) * 2 : a.vis + ((size_t)gr * a.T + (t + 1)) * 2;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(reinterpret_cast<float2*>(s_obs + r) + cs)),
                       "l"(src)
                       : "memory");
        }
        // ---- (b) attention row r against 8-column chunks cs, cs+4, .. of its scene (un-normalised), packed fp32x2 math
        {
          const float2 nx = make_float2(-cur.x, -cur.x), ny = make_float2(-cur.y, -cur.y);
          const float2 cexp = make_float2(a.neg_inv_log2e, a.neg_inv_log2e);
          float sum = 0.f;
          for (int ch = cs; ch < nch; ch += 4) {
            const int j8 = ch << 3;
            uint32_t pk[4];
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
              const float4 xj = *reinterpret_cast<const float4*>(s_px + sb + j8 + hq * 4);
              const float4 yj = *reinterpret_cast<const float4*>(s_py + sb + j8 + hq * 4);
#pragma unroll
              for (int pr = 0; pr < 2; ++pr) {
                const float2 dx = fadd2(pr ? make_float2(xj.z, xj.w) : make_float2(xj.x, xj.y), nx);
                const float2 dy = fadd2(pr ? make_float2(yj.z, yj.w) : make_float2(yj.x, yj.y), ny);
                const float2 d2 = fadd2(fmul2(dx, dx), fmul2(dy, dy));
                const float2 ka = fmul2(d2, cexp);
                const float2 kern = make_float2(ex2_fast(ka.x), ex2_fast(ka.y));   // exp(-d2 / 2 sigma^2)
                // exp(kern), kern in (0, 1]: cubic on the FMA pipe (max relative error 3.2e-4, below the bf16 rounding
                // of the operand) instead of a second MUFU per pair
                // (fp16 operands round at 4.9e-4: there a quartic, 1.6e-5)
                float2 ek;
                if constexpr (F16)
                  ek = ffma2(ffma2(ffma2(ffma2(make_float2(0.0679839998f, 0.0679839998f), kern, make_float2(0.143049359f, 0.143049359f)),
                                         kern, make_float2(0.50812006f, 0.50812006f)), kern, make_float2(0.99906832f, 0.99906832f)),
                             kern, make_float2(1.00001609f, 1.00001609f));
                else
                  ek = ffma2(ffma2(ffma2(make_float2(0.27136664f, 0.27136664f), kern, make_float2(0.43417813f, 0.43417813f)),
                                   kern, make_float2(1.01218117f, 1.01218117f)), kern, make_float2(0.99967653f, 0.99967653f));
                const float e0 = (v && d2.x < a.r2) ? ek.x : 0.f;                 // softmax numerator
                const float e1 = (v && d2.y < a.r2) ? ek.y : 0.f;
                pk[hq * 2 + pr] = pack_op2<F16>(e0, e1);
              }
            }
            if (j8 == jdiag8) {
              const int dw = ((r - sb) & 7) >> 1;
              const uint32_t dm = ((r - sb) & 1) ? 0x0000FFFFu : 0xFFFF0000u;
#pragma unroll
              for (int w = 0; w < 4; ++w) pk[w] &= (w == dw) ? dm : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) sum += op_lo<F16>(pk[w]) + op_hi<F16>(pk[w]);   // normalise by what the MMA really sums
            const int jt = sb + j8;
            *reinterpret_cast<uint4*>(smem + RS_ATT + (jt >> 6) * RO_BLK + sw128_off(r, jt & 63)) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          s_sum[cs * 128 + r] = sum;
        }
        fence_proxy_async();   // generic-proxy smem writes (att; h', c' of the previous step) -> async proxy
        tc_fence_before();     // (and the h' tcgen05.st of the previous step, already waited for)
        mbar_arrive_warp(ATT_READY);
        if (DIAG && dbg) dbg[1] = clock64();
        asm volatile("cp.async.wait_all;" ::: "memory");
        worker_sync();   // all four partial sums of every attention row are written; next observed frame landed
        const float ssum = (s_sum[r] + s_sum[128 + r]) + (s_sum[256 + r] + s_sum[384 + r]);
        const float inv = ssum > 0.f ? __fdividef(1.0f, ssum) : 0.f;
        // ---- (c) e = relu(x W_e + b_e): row r, k in [16 cs, 16 cs + 16) -> A-operand columns 8 cs .. +7.  Computed while
        //      the aggregation MMAs execute; no worker barrier between here and MH_READY, so a warp that finishes early
        //      starts its conversion early (placing e before the attention build delayed the aggregation: +900 clk)
        uint32_t pe[8];
        {
#pragma unroll
          for (int hq = 0; hq < 4; ++hq) {
            const int k = cs * 16 + hq * 4;
            const float4 w0 = *reinterpret_cast<const float4*>(s_we + k);
            const float4 w1 = *reinterpret_cast<const float4*>(s_we + 64 + k);
            const float4 w2 = *reinterpret_cast<const float4*>(s_we + 128 + k);
            const float4 w3 = *reinterpret_cast<const float4*>(s_we + 192 + k);
            const float4 bb = *reinterpret_cast<const float4*>(s_we + 256 + k);
            float e0 = fmaf(xv.w, w3.x, fmaf(xv.z, w2.x, fmaf(xv.y, w1.x, fmaf(xv.x, w0.x, bb.x))));
            float e1 = fmaf(xv.w, w3.y, fmaf(xv.z, w2.y, fmaf(xv.y, w1.y, fmaf(xv.x, w0.y, bb.y))));
            float e2 = fmaf(xv.w, w3.z, fmaf(xv.z, w2.z, fmaf(xv.y, w1.z, fmaf(xv.x, w0.z, bb.z))));
            float e3 = fmaf(xv.w, w3.w, fmaf(xv.z, w2.w, fmaf(xv.y, w1.w, fmaf(xv.x, w0.w, bb.w))));
            e0 = rok ? fmaxf(e0, 0.f) : 0.f;
            e1 = rok ? fmaxf(e1, 0.f) : 0.f;
            e2 = rok ? fmaxf(e2, 0.f) : 0.f;
            e3 = rok ? fmaxf(e3, 0.f) : 0.f;
            pe[hq * 2] = pack_op2<F16>(e0, e1);
            pe[hq * 2 + 1] = pack_op2<F16>(e2, e3);
          }
        }
        tmem_st8(t_row + RT_A + cs * 8, pe);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_warp(E_READY);
        if (DIAG && dbg) dbg[2] = clock64();
        // ---- (d) mh: accumulator -> normalise -> bf16 -> A-operand columns (this thread: row r, units 32 cs .. +31)
        mbar_wait(AGG_FULL, par, trap, RT_AGG_FULL);
        tc_fence_after();
        if (DIAG && dbg) dbg[3] = clock64();
#pragma unroll
        for (int ch = 0; ch < 4; ch += 2) {
          float v0[8], v1[8];
          tmem_ld8(t_row + RT_MH + cs * 32 + ch * 8, v0);
          tmem_ld8(t_row + RT_MH + cs * 32 + ch * 8 + 8, v1);
          tmem_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            pk[i] = pack_op2<F16>(v0[2 * i] * inv, v0[2 * i + 1] * inv);
            pk[4 + i] = pack_op2<F16>(v1[2 * i] * inv, v1[2 * i + 1] * inv);
          }
          tmem_st8(t_row + RT_A_MH + cs * 16 + ch * 4, pk);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_warp(MH_READY);
        if (DIAG && dbg) dbg[4] = clock64();

        // ---- (e) gate epilogue: 4 passes; this thread: row r, units 32 p + 8 cs .. +7, four at a time
        float2 y2[5];
#pragma unroll
        for (int z = 0; z < 5; ++z) y2[z] = make_float2(0.f, 0.f);
        const float2 kHalf = make_float2(0.5f, 0.5f), kNegHalf = make_float2(-0.5f, -0.5f), kOne = make_float2(1.f, 1.f);
        const float2 inv2 = make_float2(inv, inv);
        // Branch-free body (emitting / observed steps are two instantiations): one basic block per pass, so the two unit quads of a thread interleave and the
        // MUFU results of one hide behind the FMAs of the other.
        auto gate_epilogue = [&](auto emit_tag) {
          constexpr bool EMIT = decltype(emit_tag)::value;
#pragma unroll 1
          for (int p = 0; p < RO_NP; ++p) {
            const uint32_t pc = sc * RO_NP + p;
            const uint32_t b = p & 1u, bph = (pc >> 1) & 1u;
            mbar_wait(ACC_FULL + 8 * b, bph, trap, RT_ACC_FULL);
            tc_fence_after();
            if (DIAG && dbg) dbg[5 + 2 * p] = clock64();
            const uint32_t t_acc = t_row + (b ? RT_ACC1 : RT_ACC0) + cs * 8;
            const int u0 = p * RO_UN + cs * 8;        // first of this thread's 8 units
            uint32_t hw[4], cw[4];                    // h', c' as bf16 pairs
            float zi[2][4], zj[2][4], zo[2][4], zm[2][4];   // [unit quad][unit]: one 8-column load per array
            tmem_ld8(t_acc, &zi[0][0]);
            tmem_ld8(t_acc + RO_UN, &zj[0][0]);
            tmem_ld8(t_acc + 2 * RO_UN, &zo[0][0]);
            tmem_ld8(t_row + RT_MC + u0, &zm[0][0]);
            tmem_wait_ld();
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
              const int u = u0 + hq * 4;
              float4 c4 = *reinterpret_cast<const float4*>(cf_row + (u >> 2) * 2048);
              float ho[4], fo[4];
              const float4 bI = *reinterpret_cast<const float4*>(s_bias + u);
              const float4 bJ = *reinterpret_cast<const float4*>(s_bias + 128 + u);
              const float4 bO = *reinterpret_cast<const float4*>(s_bias + 256 + u);
              const float4 pIf = *reinterpret_cast<const float4*>(s_bias + 384 + u);
              const float4 pIt = *reinterpret_cast<const float4*>(s_bias + 512 + u);
              const float4 pOf = *reinterpret_cast<const float4*>(s_bias + 640 + u);
              const float4 pOt = *reinterpret_cast<const float4*>(s_bias + 768 + u);
#pragma unroll
              for (int pr = 0; pr < 2; ++pr) {
                const int i0 = pr * 2;
                auto sel = [&](const float4& f) { return pr ? make_float2(f.z, f.w) : make_float2(f.x, f.y); };
                // sigmoid = 0.5 + 0.5 tanh (the inner 1/2 is folded into weights/biases/peepholes):
                //   c_x' = x + g (tj - x) = x + (1 + th) d,  d = (tj - x) / 2        (x = mc, c)
                //   h'   = q tanh(c_t') = to a + a,          a = tanh(c_t') / 2
                const float2 c2 = sel(c4);
                const float2 m2 = fmul2(make_float2(zm[hq][i0], zm[hq][i0 + 1]), inv2);
                const float2 ai = ffma2(sel(pIf), m2, ffma2(sel(pIt), c2, fadd2(make_float2(zi[hq][i0], zi[hq][i0 + 1]), sel(bI))));
                const float2 th = tanh2(ai);
                const float2 tjh = fmul2(tanh2(fadd2(make_float2(zj[hq][i0], zj[hq][i0 + 1]), sel(bJ))), kHalf);
                const float2 dm = ffma2(m2, kNegHalf, tjh), dc = ffma2(c2, kNegHalf, tjh);
                const float2 g2 = fadd2(th, kOne);                     // 2 g
                const float2 cf = ffma2(dm, g2, m2);                   // (1-g) mc + g tanh j
                const float2 ct = ffma2(dc, g2, c2);                   // (1-g) c  + g tanh j
                const float2 o1 = ffma2(sel(pOf), cf, fadd2(make_float2(zo[hq][i0], zo[hq][i0 + 1]), sel(bO)));
                const float2 to = tanh2(ffma2(sel(pOt), ct, o1));
                // rows of invalid agents carry a bounded state of their own (x = 0 input, no neighbours, nobody's neighbour:
                // their attention column is zero) instead of being re-zeroed by 16 selects per pass
                if constexpr (EMIT) {
                  const float2 q2 = ffma2(to, kHalf, kHalf);           // output gate
                  // tanh(c_f) stays on the MUFU pipe: with both on the FMA pipe the emitting pass becomes issue-bound (2.02 ms)
                  const float2 h2 = fmul2(tanh_unit2(ct), q2), f2 = fmul2(tanh2(cf), q2);
                  ho[i0] = h2.x; ho[i0 + 1] = h2.y;
                  fo[i0] = f2.x; fo[i0 + 1] = f2.y;
                } else {
                  const float2 ha = half_tanh_unit2(ct);
                  const float2 h2 = ffma2(to, ha, ha);
                  ho[i0] = h2.x; ho[i0 + 1] = h2.y;
                }
                if (pr) { c4.z = ct.x; c4.w = ct.y; } else { c4.x = ct.x; c4.y = ct.y; }
              }
              *reinterpret_cast<float4*>(cf_row + (u >> 2) * 2048) = c4;
              if constexpr (EMIT) {
                // head partial sums, two units per packed FMA: y2[z] += (v_u, v_u+1) * (W_hT[z][u], W_hT[z][u+1])
                // (rows of invalid agents accumulate values nobody reads)
#pragma unroll
                for (int hsrc = 0; hsrc < 2; ++hsrc) {
                  const float2 va = hsrc ? make_float2(fo[0], fo[1]) : make_float2(ho[0], ho[1]);
                  const float2 vb = hsrc ? make_float2(fo[2], fo[3]) : make_float2(ho[2], ho[3]);
#pragma unroll
                  for (int z = 0; z < 5; ++z) {
                    const float4 w4 = *reinterpret_cast<const float4*>(s_wht + z * 256 + hsrc * RO_U + u);
                    y2[z] = ffma2(va, make_float2(w4.x, w4.y), y2[z]);
                    y2[z] = ffma2(vb, make_float2(w4.z, w4.w), y2[z]);
                  }
                }
              }
              hw[hq * 2] = pack_op2<F16>(ho[0], ho[1]);
              hw[hq * 2 + 1] = pack_op2<F16>(ho[2], ho[3]);
              cw[hq * 2] = pack_op2<F16>(c4.x, c4.y);
              cw[hq * 2 + 1] = pack_op2<F16>(c4.z, c4.w);
            }
            {
              // c', h' (bf16) -> shared-memory B operands of the next step's aggregation (its MMAs of this step are
              // complete; the gate MMAs read h from TMEM, whose copy follows after the last pass)
              const uint32_t so = (u0 >> 6) * RO_BLK + r * 128 + ((((u0 & 63) >> 3) ^ (r & 7)) << 4);
              *reinterpret_cast<uint4*>(smem + RS_C + so) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
              *reinterpret_cast<uint4*>(smem + RS_H + so) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            }
            if (p == RO_NP - 1) {
              // every gate MMA of this step has completed (ACC_FULL of the last pass): the h columns of the TMEM A
              // operand may be overwritten.  Each thread re-reads the 4 x 8 units it stored (its own writes) and
              // copies them: units u, u+1 -> column u/2.
#pragma unroll
              for (int pp = 0; pp < RO_NP; ++pp) {
                const int u = pp * RO_UN + cs * 8;
                const uint4 t4 = *reinterpret_cast<const uint4*>(smem + RS_H + (u >> 6) * RO_BLK + r * 128 +
                                                                 ((((u & 63) >> 3) ^ (r & 7)) << 4));
                const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
                tmem_st4(t_row + RT_A_H + pp * 16 + cs * 4, w4);
              }
              tmem_wait_st();
            }
            tc_fence_before();
            mbar_arrive_warp(ACC_EMPTY + 8 * b);
          }
        };
        if (emit) gate_epilogue(std::true_type{}); else gate_epilogue(std::false_type{});
        if (DIAG && dbg) dbg[13] = clock64();
        // ---- (f) head: combine the four column slices of each row, emit the 5 parameters and the next position
        if (emit) {
          {
            uint32_t yw[8];
#pragma unroll
            for (int z = 0; z < 5; ++z) yw[z] = __float_as_uint(y2[z].x + y2[z].y);
            yw[5] = yw[6] = yw[7] = 0u;
            tmem_st8(t_row + RT_HEAD + cs * 8, yw);
            tmem_wait_st();
            tc_fence_before();
          }
          worker_sync();
          if (cs == 0) {
            tc_fence_after();
            float p0[8], p1[8], p2[8], p3[8];
            tmem_ld8(t_row + RT_HEAD, p0);
            tmem_ld8(t_row + RT_HEAD + 8, p1);
            tmem_ld8(t_row + RT_HEAD + 16, p2);
            tmem_ld8(t_row + RT_HEAD + 24, p3);
            tmem_wait_ld();
            float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (v) {
#pragma unroll
              for (int z = 0; z < 5; ++z) o[z] = (p0[z] + p1[z]) + (p2[z] + p3[z]) + __ldg(a.b_h + z);
              o[2] = __expf(o[2]);
              o[3] = __expf(o[3]);
              o[4] = tanh_fast(o[4]);
            }
            if (rok) {
              float* po = a.params + ((size_t)gr * a.P + (t - (a.T - 1))) * 5;
#pragma unroll
              for (int z = 0; z < 5; ++z) po[z] = o[z];
            }
            s_next[r] = make_float2(cur.x + o[0], cur.y + o[1]);
            tc_fence_before();
          }
          worker_sync();
        }
        if (DIAG && dbg) dbg[14] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RO_WWARPS) tmem_dealloc(tmem_base, 512);
}

// pos[R,F,2], vis[R,T,2], valid[R] -> params[R,P,5].  Requires 128 % N == 0, N >= 8, U = 128, E = 64.
int launch_rollout_tc(const float* pos, const float* vis, const uint8_t* valid, const mmt_cell_weights* w, int S, int N,
                      int T, int P, float r2, float inv_2sigma2, float* params, long long* dbg, int f16, cudaStream_t stream) {
  RoArgs a = {};
  a.pos = pos; a.vis = vis; a.valid = valid;
  a.W_e = w->W_e; a.b_e = w->b_e; a.b = w->b; a.w_If = w->w_If; a.w_It = w->w_It; a.w_Of = w->w_Of; a.w_Ot = w->w_Ot;
  a.W_h = w->W_h; a.b_h = w->b_h; a.Wp = reinterpret_cast<const uint8_t*>(f16 ? w->W_packed_f16 : w->W_packed_bf16);
  a.params = params;
  a.R = S * N; a.N = N; a.T = T; a.P = P; a.F = T + P;
  a.num_tiles = (a.R + 127) / 128;
  a.r2 = r2; a.neg_inv_log2e = -inv_2sigma2 * 1.4426950408889634f;
  a.dbg = dbg;
  a.trap = trap_record();
  // diagnostic switches of scratch/ro_*.py (timing experiments, fault injection): read once per process
  static std::atomic<int> env_flags{INT_MIN}, env_grid{INT_MIN};
  a.flags = env_int_once("MMT_RO_FLAGS", &env_flags);
  static DeviceMask smem_opted[4];   // per kernel: devices already opted in
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&rollout_tc_kernel<false, false>), RS_TOTAL + 1024, &smem_opted[0])) return rc;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&rollout_tc_kernel<true, false>), RS_TOTAL + 1024, &smem_opted[1])) return rc;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&rollout_tc_kernel<false, true>), RS_TOTAL + 1024, &smem_opted[2])) return rc;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&rollout_tc_kernel<true, true>), RS_TOTAL + 1024, &smem_opted[3])) return rc;
  int grid = a.num_tiles < num_sms() ? a.num_tiles : num_sms();
  const int dgrid = env_int_once("MMT_RO_GRID", &env_grid);
  if (dgrid > 0 && dgrid < grid) grid = dgrid;
  const bool diag = a.dbg || a.flags;
  if (f16) {
    if (diag) rollout_tc_kernel<true, true><<<grid, RO_THREADS, RS_TOTAL + 1024, stream>>>(a);
    else rollout_tc_kernel<false, true><<<grid, RO_THREADS, RS_TOTAL + 1024, stream>>>(a);
  } else {
    if (diag) rollout_tc_kernel<true, false><<<grid, RO_THREADS, RS_TOTAL + 1024, stream>>>(a);
    else rollout_tc_kernel<false, false><<<grid, RO_THREADS, RS_TOTAL + 1024, stream>>>(a);
  }
  count_launch();
  return check_launch("rollout_tc_kernel");
}

}  // namespace mmt
