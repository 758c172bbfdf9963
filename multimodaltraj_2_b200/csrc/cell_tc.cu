// gsk_lstm_cell on the 5th-generation tensor cores (sm_100a): bf16 operands, fp32 accumulation in
// TMEM, gate update + head fused as the epilogue (include/mmt.h mmt_gsk_cell
// with MMT_PREC_BF16).
//
// Tile = 128 rows (agents) x 384 gate columns x K = E + 2U = 320.
//   * A operand [e | h | mh] is built in shared memory by the worker warps (e = relu(x W_e + b_e)
//     computed on the fly, h / mh converted fp32 -> bf16) in the canonical K-major SWIZZLE_128B
//     layout (5 blocks of [128 rows x 128 B]).  It is never written to HBM.
//   * B operand (the gate weights) is pre-packed once by mmt_pack_gate_weights_bf16 into the exact
//     shared-memory image, ordered [pass][k-chunk]; a producer thread streams 12 KB stages with
//     cp.async.bulk (TMA bulk copy, SASS UBLKCP) completing on mbarriers.  The image is 240 KB and
//     stays L2-resident.
//   * The 384 gate columns are processed as 4 passes of 32 units; the packed column order puts the
//     i, j, o columns of a pass next to each other so one tcgen05.mma (M=128, N=96, K=16) per
//     k-step feeds a 96-column accumulator; two accumulators (TMEM columns 0 and 128) are
//     double-buffered so the epilogue of pass p overlaps the MMAs of pass p+1.
//   * Epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).. (its rows) and units 16*(w/4).. of the
//     pass with tcgen05.ld.32x32b.x8, applies the peephole gates, writes h', c' (and m_f) and
//     accumulates the 5-wide head.  2 CTAs are resident per SM (256 TMEM columns, ~105 KB smem each)
//     so one CTA's operand build overlaps the other's MMA/epilogue.
#include <cuda_bf16.h>

#include "mmt_common.cuh"
#include "tc_common.cuh"

namespace mmt {

constexpr int TC_E = 64, TC_U = 128, TC_K = TC_E + 2 * TC_U;  // 320
constexpr int TC_M = 128;                                     // rows per tile
constexpr int TC_UN = 32;                                     // units per pass
constexpr int TC_NP = TC_U / TC_UN;                           // 4 passes
constexpr int TC_N = 3 * TC_UN;                               // 96 accumulator columns per pass
constexpr int TC_KC = 64;                                     // k-chunk (one 128-byte swizzle row)
constexpr int TC_NKC = TC_K / TC_KC;                          // 5
constexpr int TC_STAGE_BYTES = TC_N * TC_KC * 2;              // 12288
constexpr int TC_NSTAGE = 2;                                  // weight stages of the bf16 kernel (two CTAs per SM)
#ifndef TC_X3_STAGING
#define TC_X3_STAGING 1   // 1: c / mc / h' / c' of the epilogue staged through 33 KB of shared memory (coalesced global access), which
#endif                    //    leaves room for only two weight stages; 0: four weight stages, per-lane global access.
// Measured on a B200 at C3 (scratch/timeline_x3.py): staging + 2 stages 0.407 ms per step (a pass = 7.0 k clk, paced by the weight
// stream: 24 KB in flight; the epilogue alone needs 4.3-6 k), no staging + 4 stages 0.527 ms (per-lane 32-byte global pieces: epilogue
// 10-12 k clk per pass).  Both at once do not fit in 227 KB next to the two 80 KB operand images.
constexpr int TC_NSTAGE_X3 = TC_X3_STAGING ? 2 : 4;
constexpr int TC_NSTAGE_MAX = 4;                              // barrier slots
constexpr int TC_A_BLOCK = TC_M * TC_KC * 2;                  // 16384 bytes per k-chunk block
constexpr int TC_A_BYTES = TC_NKC * TC_A_BLOCK;               // 81920
constexpr int TC_WORKERS = 256;                               // worker threads of the bf16 kernel (two CTAs per SM)
constexpr int TC_THREADS = TC_WORKERS + 64;                   // + producer warp + MMA warp
constexpr int TC_WORKERS_X3 = 512;                            // split-bf16 kernel: one CTA per SM, so 16 worker warps (the 8-warp
constexpr int TC_THREADS_X3 = TC_WORKERS_X3 + 64;             // version was latency-bound: 2 warps per scheduler, 57 k clk per tile)
constexpr int TC_TMEM_COLS = 256;
constexpr int TC_ACC_STRIDE = 128;                            // TMEM column stride between the two accumulators

// shared memory map (dynamic, 1024-aligned)
constexpr int SM_A = 0;
constexpr int SM_W = SM_A + TC_A_BYTES;                       // stages
constexpr int SM_BAR = SM_W + TC_NSTAGE * TC_STAGE_BYTES;     // mbarriers (16 x 8 B)
constexpr int SM_TMEM = SM_BAR + 128;                         // tmem base address
constexpr int SM_BIAS = SM_TMEM + 16;                         // b[384] w_If,w_It,w_Of,w_Ot [4][128]
constexpr int SM_WE = SM_BIAS + (384 + 512) * 4;              // W_e[4][64], b_e[64]
constexpr int SM_HEAD = SM_WE + (256 + 64) * 4;               // head partials [128][5]
constexpr int SM_TOTAL = SM_HEAD + 128 * 5 * 4;
static_assert(SM_TOTAL + 1024 <= 113 * 1024, "two CTAs per SM must fit");


// Split-bf16 ("bf16x3") mode: the fp32 operands are split a = a_hi + a_lo, w = w_hi + w_lo (bf16 each: 16 mantissa bits
// together) and the product is taken as a_hi w_hi + a_lo w_hi + a_hi w_lo in the fp32 accumulator (the dropped a_lo w_lo
// term is 2^-16 of a product that is itself rounded to 2^-24): the gate GEMM at fp32-grade accuracy on the bf16 tensor
// pipe, at three MMAs per k-step.  To the kernel it is the same GEMM with K' = 3 K: the k-chunks of a pass run
// [A_hi | A_lo | A_hi] against the packed stream [W_hi ; W_hi ; W_lo].  The epilogue then uses tanhf / expf.
// Chunk order of a pass: W_hi[kc] (kc = 0..4), each used for TWO groups of MMAs (x A_hi[kc], then x A_lo[kc]) while it
// sits in its stage, then W_lo[kc] (x A_hi[kc]): 10 streamed chunks per pass instead of 15.
constexpr int TC_X3_NKC = 2 * TC_NKC;                         // 10 weight chunks per pass
constexpr int SM_ALO = ((SM_TOTAL + 1023) / 1024) * 1024;     // A_lo blocks (x3 only), after the common map
constexpr int SM_W23 = SM_ALO + TC_A_BYTES;                   // (weight stages 2 and 3 when TC_NSTAGE_X3 == 4)
// Epilogue staging of the split-bf16 kernel: c and mc of a pass (128 rows x 32 units, fp32) come in, c' and h' go out,
// through shared memory, so that every global access of the epilogue is a coalesced 128-byte row piece instead of one
// 32-byte piece per lane (row-major fp32 state: the per-lane pieces were 36 k of the kernel's 57 k clk per tile).
// [8 unit quads][129 rows (128 + 1 pad: the transposing accesses fall on different banks)][4 units], two arrays.
constexpr int TC_STG_Q = 129;
constexpr int TC_STG_BYTES = 8 * TC_STG_Q * 16;               // 16512 per array
constexpr int SM_STG = SM_W23 + (TC_NSTAGE_X3 > 2 ? 2 * TC_STAGE_BYTES : 0);
constexpr int SM_TOTAL_X3 = SM_STG + (TC_X3_STAGING ? 2 * TC_STG_BYTES : 0);
__device__ __forceinline__ uint32_t tc_stage_off(uint32_t s) {   // byte offset of weight stage s
  return s < 2 ? (uint32_t)(SM_W + s * TC_STAGE_BYTES) : (uint32_t)(SM_W23 + (s - 2) * TC_STAGE_BYTES);
}
// tanh for the split-bf16 mode: 1 - 2 / (exp(2x) + 1) with ex2.approx / rcp.approx (two MUFU + three FMA-pipe
// instructions; absolute error < 4e-7 -- the 1e-4 tolerance of the mode is relative to O(1) states) instead of the
// ~25-instruction tanhf; saturates correctly for |x| large (exp -> inf -> 1, exp -> 0 -> -1).
__device__ __forceinline__ float tanh_acc(float x) {
  const float e = ex2_fast(x * 2.885390081777927f);           // exp(2x)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
static_assert(SM_TOTAL_X3 + 1024 <= 227 * 1024, "x3 mode: one CTA per SM");
__device__ __forceinline__ uint32_t pack_bf16x2_lo(float a, float b, uint32_t hi) {   // residuals of a pair
  return pack_bf16x2(a - bf16_lo(hi), b - bf16_hi(hi));
}

struct TcArgs {
  const float *x, *h, *c, *mh, *mc;
  const __nv_bfloat16 *hb, *mhb, *mcb;  // bf16-state mode: h, mh, mc as [R,U] bf16 (c stays fp32, row stride ld)
  __nv_bfloat16* hb_out;
  const uint8_t* valid;
  const float *W_e, *b_e, *b, *w_If, *w_It, *w_Of, *w_Ot, *W_h, *b_h;
  const uint8_t* Wp;  // packed bf16 operand image
  float *h_out, *c_out, *mf_out;
  const float* cur_pos;
  float *params_out, *next_pos;
  int R, ld, ld_mf, params_stride, num_tiles;
  long long* dbg;  // optional [CTA][tile_iter][16] clock64 timestamps of worker thread 0 (diagnostics)
  uint32_t* trap;  // host-mapped trap record (tc_common.cuh: trap_report); may be null
};

// State layouts.  LAY 0: fp32 row-major [R,ld] in/out (the mmt_gsk_cell API).  LAY 1: bf16 h/mh/mc, fp32 c,
// row-major [R,U].  LAY 2: same types, TILE-BLOCKED [tile][U/8][128 rows][8 units]: the 8 units a thread owns
// in the epilogue are contiguous and consecutive lanes (= consecutive rows) are adjacent, so every warp-level
// global access of the epilogue and of the operand build is one contiguous 512 B / 1 KB segment instead of 32
// scattered 16/32 B pieces (the L1TEX wavefront count was the limiter of the row-major version).
__host__ __device__ __forceinline__ size_t blk_off(int tile, int r, int u) {  // element offset, u % 8 == 0
  return ((size_t)(tile * (TC_U / 8) + (u >> 3)) * TC_M + r) * 8;
}

// F16: fp16 instead of bf16 in every 16-bit operand and state word of the kernel (MMT_PREC_F16; the "bf16" state buffers of
// LAY 1 / 2 then hold fp16 bits -- producers and consumers of a forecast all run in the same mode).
template <int LAY, bool X3 = false, bool F16 = false>
__global__ void __launch_bounds__(X3 ? TC_THREADS_X3 : TC_THREADS, X3 ? 1 : 2) gsk_cell_tc_kernel(TcArgs a) {
  constexpr bool BF = LAY != 0;
  static_assert(!X3 || LAY == 0, "split-bf16 mode takes the fp32 state layout");
  static_assert(!(X3 && F16), "the split mode is built on bf16 pairs");
  constexpr uint32_t kIdesc = make_idesc_op<F16>(TC_M, TC_N);
  constexpr int NKC = X3 ? TC_X3_NKC : TC_NKC;
  constexpr int NSTAGE = X3 ? TC_NSTAGE_X3 : TC_NSTAGE;
  constexpr bool STG = X3 && TC_X3_STAGING;
  constexpr int NWT = X3 ? TC_WORKERS_X3 : TC_WORKERS;   // worker threads
  constexpr int NW = NWT / 32;                            // worker warps: 8 or 16; producer = warp NW, MMA issuer = warp NW + 1
  constexpr int NTHR = NWT + 64;
  constexpr int NH = NW / 4;                              // column slices of a pass (one per group of four warps)
  constexpr int UPT = TC_UN / NH;                         // units per thread per pass: 16 or 8
  constexpr int NSUB = UPT / 8;                           // 8-unit sub-chunks per pass: 2 or 1
  // SWIZZLE_128B atoms need a 1024-byte aligned base: requested from the toolchain, so that the base is a link-time
  // constant and the barrier addresses / descriptors derived from it are uniform
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* const smem = smem_dyn;
  require_smem_alignment(smem, a.trap, 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + SM_BAR;
  // barrier indices
  const uint32_t W_FULL = bar0, W_EMPTY = bar0 + 8 * TC_NSTAGE_MAX, ACC_FULL = bar0 + 8 * 2 * TC_NSTAGE_MAX,
                 ACC_EMPTY = ACC_FULL + 16, A_READY = ACC_EMPTY + 16;
  float* s_bias = reinterpret_cast<float*>(smem + SM_BIAS);
  float* s_we = reinterpret_cast<float*>(smem + SM_WE);
  float* s_head = reinterpret_cast<float*>(smem + SM_HEAD);
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(W_FULL + 8 * s, 1);
      mbar_init(W_EMPTY + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(ACC_FULL + 8 * b, 1);
      mbar_init(ACC_EMPTY + 8 * b, NW);     // one arrival per worker WARP (per-thread arrivals serialise in the smem atomic unit)
    }
    mbar_init(A_READY, NW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NW) tmem_alloc(sbase + SM_TMEM, TC_TMEM_COLS);
  // sigmoid(z) = 0.5 tanh(z/2) + 0.5: the 1/2 is folded into everything that feeds the i and o gates
  // (their packed weight columns, biases and peephole diagonals), so a gate is one MUFU.TANH + one FMA
  for (int i = tid; i < 384; i += NTHR) s_bias[i] = (i >= 128 && i < 256) ? a.b[i] : 0.5f * a.b[i];
  for (int i = tid; i < 128; i += NTHR) {
    s_bias[384 + i] = 0.5f * a.w_If[i];
    s_bias[512 + i] = 0.5f * a.w_It[i];
    s_bias[640 + i] = 0.5f * a.w_Of[i];
    s_bias[768 + i] = 0.5f * a.w_Ot[i];
  }
  for (int i = tid; i < 256; i += NTHR) s_we[i] = a.W_e[i];
  for (int i = tid; i < 64; i += NTHR) s_we[256 + i] = a.b_e[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == NW) {
    // =============================== weight-stage producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        for (int p = 0; p < TC_NP; ++p)
          for (int kc = 0; kc < NKC; ++kc, ++it) {
            const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
            mbar_wait(W_EMPTY + 8 * s, ph ^ 1u, a.trap, 0x201);
            mbar_arrive_expect_tx(W_FULL + 8 * s, TC_STAGE_BYTES);
            bulk_g2s(sbase + tc_stage_off(s), a.Wp + (size_t)(p * NKC + kc) * TC_STAGE_BYTES,
                     TC_STAGE_BYTES, W_FULL + 8 * s);
          }
      }
    }
  } else if (warp == NW + 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t it = 0, pc = 0, tc = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tc) {
        mbar_wait(A_READY, tc & 1u, a.trap, 0x202);
        tc_fence_after();
        for (int p = 0; p < TC_NP; ++p, ++pc) {
          const uint32_t b = pc & 1u, bph = (pc >> 1) & 1u;
          mbar_wait(ACC_EMPTY + 8 * b, bph ^ 1u, a.trap, 0x203);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + b * TC_ACC_STRIDE;
          for (int kc = 0; kc < NKC; ++kc, ++it) {
            const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
            mbar_wait(W_FULL + 8 * s, ph, a.trap, 0x204);   // written by the async proxy: no tcgen05 fence needed
            const uint64_t db = make_desc_sw128(sbase + tc_stage_off(s));
            if constexpr (X3) {
              // chunks 0-4: W_hi[kc] against A_hi[kc] and A_lo[kc]; chunks 5-9: W_lo[kc] against A_hi[kc]
              const int kk = kc % TC_NKC;
              const uint64_t da_hi = make_desc_sw128(sbase + SM_A + kk * TC_A_BLOCK);
#pragma unroll
              for (int ks = 0; ks < TC_KC / 16; ++ks)
                umma_bf16(d_tmem, da_hi + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), kIdesc, (kc | ks) ? 1u : 0u);
              if (kc < TC_NKC) {
                const uint64_t da_lo = make_desc_sw128(sbase + SM_ALO + kk * TC_A_BLOCK);
#pragma unroll
                for (int ks = 0; ks < TC_KC / 16; ++ks)
                  umma_bf16(d_tmem, da_lo + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), kIdesc, 1u);
              }
            } else {
              const uint64_t da = make_desc_sw128(sbase + SM_A + kc * TC_A_BLOCK);
#pragma unroll
              for (int ks = 0; ks < TC_KC / 16; ++ks)  // advance 32 B (16 bf16) inside the swizzle row
                umma_bf16(d_tmem, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), kIdesc, (kc | ks) ? 1u : 0u);
            }
            umma_commit(W_EMPTY + 8 * s);  // stage reusable once these MMAs have read it
          }
          umma_commit(ACC_FULL + 8 * b);  // accumulator (and, on the last pass, the A operand) done
        }
      }
    }
  } else {
    // =============================== workers: operand build + epilogue ===============================
    const int q = warp & 3, hsel = warp >> 2;
    auto worker_bar = [] { asm volatile("bar.sync 1, %0;" ::"n"(NWT) : "memory"); };
    uint32_t pc = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int row0 = tile * TC_M;
      long long* dbg = (a.dbg && tid == 0) ? a.dbg + ((size_t)blockIdx.x * 16 + (tile / gridDim.x)) * 16 : nullptr;
      if (dbg) dbg[0] = clock64();
      // ---- e = relu(x W_e + b_e) -> block 0.  thread -> (row tid/2, 32 k's)
      {
        constexpr int EPT = 64 * TC_M / NWT;   // k's of e per thread: 32 or 16
        const int r = tid / (64 / EPT), k0 = (tid % (64 / EPT)) * EPT;
        const int gr = row0 + r;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool ok = gr < a.R;
        if (ok) xv = __ldg(reinterpret_cast<const float4*>(a.x) + gr);
#pragma unroll
        for (int kk = 0; kk < EPT; kk += 8) {
          uint32_t pk[4];
          [[maybe_unused]] uint32_t pl[4];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            float e2[2];
#pragma unroll
            for (int z = 0; z < 2; ++z) {
              const int k = k0 + kk + j + z;
              float s = s_we[256 + k];
              s = fmaf(xv.x, s_we[k], s);
              s = fmaf(xv.y, s_we[64 + k], s);
              s = fmaf(xv.z, s_we[128 + k], s);
              s = fmaf(xv.w, s_we[192 + k], s);
              e2[z] = ok ? fmaxf(s, 0.f) : 0.f;
            }
            pk[j >> 1] = pack_op2<F16>(e2[0], e2[1]);
            if constexpr (X3) pl[j >> 1] = pack_bf16x2_lo(e2[0], e2[1], pk[j >> 1]);
          }
          *reinterpret_cast<uint4*>(smem + SM_A + sw128_off(r, k0 + kk)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if constexpr (X3)
            *reinterpret_cast<uint4*>(smem + SM_ALO + sw128_off(r, k0 + kk)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
      }
      const int r = q * 32 + lane;  // epilogue row within tile == TMEM lane
      const int gr = row0 + r;
      const bool rok = gr < a.R;
      const bool v = rok && a.valid[gr] != 0;
      // epilogue operands of sub-chunk idx (pass idx/2, half idx%2): 8 units of c (fp32) and mc
      struct CM { float4 c0, c1; uint4 m; float4 mf0, mf1; };   // m: 8 bf16 (bf16 state); mf0/mf1: fp32 state only
      auto load_cm = [&](int idx) {
        CM o;
        o.c0 = o.c1 = make_float4(0.f, 0.f, 0.f, 0.f);
        o.m = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (!BF) o.mf0 = o.mf1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v) {
          const int u = (idx / NSUB) * TC_UN + hsel * UPT + (idx % NSUB) * 8;
          const size_t so = LAY == 2 ? blk_off(tile, r, u) : (size_t)gr * a.ld + u;
          const float4* cp = reinterpret_cast<const float4*>(a.c + so);
          o.c0 = cp[0];
          o.c1 = cp[1];
          if constexpr (BF) {
            o.m = *reinterpret_cast<const uint4*>(a.mcb + so);
          } else {
            const float4* mp = reinterpret_cast<const float4*>(a.mc + (size_t)gr * a.ld + u);
            o.mf0 = mp[0];
            o.mf1 = mp[1];
          }
        }
        return o;
      };
      // ---- h -> blocks 1,2 ; mh -> blocks 3,4
      if constexpr (LAY == 2) {
        // blocked bf16 state: piece (g, row) is 16 B; thread t takes pieces t + 256k -> consecutive lanes read
        // consecutive rows of one group: 512 B contiguous per warp instruction; 8 loads in flight per array
        const uint4* hsrc = reinterpret_cast<const uint4*>(a.hb + (size_t)tile * TC_M * TC_U);
        const uint4* msrc = reinterpret_cast<const uint4*>(a.mhb + (size_t)tile * TC_M * TC_U);
#pragma unroll
        for (int arr = 0; arr < 2; ++arr) {
          const uint4* src = arr ? msrc : hsrc;
          uint4 v8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v8[k] = src[tid + 256 * k];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int pidx = tid + 256 * k, g = pidx >> 7, rr = pidx & 127;
            const uint32_t off = (uint32_t)rr * 128u + (uint32_t)(((g & 7) ^ (rr & 7)) << 4);
            *reinterpret_cast<uint4*>(smem + SM_A + (1 + 2 * arr + (g >> 3)) * TC_A_BLOCK + off) = v8[k];
          }
        }
      } else if constexpr (LAY == 1) {
        // bf16 state: a warp owns 16 rows; lane -> (row parity, 16-byte unit); 8 independent 16 B loads in flight
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint4 hv[4], mv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = warp * 16 + half * 8 + i * 2 + (lane >> 4);
            const int g2 = row0 + rr;
            hv[i] = mv[i] = make_uint4(0u, 0u, 0u, 0u);
            if (g2 < a.R) {
              hv[i] = *reinterpret_cast<const uint4*>(a.hb + (size_t)g2 * TC_U + (lane & 15) * 8);
              mv[i] = *reinterpret_cast<const uint4*>(a.mhb + (size_t)g2 * TC_U + (lane & 15) * 8);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = warp * 16 + half * 8 + i * 2 + (lane >> 4);
            const int un = lane & 15, blk = un >> 3;
            const uint32_t off = (uint32_t)rr * 128u + (uint32_t)(((un & 7) ^ (rr & 7)) << 4);
            *reinterpret_cast<uint4*>(smem + SM_A + (1 + blk) * TC_A_BLOCK + off) = hv[i];
            *reinterpret_cast<uint4*>(smem + SM_A + (3 + blk) * TC_A_BLOCK + off) = mv[i];
          }
        }
      } else {
        // fp32 state: one warp per row, lane -> 4 consecutive k, converted to bf16 on the way.  Eight rows (16 independent
        // 512-byte loads) are in flight per warp: one row at a time left the build phase latency-bound (with the
        // split-bf16 kernel's one CTA per SM nothing else hides it: 24 k of its 77 k clk per tile)
        constexpr int RB = 4;   // rows in flight per warp (96 / 112 registers per thread)
#pragma unroll 1
        for (int r8 = warp; r8 < TC_M; r8 += NW * RB) {
          float4 hv[RB], mv[RB];
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int g2 = row0 + r8 + NW * i;
            hv[i] = mv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g2 < a.R) {
              hv[i] = *reinterpret_cast<const float4*>(a.h + (size_t)g2 * a.ld + lane * 4);
              mv[i] = *reinterpret_cast<const float4*>(a.mh + (size_t)g2 * a.ld + lane * 4);
            }
          }
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int rr = r8 + NW * i;
            const int k = lane * 4;  // 0..124 within the 128-wide part
            const int blk = k >> 6, kk = k & 63;
            const uint2 hh = make_uint2(pack_op2<F16>(hv[i].x, hv[i].y), pack_op2<F16>(hv[i].z, hv[i].w));
            const uint2 mm = make_uint2(pack_op2<F16>(mv[i].x, mv[i].y), pack_op2<F16>(mv[i].z, mv[i].w));
            *reinterpret_cast<uint2*>(smem + SM_A + (1 + blk) * TC_A_BLOCK + sw128_off(rr, kk)) = hh;
            *reinterpret_cast<uint2*>(smem + SM_A + (3 + blk) * TC_A_BLOCK + sw128_off(rr, kk)) = mm;
            if constexpr (X3) {
              *reinterpret_cast<uint2*>(smem + SM_ALO + (1 + blk) * TC_A_BLOCK + sw128_off(rr, kk)) =
                  make_uint2(pack_bf16x2_lo(hv[i].x, hv[i].y, hh.x), pack_bf16x2_lo(hv[i].z, hv[i].w, hh.y));
              *reinterpret_cast<uint2*>(smem + SM_ALO + (3 + blk) * TC_A_BLOCK + sw128_off(rr, kk)) =
                  make_uint2(pack_bf16x2_lo(mv[i].x, mv[i].y, mm.x), pack_bf16x2_lo(mv[i].z, mv[i].w, mm.y));
            }
          }
        }
      }
      if (dbg) dbg[1] = clock64();
      CM pre = {}, pre2 = {};
      if constexpr (!STG) {
        pre = load_cm(0);
        pre2 = load_cm(1);   // two sub-chunks in flight while the MMAs of pass 0 run
      }
      // split-bf16 kernel: operands of the epilogue staged through shared memory, pass by pass (see TC_STG_Q)
      constexpr int NSTG = TC_M * TC_UN / 4 / NWT;   // float4 per thread and array in the staging copies: 4 or 2
      float4* const s_c4 = reinterpret_cast<float4*>(smem + SM_STG);
      float4* const s_m4 = reinterpret_cast<float4*>(smem + SM_STG + TC_STG_BYTES);
      [[maybe_unused]] auto stage_load = [&](int pass, float4 (&cin)[NSTG], float4 (&min)[NSTG]) {   // global -> registers (coalesced)
#pragma unroll
        for (int k = 0; k < NSTG; ++k) {
          const int i = tid + NWT * k, row = i >> 3, chunk = i & 7;
          const int g2 = row0 + row;
          cin[k] = min[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g2 < a.R) {
            cin[k] = *reinterpret_cast<const float4*>(a.c + (size_t)g2 * a.ld + pass * TC_UN + chunk * 4);
            min[k] = *reinterpret_cast<const float4*>(a.mc + (size_t)g2 * a.ld + pass * TC_UN + chunk * 4);
          }
        }
      };
      [[maybe_unused]] auto stage_put = [&](const float4 (&cin)[NSTG], const float4 (&min)[NSTG]) {   // registers -> staging
#pragma unroll
        for (int k = 0; k < NSTG; ++k) {
          const int i = tid + NWT * k, row = i >> 3, chunk = i & 7;
          s_c4[chunk * TC_STG_Q + row] = cin[k];
          s_m4[chunk * TC_STG_Q + row] = min[k];
        }
      };
      if constexpr (STG) {
        float4 cin[NSTG], min[NSTG];
        stage_load(0, cin, min);
        stage_put(cin, min);   // visible to the epilogue threads after the bar.sync that follows the fence below
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      __syncwarp();         // every lane has fenced its own writes; lane 0's arrive releases them
      if (lane == 0) mbar_arrive(A_READY);
      if constexpr (STG) worker_bar();   // pass 0's c / mc staged

      // ---- epilogue: 8 sub-chunks (4 passes x 2 halves), operands prefetched one sub-chunk ahead
      float y[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      [[maybe_unused]] float4 nxt_c[NSTG], nxt_m[NSTG];   // split-bf16 kernel: next pass's c / mc, in flight during this pass's epilogue
#pragma unroll 1
      for (int idx = 0; idx < NSUB * TC_NP; ++idx) {
        const int p = idx / NSUB, sub = idx % NSUB;
        const uint32_t b = pc & 1u, bph = (pc >> 1) & 1u;
        if constexpr (STG) {
          if (sub == 0 && p + 1 < TC_NP) stage_load(p + 1, nxt_c, nxt_m);
        }
        if (sub == 0) {
          if (dbg) dbg[2 + 3 * p] = clock64();
          mbar_wait(ACC_FULL + 8 * b, bph, a.trap, 0x205);
          if (dbg) dbg[3 + 3 * p] = clock64();
          tc_fence_after();
        }
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + b * TC_ACC_STRIDE;
        const int ul = hsel * UPT + sub * 8;  // unit within pass
        const int u = p * TC_UN + ul;        // global unit
        float zi[8], zj[8], zo[8];
        tmem_ld8(t_row + ul, zi);
        tmem_ld8(t_row + TC_UN + ul, zj);
        tmem_ld8(t_row + 2 * TC_UN + ul, zo);
        CM cur;
        if constexpr (STG) {
          const int qd = ul >> 2;   // unit quad within the pass
          cur.c0 = s_c4[qd * TC_STG_Q + r];
          cur.c1 = s_c4[(qd + 1) * TC_STG_Q + r];
          cur.mf0 = s_m4[qd * TC_STG_Q + r];
          cur.mf1 = s_m4[(qd + 1) * TC_STG_Q + r];
          cur.m = make_uint4(0u, 0u, 0u, 0u);
        } else {
          cur = pre;
          pre = pre2;
          if (idx + 2 < NSUB * TC_NP) pre2 = load_cm(idx + 2);
        }
        tmem_wait_ld();
        float ho[8], co[8], fo[8];
        if (v) {
          const float2 kHalf = make_float2(0.5f, 0.5f), kNeg = make_float2(-1.f, -1.f);
          auto tanh2 = [](float2 t) {   // x3: accurate tanhf; bf16: tanh.approx (error below the operand rounding)
#ifdef TC_X3_EXP_FAST_TANH   // timing experiment only (scratch/timeline_x3.py): what the accurate tanh costs
            if constexpr (X3) return mmt::tanh2(t);
#endif
            if constexpr (X3) return make_float2(tanh_acc(t.x), tanh_acc(t.y));
            else return mmt::tanh2(t);
          };
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {   // 4 units at a time: parameters come as 128-bit smem loads
            const int uu = u + hq * 4;
            const float4 bI = *reinterpret_cast<const float4*>(s_bias + uu);
            const float4 bJ = *reinterpret_cast<const float4*>(s_bias + 128 + uu);
            const float4 bO = *reinterpret_cast<const float4*>(s_bias + 256 + uu);
            const float4 pIf = *reinterpret_cast<const float4*>(s_bias + 384 + uu);
            const float4 pIt = *reinterpret_cast<const float4*>(s_bias + 512 + uu);
            const float4 pOf = *reinterpret_cast<const float4*>(s_bias + 640 + uu);
            const float4 pOt = *reinterpret_cast<const float4*>(s_bias + 768 + uu);
            const float4 c4 = hq ? cur.c1 : cur.c0;
            float4 m4;
            if constexpr (BF) {
              const uint32_t w0 = hq ? cur.m.z : cur.m.x, w1 = hq ? cur.m.w : cur.m.y;
              m4 = make_float4(op_lo<F16>(w0), op_hi<F16>(w0), op_lo<F16>(w1), op_hi<F16>(w1));
            } else {
              m4 = hq ? cur.mf1 : cur.mf0;
            }
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {  // a pair of units per packed FFMA2
              const int i0 = hq * 4 + pr * 2;
              const float2 c2 = pr ? make_float2(c4.z, c4.w) : make_float2(c4.x, c4.y);
              const float2 m2 = pr ? make_float2(m4.z, m4.w) : make_float2(m4.x, m4.y);
              auto sel = [&](const float4& f) { return pr ? make_float2(f.z, f.w) : make_float2(f.x, f.y); };
              float2 t = fadd2(make_float2(zi[i0], zi[i0 + 1]), sel(bI));
              t = ffma2(sel(pIf), m2, t);
              t = ffma2(sel(pIt), c2, t);
              const float2 g = ffma2(tanh2(t), kHalf, kHalf);
              const float2 tj = tanh2(fadd2(make_float2(zj[i0], zj[i0 + 1]), sel(bJ)));
              const float2 cf = ffma2(g, ffma2(m2, kNeg, tj), m2);   // (1-g)*mc + g*tj
              const float2 ct = ffma2(g, ffma2(c2, kNeg, tj), c2);
              float2 o = fadd2(make_float2(zo[i0], zo[i0 + 1]), sel(bO));
              o = ffma2(sel(pOf), cf, o);
              o = ffma2(sel(pOt), ct, o);
              const float2 qq = ffma2(tanh2(o), kHalf, kHalf);
              const float2 f2 = fmul2(qq, tanh2(cf)), h2 = fmul2(qq, tanh2(ct));
              fo[i0] = f2.x; fo[i0 + 1] = f2.y;
              ho[i0] = h2.x; ho[i0 + 1] = h2.y;
              co[i0] = ct.x; co[i0 + 1] = ct.y;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) ho[i] = co[i] = fo[i] = 0.f;
        }
        if constexpr (STG) {
          // c' and h' take the staging slots their inputs came from (this thread's own); stored coalesced after the pass
          const int qd = ul >> 2;
          s_c4[qd * TC_STG_Q + r] = make_float4(co[0], co[1], co[2], co[3]);
          s_c4[(qd + 1) * TC_STG_Q + r] = make_float4(co[4], co[5], co[6], co[7]);
          s_m4[qd * TC_STG_Q + r] = make_float4(ho[0], ho[1], ho[2], ho[3]);
          s_m4[(qd + 1) * TC_STG_Q + r] = make_float4(ho[4], ho[5], ho[6], ho[7]);
          if (a.mf_out && rok) {
            float4* fp = reinterpret_cast<float4*>(a.mf_out + (size_t)gr * a.ld_mf + u);
            fp[0] = make_float4(fo[0], fo[1], fo[2], fo[3]);
            fp[1] = make_float4(fo[4], fo[5], fo[6], fo[7]);
          }
        } else if (rok || LAY == 2) {
          const size_t so = LAY == 2 ? blk_off(tile, r, u) : (size_t)gr * a.ld + u;
          float4* cp = reinterpret_cast<float4*>(a.c_out + so);
          cp[0] = make_float4(co[0], co[1], co[2], co[3]);
          cp[1] = make_float4(co[4], co[5], co[6], co[7]);
          if constexpr (BF) {
            *reinterpret_cast<uint4*>(a.hb_out + so) =
                make_uint4(pack_op2<F16>(ho[0], ho[1]), pack_op2<F16>(ho[2], ho[3]), pack_op2<F16>(ho[4], ho[5]),
                           pack_op2<F16>(ho[6], ho[7]));
          } else {
            float4* hp = reinterpret_cast<float4*>(a.h_out + (size_t)gr * a.ld + u);
            hp[0] = make_float4(ho[0], ho[1], ho[2], ho[3]);
            hp[1] = make_float4(ho[4], ho[5], ho[6], ho[7]);
            if (a.mf_out) {
              float4* fp = reinterpret_cast<float4*>(a.mf_out + (size_t)gr * a.ld_mf + u);
              fp[0] = make_float4(fo[0], fo[1], fo[2], fo[3]);
              fp[1] = make_float4(fo[4], fo[5], fo[6], fo[7]);
            }
          }
        }
#ifdef TC_X3_EXP_NO_HEAD      // timing experiment only
        if (false) {
#else
        if (a.params_out) {
#endif
          // head partial sums: W_h rows u..u+7 are 40 contiguous floats (160 B, 16-byte aligned): 10 uniform
          // 128-bit loads per half instead of 40 scalar ones (the scalar version saturated L1TEX)
          float wv[40];
#pragma unroll
          for (int hsrc = 0; hsrc < 2; ++hsrc) {
            const float4* wp = reinterpret_cast<const float4*>(a.W_h + (size_t)(hsrc * TC_U + u) * 5);
#pragma unroll
            for (int k4 = 0; k4 < 10; ++k4) {
              const float4 t4 = __ldg(wp + k4);
              wv[4 * k4] = t4.x; wv[4 * k4 + 1] = t4.y; wv[4 * k4 + 2] = t4.z; wv[4 * k4 + 3] = t4.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int z = 0; z < 5; ++z) y[z] = fmaf(hsrc ? fo[i] : ho[i], wv[i * 5 + z], y[z]);
          }
        }
        if (sub == NSUB - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ACC_EMPTY + 8 * b);
          if constexpr (STG) {
            worker_bar();           // every h', c' of the pass is in the staging slots
            float4 cout[NSTG], hout[NSTG];
#pragma unroll
            for (int k = 0; k < NSTG; ++k) {
              const int i = tid + NWT * k, row = i >> 3, chunk = i & 7;
              cout[k] = s_c4[chunk * TC_STG_Q + row];
              hout[k] = s_m4[chunk * TC_STG_Q + row];
            }
            worker_bar();           // all staging slots have been read (also guards the next tile)
            if (p + 1 < TC_NP) stage_put(nxt_c, nxt_m);               // next pass's operands (loaded during this pass's epilogue)
#pragma unroll
            for (int k = 0; k < NSTG; ++k) {
              const int i = tid + NWT * k, row = i >> 3, chunk = i & 7;
              const int g2 = row0 + row;
              if (g2 < a.R) {
                *reinterpret_cast<float4*>(a.c_out + (size_t)g2 * a.ld + p * TC_UN + chunk * 4) = cout[k];
                *reinterpret_cast<float4*>(a.h_out + (size_t)g2 * a.ld + p * TC_UN + chunk * 4) = hout[k];
              }
            }
            if (p + 1 < TC_NP) worker_bar();         // staged for pass p + 1
          }
          if (dbg) dbg[4 + 3 * p] = clock64();
          ++pc;
        }
      }
      if (dbg) dbg[14] = clock64();
      // ---- head: combine the two column halves of each row
      if (a.params_out) {
        // slices 1 .. NH-1 hand their partial sums to slice 0.  The split-bf16 kernel has three of them: they go through the
        // A_lo operand region, free once the last pass has completed (every worker has seen its ACC_FULL) until the next build
        float* const s_part = X3 ? reinterpret_cast<float*>(smem + SM_ALO) : s_head;
        if (hsel >= 1) {
#pragma unroll
          for (int z = 0; z < 5; ++z) s_part[((hsel - 1) * TC_M + r) * 5 + z] = y[z];
        }
        worker_bar();
        if (hsel == 0 && rok) {
          float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
          if (v) {
#pragma unroll
            for (int z = 0; z < 5; ++z) {
              float acc = y[z];
#pragma unroll
              for (int hh = 0; hh < NH - 1; ++hh) acc += s_part[(hh * TC_M + r) * 5 + z];
              o[z] = acc + __ldg(a.b_h + z);
            }
            o[2] = X3 ? expf(o[2]) : __expf(o[2]);
            o[3] = X3 ? expf(o[3]) : __expf(o[3]);
            o[4] = X3 ? tanh_acc(o[4]) : tanh_fast(o[4]);
          }
          float* po = a.params_out + (size_t)gr * a.params_stride;
#pragma unroll
          for (int z = 0; z < 5; ++z) po[z] = o[z];
          if (a.next_pos) {
            const float2 cp = *reinterpret_cast<const float2*>(a.cur_pos + (size_t)gr * 2);
            *reinterpret_cast<float2*>(a.next_pos + (size_t)gr * 2) = make_float2(cp.x + o[0], cp.y + o[1]);
          }
        }
        worker_bar();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// W[E+2U, 3U] fp32 row-major -> bf16 operand image [pass][k-chunk][96 rows][64 k], SWIZZLE_128B.
// Row n = g*32 + ul of a pass holds gate column g*U + pass*32 + ul.
template <bool F16>
__global__ void pack_gate_weights_kernel(const float* __restrict__ W, uint8_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TC_K * 3 * TC_U) return;
  const int k = idx / (3 * TC_U), col = idx - k * (3 * TC_U);
  const int g = col / TC_U, u = col - g * TC_U;
  const int p = u / TC_UN, ul = u - p * TC_UN;
  const int n = g * TC_UN + ul;
  const int kc = k / TC_KC, kk = k - kc * TC_KC;
  const size_t off = (size_t)(p * TC_NKC + kc) * TC_STAGE_BYTES + sw128_off(n, kk);
  // gates i (g == 0) and o (g == 2) go through sigmoid(z) = 0.5 tanh(z/2) + 0.5: fold the 1/2 (exact in bf16)
  const float w = g == 1 ? W[idx] : 0.5f * W[idx];
  if constexpr (F16) *reinterpret_cast<__half*>(out + off) = __float2half_rn(w);
  else *reinterpret_cast<__nv_bfloat16*>(out + off) = __float2bfloat16_rn(w);
}

// split-bf16 image: [pass][10 chunks][96 rows][64 k]: chunks 0-4 hold W_hi (against A_hi and A_lo), 5-9 W_lo (against A_hi)
__global__ void pack_gate_weights_x3_kernel(const float* __restrict__ W, uint8_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TC_K * 3 * TC_U) return;
  const int k = idx / (3 * TC_U), col = idx - k * (3 * TC_U);
  const int g = col / TC_U, u = col - g * TC_U;
  const int p = u / TC_UN, ul = u - p * TC_UN;
  const int n = g * TC_UN + ul;
  const int kc = k / TC_KC, kk = k - kc * TC_KC;
  const float w = g == 1 ? W[idx] : 0.5f * W[idx];           // sigmoid as 0.5 tanh(z/2) + 0.5: the 1/2 is exact
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const size_t base = (size_t)(p * TC_X3_NKC) * TC_STAGE_BYTES + sw128_off(n, kk);
  *reinterpret_cast<__nv_bfloat16*>(out + base + (size_t)kc * TC_STAGE_BYTES) = hi;
  *reinterpret_cast<__nv_bfloat16*>(out + base + (size_t)(TC_NKC + kc) * TC_STAGE_BYTES) = lo;
}

static void tc_fill_weights(TcArgs& a, const mmt_cell_weights* w) {
  a.W_e = w->W_e; a.b_e = w->b_e; a.b = w->b; a.w_If = w->w_If; a.w_It = w->w_It; a.w_Of = w->w_Of; a.w_Ot = w->w_Ot;
  a.W_h = w->W_h; a.b_h = w->b_h; a.Wp = reinterpret_cast<const uint8_t*>(w->W_packed_bf16);
}

template <int LAY, bool X3 = false, bool F16 = false>
static int tc_launch(TcArgs& a, cudaStream_t stream) {
  a.num_tiles = (a.R + TC_M - 1) / TC_M;
  a.trap = trap_record();
  constexpr int kSmem = (X3 ? SM_TOTAL_X3 : SM_TOTAL) + 1024, kPerSM = X3 ? 1 : 2;
  static DeviceMask smem_opted[1];   // per kernel: devices already opted in
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&gsk_cell_tc_kernel<LAY, X3, F16>), kSmem, &smem_opted[0])) return rc;
  const int grid = a.num_tiles < kPerSM * num_sms() ? a.num_tiles : kPerSM * num_sms();
  gsk_cell_tc_kernel<LAY, X3, F16><<<grid, X3 ? TC_THREADS_X3 : TC_THREADS, kSmem, stream>>>(a);
  count_launch();
  return check_launch(X3 ? "gsk_cell_tc_kernel<x3>" : "gsk_cell_tc_kernel");
}

// bf16-state variant used by the rollout: h, mh, mc as bf16 [R,U]; c fp32 [R,U]
int launch_cell_tc_bf16(const float* x, const void* hb, const float* c, const void* mhb, const void* mcb,
                        const uint8_t* valid, const mmt_cell_weights* w, int R, void* hb_out, float* c_out,
                        const float* cur_pos, float* params_out, int params_stride, float* next_pos, int blocked, int f16,
                        cudaStream_t stream) {
  TcArgs a = {};
  a.x = x; a.c = c; a.valid = valid;
  a.hb = reinterpret_cast<const __nv_bfloat16*>(hb);
  a.mhb = reinterpret_cast<const __nv_bfloat16*>(mhb);
  a.mcb = reinterpret_cast<const __nv_bfloat16*>(mcb);
  a.hb_out = reinterpret_cast<__nv_bfloat16*>(hb_out);
  tc_fill_weights(a, w);
  a.c_out = c_out; a.cur_pos = cur_pos; a.params_out = params_out; a.next_pos = next_pos;
  a.R = R; a.ld = TC_U; a.ld_mf = TC_U; a.params_stride = params_stride;
  if (f16) {
    a.Wp = reinterpret_cast<const uint8_t*>(w->W_packed_f16);
    return blocked ? tc_launch<2, false, true>(a, stream) : tc_launch<1, false, true>(a, stream);
  }
  return blocked ? tc_launch<2>(a, stream) : tc_launch<1>(a, stream);
}

int launch_cell_tc(const float* x, const float* h, const float* c, const float* mh, const float* mc, int ld,
                   const uint8_t* valid, const mmt_cell_weights* w, int R, float* h_out, float* c_out, float* mf_out,
                   int ld_mf, const float* cur_pos, float* params_out, int params_stride, float* next_pos,
                   int x3 /* 0: bf16 operands, 1: split bf16, 2: fp16 operands */, cudaStream_t stream) {
  TcArgs a = {};
  a.x = x; a.h = h; a.c = c; a.mh = mh; a.mc = mc; a.valid = valid;
  tc_fill_weights(a, w);
  if (x3 == 1) a.Wp = reinterpret_cast<const uint8_t*>(w->W_packed_bf16x3);
  if (x3 == 2) a.Wp = reinterpret_cast<const uint8_t*>(w->W_packed_f16);
  a.h_out = h_out; a.c_out = c_out; a.mf_out = mf_out; a.cur_pos = cur_pos; a.params_out = params_out;
  a.next_pos = next_pos; a.R = R; a.ld = ld; a.ld_mf = ld_mf; a.params_stride = params_stride;
  return x3 == 1 ? tc_launch<0, true>(a, stream) : x3 == 2 ? tc_launch<0, false, true>(a, stream) : tc_launch<0>(a, stream);
}

}  // namespace mmt

extern "C" size_t mmt_gate_weights_packed_x3_bytes(int E, int U) {
  if (E != mmt::TC_E || U != mmt::TC_U) return 0;
  return (size_t)mmt::TC_NP * mmt::TC_X3_NKC * mmt::TC_STAGE_BYTES;
}

extern "C" int mmt_pack_gate_weights_bf16x3(const float* W, int E, int U, void* packed, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(W && packed, "W/packed must not be NULL");
  MMT_REQUIRE(E == TC_E && U == TC_U, "packing is built for E = 64, U = 128");
  MMT_ALIGNED(packed);
  const int n = TC_K * 3 * TC_U;
  pack_gate_weights_x3_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, reinterpret_cast<uint8_t*>(packed));
  count_launch();
  return check_launch("pack_gate_weights_x3_kernel");
}

extern "C" size_t mmt_gate_weights_packed_bytes(int E, int U) {
  if (E != mmt::TC_E || U != mmt::TC_U) return 0;
  return (size_t)mmt::TC_NP * mmt::TC_NKC * mmt::TC_STAGE_BYTES;
}

extern "C" int mmt_pack_gate_weights_bf16(const float* W, int E, int U, void* packed, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(W && packed, "W/packed must not be NULL");
  MMT_REQUIRE(E == TC_E && U == TC_U, "packing is built for E = 64, U = 128");
  MMT_ALIGNED(packed);
  const int n = TC_K * 3 * TC_U;
  pack_gate_weights_kernel<false><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, reinterpret_cast<uint8_t*>(packed));
  count_launch();
  return check_launch("pack_gate_weights_kernel");
}

// the same image with fp16 entries (MMT_PREC_F16: fused rollout with fp16 operands); same size
extern "C" int mmt_pack_gate_weights_f16(const float* W, int E, int U, void* packed, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(W && packed, "W/packed must not be NULL");
  MMT_REQUIRE(E == TC_E && U == TC_U, "packing is built for E = 64, U = 128");
  MMT_ALIGNED(packed);
  const int n = TC_K * 3 * TC_U;
  pack_gate_weights_kernel<true><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, reinterpret_cast<uint8_t*>(packed));
  count_launch();
  return check_launch("pack_gate_weights_kernel");
}

// diagnostics: run the blocked bf16-state kernel once with per-tile phase timestamps (clock64 of worker 0)
extern "C" int mmt_debug_cell_tc_timeline(const float* x, const void* hb, const float* c, const void* mhb,
                                          const void* mcb, const uint8_t* valid, const mmt_cell_weights* w, int R,
                                          void* hb_out, float* c_out, const float* cur_pos, float* params_out,
                                          float* next_pos, long long* dbg, void* stream) {
  using namespace mmt;
  TcArgs a = {};
  a.x = x; a.c = c; a.valid = valid;
  a.hb = reinterpret_cast<const __nv_bfloat16*>(hb);
  a.mhb = reinterpret_cast<const __nv_bfloat16*>(mhb);
  a.mcb = reinterpret_cast<const __nv_bfloat16*>(mcb);
  a.hb_out = reinterpret_cast<__nv_bfloat16*>(hb_out);
  tc_fill_weights(a, w);
  a.c_out = c_out; a.cur_pos = cur_pos; a.params_out = params_out; a.next_pos = next_pos;
  a.R = R; a.ld = TC_U; a.ld_mf = TC_U; a.params_stride = 5; a.dbg = dbg;
  return tc_launch<2>(a, (cudaStream_t)stream);
}

// diagnostics: the fp32-state kernel (bf16 or split-bf16 gate GEMM) once with per-tile phase timestamps
extern "C" int mmt_debug_cell_tc_f32state_timeline(const float* x, const float* h, const float* c, const float* mh,
                                                   const float* mc, const uint8_t* valid, const mmt_cell_weights* w, int R,
                                                   float* h_out, float* c_out, const float* cur_pos, float* params_out,
                                                   float* next_pos, int x3, long long* dbg, void* stream) {
  using namespace mmt;
  TcArgs a = {};
  a.x = x; a.h = h; a.c = c; a.mh = mh; a.mc = mc; a.valid = valid;
  tc_fill_weights(a, w);
  if (x3) a.Wp = reinterpret_cast<const uint8_t*>(w->W_packed_bf16x3);
  a.h_out = h_out; a.c_out = c_out; a.mf_out = nullptr; a.cur_pos = cur_pos; a.params_out = params_out;
  a.next_pos = next_pos; a.R = R; a.ld = TC_U; a.ld_mf = TC_U; a.params_stride = 5; a.dbg = dbg;
  return x3 ? tc_launch<0, true>(a, (cudaStream_t)stream) : tc_launch<0>(a, (cudaStream_t)stream);
}
