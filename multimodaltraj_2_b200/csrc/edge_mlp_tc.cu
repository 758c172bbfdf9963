// Relational edge MLP of g2k_lstm_mcr on the 5th-generation tensor cores (bf16 operands, fp32 accumulation in
// TMEM): SURVEY App. C.3 / section 8(f) rank 1; include/mmt.h mmt_edge_mlp_bf16.
//
//   [a | b] = h [W1a | W1b]                   node level:  node_proj_tc_kernel, one M128 x N256 x K128 GEMM per 128 rows
//   e1_ij = elu(a_i + b_j + b1)               edge level:  edge_mlp_tc_kernel, per tile of 128 EDGES of the adjacency
//   e2_ij = elu(e1_ij W2 + b2)                             mask: gather -> e1 tile (bf16, SWIZZLE_128B) -> 8 MMAs
//   score_ij = sigmoid(w_out . e2_ij + b_out)              M128 x N128 x K16 -> epilogue from TMEM -> scatter
// The crowd graphs are sparse (~3 neighbours per agent): only edges are evaluated, and an edge costs 2 x 512 B of
// gathered node projections (L2-resident per scene) + 32 kFLOP on the tensor pipe; the fp32 CUDA-core version
// (edge_mlp.cu, parity mode) spent 3.7 ms per step on C3, 77 % of the g2k_lstm_mcr step.
#include <cuda_bf16.h>

#include "mmt_common.cuh"
#include "tc_common.cuh"

namespace mmt {

constexpr int EM_U = 128, EM_HE = 128;
constexpr int EM_W1_BYTES = 2 * 256 * 128;     // [k-block 2][256 out rows][128 B]
constexpr int EM_W2_BYTES = 2 * 128 * 128;     // [k-block 2][128 out rows][128 B]
constexpr int EM_BLK = 128 * 128;              // A operand k-block: [128 rows][128 B]

// ------------------------------------------------------------------------------------------------
// packing: W1[2U,He], W2[He,He] fp32 row-major -> bf16 K-major SWIZZLE_128B operand images
template <bool F16>
__global__ void pack_edge_weights_kernel(const float* __restrict__ W1, const float* __restrict__ W2,
                                         uint8_t* __restrict__ out) {
  auto put = [](uint8_t* p, float w) {
    if constexpr (F16) *reinterpret_cast<__half*>(p) = __float2half_rn(w);
    else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(w);
  };
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 256 * 128) {
    const int n = idx >> 7, k = idx & 127;   // output column n of [a | b], input unit k
    const float w = n < 128 ? W1[(size_t)k * EM_HE + n] : W1[(size_t)(EM_U + k) * EM_HE + (n - 128)];
    put(out + (k >> 6) * (256 * 128) + sw128_off(n, k & 63), w);
  } else if (idx < 256 * 128 + 128 * 128) {
    const int i2 = idx - 256 * 128;
    const int n = i2 >> 7, k = i2 & 127;
    put(out + EM_W1_BYTES + (k >> 6) * EM_BLK + sw128_off(n, k & 63), W2[(size_t)k * EM_HE + n]);
  }
}

// ------------------------------------------------------------------------------------------------
// node projections: out[R, 256] = h[R, 128 (ld)] x [W1a | W1b]   (bf16 operands, fp32 accumulation, bf16 result: the edge
// kernel gathers two 256-byte rows per edge instead of two 512-byte ones and keeps eight edges in flight per warp)
constexpr int NP_SM_B = 0;                          // 64 KB
constexpr int NP_SM_A = NP_SM_B + EM_W1_BYTES;      // 32 KB
constexpr int NP_SM_BAR = NP_SM_A + 34 * 1024;      // A block (32 KB); the epilogue staging [8][32][33] floats needs 33 KB
constexpr int NP_SM_TOTAL = NP_SM_BAR + 32;
template <bool F16> constexpr uint32_t kIdescNP = make_idesc_op<F16>(128, 256);

// BLOCKED: h is the tile-blocked bf16 state of the per-step fast path (cell_tc.cu: piece (g, r) = units 8g..8g+7 of
// row r of a 128-row tile, 16 bytes at ((tile 16 + g) 128 + r) 16) -- each piece is one 16-byte chunk of the K-major
// SWIZZLE_128B operand, copied as it is.
template <bool BLOCKED, bool F16 = false>
__global__ void __launch_bounds__(256, 2) node_proj_tc_kernel(const float* __restrict__ h, int ld_h, int R,
                                                              const uint8_t* __restrict__ Wp, __nv_bfloat16* __restrict__ out,
                                                              int num_tiles, uint32_t* trap) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];   // link-time constant base: uniform addresses / descriptors
  uint8_t* const smem = smem_dyn;
  require_smem_alignment(smem, trap, 4);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_w = sbase + NP_SM_BAR, bar_mma = bar_w + 8;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + NP_SM_BAR + 16);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive_expect_tx(bar_w, EM_W1_BYTES);
    bulk_g2s(sbase + NP_SM_B, Wp, EM_W1_BYTES, bar_w);
  }
  if (warp == 0) tmem_alloc(sbase + NP_SM_BAR + 16, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  mbar_wait(bar_w, 0, trap, 0x401);

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int row0 = tile * 128;
    if constexpr (BLOCKED) {
      const uint4* src = reinterpret_cast<const uint4*>(h) + (size_t)tile * 2048;   // 16 pieces x 128 rows
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int pc = tid + 256 * k, g = pc >> 7, rr = pc & 127;
        *reinterpret_cast<uint4*>(smem + NP_SM_A + (g >> 3) * EM_BLK + sw128_off(rr, (g & 7) * 8)) = __ldg(src + pc);
      }
    } else
    // A operand: one warp per row, lane -> 4 consecutive k (fp32 -> bf16), K-major SWIZZLE_128B
    for (int rr = warp; rr < 128; rr += 8) {
      const int g = row0 + rr;
      float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < R) hv = __ldg(reinterpret_cast<const float4*>(h + (size_t)g * ld_h) + lane);
      const int k = lane * 4;
      *reinterpret_cast<uint2*>(smem + NP_SM_A + (k >> 6) * EM_BLK + sw128_off(rr, k & 63)) =
          make_uint2(pack_op2<F16>(hv.x, hv.y), pack_op2<F16>(hv.z, hv.w));
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t da = make_desc_sw128(sbase + NP_SM_A + (ks >> 2) * EM_BLK) + (uint64_t)((ks & 3) * 2);
        const uint64_t db = make_desc_sw128(sbase + NP_SM_B + (ks >> 2) * (256 * 128)) + (uint64_t)((ks & 3) * 2);
        umma_bf16(tmem_base, da, db, kIdescNP<F16>, ks ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, it & 1u, trap, 0x402);
    tc_fence_after();
    // epilogue: warp w -> rows 32 (w % 4) .., columns 128 (w / 4) ..  TMEM gives a lane one ROW; stored directly that
    // is 32 rows x 16 B per instruction.  Each warp instead transposes 32 x 32 blocks through its 4 KB slice of the
    // (now free) A block, so that every global store instruction writes 32 consecutive floats of one row.
    {
      const int q = warp & 3, half = warp >> 2;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + half * 128;
      float* stg = reinterpret_cast<float*>(smem + NP_SM_A) + warp * (32 * 33);   // [32][33] padded
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        float v[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_ld8(t_row + cb * 32 + j * 8, v[j]);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) stg[lane * 33 + j * 8 + i] = v[j][i];
        __syncwarp();
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          const int g = row0 + q * 32 + rr;
          if (g < R) {
            const float pv = stg[rr * 33 + lane];
            if constexpr (F16) reinterpret_cast<__half*>(out)[(size_t)g * 256 + half * 128 + cb * 32 + lane] = __float2half_rn(pv);
            else out[(size_t)g * 256 + half * 128 + cb * 32 + lane] = __float2bfloat16_rn(pv);
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM drained and the A block free before the next tile
  }
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------
// edge level
constexpr int EE_LIST = 4096;                       // candidate pairs scanned per chunk (>= edges found)
constexpr int EE_SM_W2 = 0;                         // 32 KB
constexpr int EE_SM_A = EE_SM_W2 + EM_W2_BYTES;     // 32 KB
constexpr int EE_SM_LIST = EE_SM_A + 2 * EM_BLK;    // u64[EE_LIST + 128]: scene << 32 | i << 16 | j; + a carried partial tile
constexpr int EE_SM_PAR = EE_SM_LIST + (EE_LIST + 128) * 8; // b1[128] b2[128] w_out[128]
constexpr int EE_SM_PART = EE_SM_PAR + 3 * 128 * 4; // float[128]: partial sums of the upper column half
constexpr int EE_SM_BAR = EE_SM_PART + 512;
constexpr int EE_SM_TOTAL = EE_SM_BAR + 48 + 32;
template <bool F16> constexpr uint32_t kIdescEE = make_idesc_op<F16>(128, 128);

__device__ __forceinline__ float elu_fast(float x) { return x > 0.f ? x : ex2_fast(x * 1.4426950408889634f) - 1.0f; }

template <bool F16>
__global__ void __launch_bounds__(256, 2) edge_mlp_tc_kernel(const __nv_bfloat16* __restrict__ nab,   // [R, 256] = [a | b]
                                                             const uint8_t* __restrict__ adj, const uint8_t* __restrict__ Wp,
                                                             const float* __restrict__ b1, const float* __restrict__ b2,
                                                             const float* __restrict__ w_out, const float* __restrict__ b_out,
                                                             int S, int N, float* __restrict__ score, int zero_fill, uint32_t* trap) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];   // link-time constant base: uniform addresses / descriptors
  uint8_t* const smem = smem_dyn;
  require_smem_alignment(smem, trap, 5);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_w = sbase + EE_SM_BAR, bar_mma = bar_w + 8;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + EE_SM_BAR + 16);
  int* s_wtot = reinterpret_cast<int*>(smem + EE_SM_BAR + 48);   // 8 warp totals of the edge scan
  unsigned long long* s_list = reinterpret_cast<unsigned long long*>(smem + EE_SM_LIST);
  float* s_b1 = reinterpret_cast<float*>(smem + EE_SM_PAR);
  float* s_b2 = s_b1 + 128;
  float* s_wo = s_b2 + 128;
  float* s_part = reinterpret_cast<float*>(smem + EE_SM_PART);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive_expect_tx(bar_w, EM_W2_BYTES);
    bulk_g2s(sbase + EE_SM_W2, Wp + EM_W1_BYTES, EM_W2_BYTES, bar_w);
  }
  if (warp == 0) tmem_alloc(sbase + EE_SM_BAR + 16, 128);
  if (tid < 128) {
    s_b1[tid] = b1[tid];
    s_b2[tid] = b2[tid];
    s_wo[tid] = w_out[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const float bo = __ldg(b_out);
  mbar_wait(bar_w, 0, trap, 0x501);
  const int rows_per_chunk = EE_LIST / N > 0 ? EE_LIST / N : 1;
  uint32_t it = 0;

  // one tile of up to 128 edges of the list: gather -> e1 operand -> MMA against W2 -> epilogue -> scatter
  auto edge_tile = [&](int t0, int nt) {
        // ---- e1 tile: warp w builds edges 16 w .. 16 w + 15; lane -> 4 consecutive k, so every gather of a_i / b_j
        //      is one coalesced 512-byte row (a thread-per-edge walk made each load instruction touch 32 sectors
        //      and the tile took ~20 k clk); four edges in flight per warp
        {
          const int k = lane * 4;
          const float4 c4 = *reinterpret_cast<const float4*>(s_b1 + k);
          uint8_t* blk = smem + EE_SM_A + (k >> 6) * EM_BLK + ((k & 7) << 1);
          const int chunk = (k & 63) >> 3;
#pragma unroll
          for (int e0 = 0; e0 < 16; e0 += 8) {
            uint2 av[8], bv[8];   // 4 bf16 of a_i / b_j each: a 256-byte row per warp load
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int e = warp * 16 + e0 + u;
              av[u] = bv[u] = make_uint2(0u, 0u);
              if (e < nt) {
                const unsigned long long en = s_list[t0 + e];
                const uint32_t ij = (uint32_t)en;
                const __nv_bfloat16* nab_s = nab + (size_t)(en >> 32) * N * 256;
                av[u] = __ldg(reinterpret_cast<const uint2*>(nab_s + (size_t)(ij >> 16) * 256) + lane);
                bv[u] = __ldg(reinterpret_cast<const uint2*>(nab_s + (size_t)(ij & 0xffffu) * 256 + 128) + lane);
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int e = warp * 16 + e0 + u;   // rows beyond nt get elu(b1): finite, never read back
              *reinterpret_cast<uint2*>(blk + e * 128 + ((chunk ^ (e & 7)) << 4)) =
                  make_uint2(pack_op2<F16>(elu_fast(op_lo<F16>(av[u].x) + op_lo<F16>(bv[u].x) + c4.x),
                                           elu_fast(op_hi<F16>(av[u].x) + op_hi<F16>(bv[u].x) + c4.y)),
                             pack_op2<F16>(elu_fast(op_lo<F16>(av[u].y) + op_lo<F16>(bv[u].y) + c4.z),
                                           elu_fast(op_hi<F16>(av[u].y) + op_hi<F16>(bv[u].y) + c4.w)));
            }
          }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = make_desc_sw128(sbase + EE_SM_A + (ks >> 2) * EM_BLK) + (uint64_t)((ks & 3) * 2);
            const uint64_t db = make_desc_sw128(sbase + EE_SM_W2 + (ks >> 2) * EM_BLK) + (uint64_t)((ks & 3) * 2);
            umma_bf16(tmem_base, da, db, kIdescEE<F16>, ks ? 1u : 0u);
          }
          umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, it & 1u, trap, 0x502);
        tc_fence_after();
        // ---- epilogue: thread -> (edge row, 64 columns): sum_c elu(acc + b2) * w_out
        {
          const int q = warp & 3, half = warp >> 2;
          const int r = q * 32 + lane;
          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + half * 64;
          float part = 0.f;
#pragma unroll
          for (int ch = 0; ch < 8; ch += 2) {
            float v0[8], v1[8];
            tmem_ld8(t_row + ch * 8, v0);
            tmem_ld8(t_row + ch * 8 + 8, v1);
            tmem_wait_ld();
            const int c = half * 64 + ch * 8;
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(s_b2 + c + i), ww = *reinterpret_cast<const float4*>(s_wo + c + i);
              part = fmaf(elu_fast(v0[i] + bb.x), ww.x, part);
              part = fmaf(elu_fast(v0[i + 1] + bb.y), ww.y, part);
              part = fmaf(elu_fast(v0[i + 2] + bb.z), ww.z, part);
              part = fmaf(elu_fast(v0[i + 3] + bb.w), ww.w, part);
              const float4 bc = *reinterpret_cast<const float4*>(s_b2 + c + 8 + i), wc = *reinterpret_cast<const float4*>(s_wo + c + 8 + i);
              part = fmaf(elu_fast(v1[i] + bc.x), wc.x, part);
              part = fmaf(elu_fast(v1[i + 1] + bc.y), wc.y, part);
              part = fmaf(elu_fast(v1[i + 2] + bc.z), wc.z, part);
              part = fmaf(elu_fast(v1[i + 3] + bc.w), wc.w, part);
            }
          }
          if (half == 1) s_part[r] = part;
          tc_fence_before();
          __syncthreads();   // partial sums visible; TMEM drained; A block free
          if (half == 0 && r < nt) {
            const unsigned long long en = s_list[t0 + r];
            const uint32_t ij = (uint32_t)en;
            const float z = part + s_part[r] + bo;
            score[((size_t)(en >> 32) * N + (ij >> 16)) * N + (ij & 0xffffu)] = 1.0f / (1.0f + __expf(-z));
          }
        }
    ++it;
  };
  // A 64-agent scene has ~190 edges = 1.5 tiles: the partial tile at the end of a scene (or chunk) is not run half empty but
  // carried to the front of the next list (entries name their scene), and flushed once after the CTA's last scene
  int carry = 0;
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    float* sc = score + (size_t)s * N * N;
    const uint8_t* ad = adj + (size_t)s * N * N;
    if (zero_fill)
      for (int i = tid; i < (N * N) >> 2; i += 256) reinterpret_cast<float4*>(sc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < N; r0 += rows_per_chunk) {
      // ---- compact the edges of rows [r0, r1): every thread takes 16 consecutive entries of the adjacency block (one
      //      128-bit load; the byte-at-a-time walk with a ballot and a shared atomic per 256 entries was 57 % of the
      //      kernel's stall samples, almost all of it the latency of 16 dependent rounds of 1-byte loads), counts its
      //      edges, and a block-wide exclusive scan gives its place in the list -- ascending (row, column) order
      const int r1 = min(N, r0 + rows_per_chunk);
      const int tot = (r1 - r0) * N;             // <= EE_LIST = 16 x 256
      const uint8_t* src = ad + (size_t)r0 * N;
      const int e_base = tid * 16;
      uint32_t wv[4] = {0u, 0u, 0u, 0u};
      if (e_base < tot) {
        if (e_base + 16 <= tot && (reinterpret_cast<uintptr_t>(src + e_base) & 15u) == 0u) {
          const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(src + e_base));
          wv[0] = v4.x; wv[1] = v4.y; wv[2] = v4.z; wv[3] = v4.w;
        } else {
          for (int bb = 0; bb < 16; ++bb)
            if (e_base + bb < tot && src[e_base + bb] != 0) wv[bb >> 2] |= 1u << ((bb & 3) * 8);
        }
      }
      uint32_t em = 0;                           // bit b: entry e_base + b is an edge
#pragma unroll
      for (int bb = 0; bb < 16; ++bb) em |= ((wv[bb >> 2] >> ((bb & 3) * 8)) & 0xffu) ? (1u << bb) : 0u;
      const int cnt = __popc(em);
      int incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      __syncthreads();                           // the previous chunk's list and warp totals are consumed
      if (lane == 31) s_wtot[warp] = incl;
      __syncthreads();
      int base = carry + incl - cnt, ne = 0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) {
        const int tw = s_wtot[w8];
        if (w8 < warp) base += tw;
        ne += tw;
      }
      if (em) {
        int row = r0 + e_base / N, col = e_base % N;
        for (int bb = 0; bb < 16; ++bb) {
          if (em & (1u << bb)) s_list[base++] = ((unsigned long long)s << 32) | ((uint32_t)row << 16) | (uint32_t)col;
          if (++col == N) {
            col = 0;
            ++row;
          }
        }
      }
      __syncthreads();
      const int n_all = carry + ne;
      const int n_full = n_all & ~127;
      for (int t0 = 0; t0 < n_full; t0 += 128) edge_tile(t0, 128);
      __syncthreads();   // the scatter of the last tile has read the list
      carry = n_all - n_full;
      if (n_full > 0 && tid < carry) s_list[tid] = s_list[n_full + tid];   // n_full >= 128 > carry: no overlap
      __syncthreads();   // the list is rebuilt by the next chunk
    }
  }
  if (carry > 0) edge_tile(0, carry);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// h[R, U] (row stride ld_h) -> score[S,N,N] on the edges of adj; nab: >= R*256 floats of scratch.
// ld_h < 0: h is the tile-blocked bf16 state (rows padded to whole tiles) instead of fp32 rows.
template <bool F16>
static int edge_mlp_tc_run(const float* h, int ld_h, const uint8_t* adj, const void* packed, const mmt_edge_weights* w, int S,
                           int N, float* score, float* nab, int zero_fill, cudaStream_t stream) {
  const int R = S * N, tiles = (R + 127) / 128;
  static DeviceMask smem_opted[3];   // per kernel: devices already opted in
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&node_proj_tc_kernel<false, F16>), NP_SM_TOTAL + 1024, &smem_opted[0])) return rc;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&node_proj_tc_kernel<true, F16>), NP_SM_TOTAL + 1024, &smem_opted[1])) return rc;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&edge_mlp_tc_kernel<F16>), EE_SM_TOTAL + 1024, &smem_opted[2])) return rc;
  const uint8_t* Wp = reinterpret_cast<const uint8_t*>(packed);
  int grid = tiles < 2 * num_sms() ? tiles : 2 * num_sms();
  __nv_bfloat16* nabh = reinterpret_cast<__nv_bfloat16*>(nab);   // R x 256 16-bit words inside the R x 256 float scratch
  if (ld_h < 0) node_proj_tc_kernel<true, F16><<<grid, 256, NP_SM_TOTAL + 1024, stream>>>(h, 0, R, Wp, nabh, tiles, trap_record());
  else node_proj_tc_kernel<false, F16><<<grid, 256, NP_SM_TOTAL + 1024, stream>>>(h, ld_h, R, Wp, nabh, tiles, trap_record());
  count_launch();
  int rc = check_launch("node_proj_tc_kernel");
  if (rc) return rc;
  grid = S < 2 * num_sms() ? S : 2 * num_sms();
  edge_mlp_tc_kernel<F16><<<grid, 256, EE_SM_TOTAL + 1024, stream>>>(nabh, adj, Wp, w->b1, w->b2, w->w_out, w->b_out, S, N, score,
                                                                     zero_fill, trap_record());
  count_launch();
  return check_launch("edge_mlp_tc_kernel");
}

int launch_edge_mlp_tc(const float* h, int ld_h, const uint8_t* adj, const void* packed, const mmt_edge_weights* w, int S,
                       int N, float* score, float* nab, int zero_fill, int f16, cudaStream_t stream) {
  return f16 ? edge_mlp_tc_run<true>(h, ld_h, adj, packed, w, S, N, score, nab, zero_fill, stream)
             : edge_mlp_tc_run<false>(h, ld_h, adj, packed, w, S, N, score, nab, zero_fill, stream);
}

// node projections only (the backward kernel of edge_mlp_bwd_tc.cu gathers the same [a | b] rows)
int launch_node_proj_tc(const float* h, int ld_h, const void* packed, int R, __nv_bfloat16* out, cudaStream_t stream) {
  static DeviceMask smem_opted[1];
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&node_proj_tc_kernel<false>), NP_SM_TOTAL + 1024, &smem_opted[0])) return rc;
  const int tiles = (R + 127) / 128;
  const int grid = tiles < 2 * num_sms() ? tiles : 2 * num_sms();
  node_proj_tc_kernel<false><<<grid, 256, NP_SM_TOTAL + 1024, stream>>>(h, ld_h, R, reinterpret_cast<const uint8_t*>(packed), out,
                                                                        tiles, trap_record());
  count_launch();
  return check_launch("node_proj_tc_kernel");
}

int launch_pack_edge_weights(const float* W1, const float* W2, void* packed, int f16, cudaStream_t stream) {
  const int n = 256 * 128 + 128 * 128;
  if (f16) pack_edge_weights_kernel<true><<<(n + 255) / 256, 256, 0, stream>>>(W1, W2, reinterpret_cast<uint8_t*>(packed));
  else pack_edge_weights_kernel<false><<<(n + 255) / 256, 256, 0, stream>>>(W1, W2, reinterpret_cast<uint8_t*>(packed));
  count_launch();
  return check_launch("pack_edge_weights_kernel");
}

}  // namespace mmt

extern "C" size_t mmt_edge_weights_packed_bytes(int U, int He) {
  if (U != mmt::EM_U || He != mmt::EM_HE) return 0;
  return (size_t)mmt::EM_W1_BYTES + mmt::EM_W2_BYTES;
}

extern "C" int mmt_pack_edge_weights_bf16(const float* W1, const float* W2, int U, int He, void* packed, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(W1 && W2 && packed, "W1/W2/packed must not be NULL");
  MMT_REQUIRE(U == EM_U && He == EM_HE, "packing is built for U = 128, He = 128");
  MMT_ALIGNED(packed);
  return launch_pack_edge_weights(W1, W2, packed, 0, (cudaStream_t)stream);
}

extern "C" int mmt_edge_mlp_bf16(const float* h, const uint8_t* adj, const void* packed, const float* b1, const float* b2,
                                 const float* w_out, const float* b_out, int S, int N, int U, int He, float* score,
                                 float* work, size_t work_bytes, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && N % 4 == 0 && N <= 1024, "need 0 < N <= 1024, N % 4 == 0");
  MMT_REQUIRE(U == EM_U && He == EM_HE, "tensor-core edge MLP is built for U = 128, He = 128");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(h && adj && packed && b1 && b2 && w_out && b_out && score && work, "all pointers required");
  MMT_ALIGNED(h);
  MMT_ALIGNED(packed);
  MMT_ALIGNED(score);
  MMT_ALIGNED(work);
  if (work_bytes < sizeof(float) * 2 * (size_t)S * N * He) {
    set_error("mmt_edge_mlp_bf16: workspace too small");
    return MMT_EWORKSPACE;
  }
  mmt_edge_weights w{nullptr, b1, nullptr, b2, w_out, b_out, He};
  return launch_edge_mlp_tc(h, U, adj, packed, &w, S, N, score, work, 1, 0, (cudaStream_t)stream);
}
