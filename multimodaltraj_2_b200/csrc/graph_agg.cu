// Fused graph step of the bf16 rollout: pairwise distance kernel + adjacency + masked softmax +
// aggregation of neighbour states in ONE kernel, so neither the N x N kernel matrix nor the
// adjacency mask ever goes to HBM (SURVEY 8d "fused ideal").  Same arithmetic as pairwise.cu
// (d2 with separately rounded ops -> identical adjacency) followed by aggregate.cu's warp-per-row
// softmax + sparse gather; inputs h (bf16) and c (fp32), outputs mh, mc (bf16).
// HBM-bound: reads 8 + 1 B/agent of positions/validity and ~768 B/agent of state (neighbour rows
// come from L2), writes 512 B/agent.
#include <cuda_bf16.h>

#include <cuda_fp16.h>

#include "mmt_common.cuh"

namespace mmt {

constexpr int kGaWarps = 8;

// 16-bit state words: bf16 (MMT_PREC_BF16*) or fp16 (MMT_PREC_F16) -- two per 32-bit word, low half first
template <bool F16>
__device__ __forceinline__ float st_lo(uint32_t w) {
  if constexpr (F16) return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu)));
  else return __uint_as_float(w << 16);
}
template <bool F16>
__device__ __forceinline__ float st_hi(uint32_t w) {
  if constexpr (F16) return __half2float(__ushort_as_half((unsigned short)(w >> 16)));
  else return __uint_as_float(w & 0xffff0000u);
}
template <bool F16>
__device__ __forceinline__ uint32_t st_pack(float lo, float hi) {
  if constexpr (F16) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&b);
  }
}

template <bool F16>
__global__ void __launch_bounds__(kGaWarps * 32) graph_aggregate_bf16_kernel(
    const float* __restrict__ pos, const uint8_t* __restrict__ valid, const __nv_bfloat16* __restrict__ hb,
    const float* __restrict__ c, int rows, int N, int U, float r2, float inv_2sigma2,
    __nv_bfloat16* __restrict__ mhb, __nv_bfloat16* __restrict__ mcb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* nb_idx = reinterpret_cast<int*>(smem_raw) + warp * N;
  float* nb_w = reinterpret_cast<float*>(smem_raw) + kGaWarps * N + warp * N;

  for (int r = blockIdx.x * kGaWarps + warp; r < rows; r += gridDim.x * kGaWarps) {
    const int s = r / N, i = r - s * N;
    const float2 pi = __ldg(reinterpret_cast<const float2*>(pos) + r);
    const bool vi = valid[r] != 0;
    int n = 0;
    float mx = -INFINITY;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      bool a = false;
      float kv = 0.f;
      if (j < N && vi) {
        const float2 pj = __ldg(reinterpret_cast<const float2*>(pos) + (size_t)s * N + j);
        const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        a = valid[(size_t)s * N + j] != 0 && j != i && d2 < r2;
        if (a) kv = expf(-(d2 * inv_2sigma2));
      }
      const unsigned m = __ballot_sync(0xffffffffu, a);
      if (a) {
        const int p = n + __popc(m & ((1u << lane) - 1u));
        nb_idx[p] = j;
        nb_w[p] = kv;
        mx = fmaxf(mx, kv);
      }
      n += __popc(m);
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.f;
    for (int k = lane; k < n; k += 32) {
      const float e = expf(nb_w[k] - mx);
      nb_w[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = n > 0 ? 1.0f / sum : 0.0f;
    __syncwarp();
    // gather: lane owns 4 consecutive units of h (8 B of bf16) and of c (16 B of fp32); U == 128
    float ah[4] = {0.f, 0.f, 0.f, 0.f}, ac[4] = {0.f, 0.f, 0.f, 0.f};
    const size_t sbase = (size_t)s * N;
    for (int k = 0; k < n; ++k) {
      const float w = nb_w[k] * inv;
      const size_t jr = sbase + nb_idx[k];
      const uint2 hv = __ldg(reinterpret_cast<const uint2*>(hb + jr * U) + lane);
      const float4 cv = __ldg(reinterpret_cast<const float4*>(c + jr * U) + lane);
      ah[0] = fmaf(w, st_lo<F16>(hv.x), ah[0]);
      ah[1] = fmaf(w, st_hi<F16>(hv.x), ah[1]);
      ah[2] = fmaf(w, st_lo<F16>(hv.y), ah[2]);
      ah[3] = fmaf(w, st_hi<F16>(hv.y), ah[3]);
      ac[0] = fmaf(w, cv.x, ac[0]);
      ac[1] = fmaf(w, cv.y, ac[1]);
      ac[2] = fmaf(w, cv.z, ac[2]);
      ac[3] = fmaf(w, cv.w, ac[3]);
    }
    reinterpret_cast<uint2*>(mhb + (size_t)r * U)[lane] = make_uint2(st_pack<F16>(ah[0], ah[1]), st_pack<F16>(ah[2], ah[3]));
    reinterpret_cast<uint2*>(mcb + (size_t)r * U)[lane] = make_uint2(st_pack<F16>(ac[0], ac[1]), st_pack<F16>(ac[2], ac[3]));
    __syncwarp();
  }
}

int launch_graph_aggregate_bf16(const float* pos, const uint8_t* valid, const void* hb, const float* c, int S, int N,
                                int U, float r2, float inv_2sigma2, void* mhb, void* mcb, int f16, cudaStream_t stream) {
  if (U != 128) {
    set_error("graph_aggregate_bf16: built for U = 128");
    return MMT_EARG;
  }
  const long rows = (long)S * N;
  long blocks = (rows + kGaWarps - 1) / kGaWarps;
  int grid = blocks < (long)num_sms() * 8 ? (int)blocks : num_sms() * 8;
  const size_t smem = (size_t)kGaWarps * N * 8;
  if (f16)
    graph_aggregate_bf16_kernel<true><<<grid, kGaWarps * 32, smem, stream>>>(
        pos, valid, reinterpret_cast<const __nv_bfloat16*>(hb), c, (int)rows, N, U, r2, inv_2sigma2,
        reinterpret_cast<__nv_bfloat16*>(mhb), reinterpret_cast<__nv_bfloat16*>(mcb));
  else
    graph_aggregate_bf16_kernel<false><<<grid, kGaWarps * 32, smem, stream>>>(
        pos, valid, reinterpret_cast<const __nv_bfloat16*>(hb), c, (int)rows, N, U, r2, inv_2sigma2,
        reinterpret_cast<__nv_bfloat16*>(mhb), reinterpret_cast<__nv_bfloat16*>(mcb));
  count_launch();
  return check_launch("graph_aggregate_bf16_kernel");
}

}  // namespace mmt

// ---------------------------------------------------------------------------------------------------------
// Tile-blocked variant (state layout [tile][U/8][128 rows][8 units], see cell_tc.cu): one CTA per 128-row tile
// (= 128/N whole scenes).  The tile's h (32 KB bf16) and c (64 KB fp32) are staged once in shared memory with
// fully coalesced 16-byte loads, so every neighbour gather is a shared-memory read; adjacency, softmax and
// the weighted sums are one warp per row as above.  Requires 128 % N == 0 and U == 128.
namespace mmt {

constexpr int GB_THREADS = 256;
constexpr int GB_HROW = 128 + 8;    // bf16 elements per staged h row (+16 B pad: conflict-free 16 B stores)
constexpr int GB_CROW = 128 + 4;    // floats per staged c row (+16 B pad)

template <bool F16>
__global__ void __launch_bounds__(GB_THREADS, 2) graph_aggregate_blocked_kernel(
    const float* __restrict__ pos, const uint8_t* __restrict__ valid, const __nv_bfloat16* __restrict__ hb,
    const float* __restrict__ c, int R, int N, float r2, float inv_2sigma2, __nv_bfloat16* __restrict__ mhb,
    __nv_bfloat16* __restrict__ mcb, int num_tiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(smem_raw);                       // [128][GB_HROW]
  float* sc = reinterpret_cast<float*>(smem_raw + 128 * GB_HROW * 2);                   // [128][GB_CROW]
  float2* spos = reinterpret_cast<float2*>(sc + 128 * GB_CROW);                         // [128]
  int* nb_all = reinterpret_cast<int*>(spos + 128);                                     // [8 warps][N]
  float* nw_all = reinterpret_cast<float*>(nb_all + 8 * N);                             // [8 warps][N]
  uint8_t* sval = reinterpret_cast<uint8_t*>(nw_all + 8 * N);                           // [128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int* nb_idx = nb_all + warp * N;
  float* nb_w = nw_all + warp * N;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int row0 = tile * 128;
    // ---- stage the tile's state (blocked global -> row-major padded smem), 24 independent 16 B loads per thread
    {
      const uint4* hsrc = reinterpret_cast<const uint4*>(hb + (size_t)tile * 128 * 128);
      uint4 hv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) hv[k] = __ldg(hsrc + tid + 256 * k);
      const uint4* csrc = reinterpret_cast<const uint4*>(c + (size_t)tile * 128 * 128);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint4 cv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cv[k] = __ldg(csrc + tid + 256 * (k + 8 * half));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int qd = tid + 256 * (k + 8 * half), piece = qd >> 1, hf = qd & 1;
          const int g = piece >> 7, rr = piece & 127;
          *reinterpret_cast<uint4*>(sc + rr * GB_CROW + g * 8 + hf * 4) = cv[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int p = tid + 256 * k, g = p >> 7, rr = p & 127;
        *reinterpret_cast<uint4*>(sh + rr * GB_HROW + g * 8) = hv[k];
      }
      if (tid < 128) {
        const int gr = row0 + tid;
        spos[tid] = gr < R ? __ldg(reinterpret_cast<const float2*>(pos) + gr) : make_float2(0.f, 0.f);
        sval[tid] = gr < R ? valid[gr] : 0;
      }
    }
    __syncthreads();
    // ---- one warp per row: adjacency -> softmax -> gather from smem
    for (int rr = warp * 16; rr < warp * 16 + 16; ++rr) {
      const int sb = (rr / N) * N, i = rr - sb;
      const float2 pi = spos[rr];
      const bool vi = sval[rr] != 0;
      int n = 0;
      float mx = -INFINITY;
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        bool a = false;
        float kv = 0.f;
        if (j < N && vi) {
          const float2 pj = spos[sb + j];
          const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y);
          const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          a = sval[sb + j] != 0 && j != i && d2 < r2;
          if (a) kv = expf(-(d2 * inv_2sigma2));
        }
        const unsigned m = __ballot_sync(0xffffffffu, a);
        if (a) {
          const int p = n + __popc(m & ((1u << lane) - 1u));
          nb_idx[p] = sb + j;
          nb_w[p] = kv;
          mx = fmaxf(mx, kv);
        }
        n += __popc(m);
      }
      mx = warp_max(mx);
      __syncwarp();
      float sum = 0.f;
      for (int k = lane; k < n; k += 32) {
        const float e = expf(nb_w[k] - mx);
        nb_w[k] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = n > 0 ? 1.0f / sum : 0.0f;
      __syncwarp();
      float ah[4] = {0.f, 0.f, 0.f, 0.f}, ac[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < n; ++k) {
        const float w = nb_w[k] * inv;
        const int j = nb_idx[k];
        const uint2 hv = *reinterpret_cast<const uint2*>(sh + j * GB_HROW + lane * 4);
        const float4 cv = *reinterpret_cast<const float4*>(sc + j * GB_CROW + lane * 4);
        ah[0] = fmaf(w, st_lo<F16>(hv.x), ah[0]);
        ah[1] = fmaf(w, st_hi<F16>(hv.x), ah[1]);
        ah[2] = fmaf(w, st_lo<F16>(hv.y), ah[2]);
        ah[3] = fmaf(w, st_hi<F16>(hv.y), ah[3]);
        ac[0] = fmaf(w, cv.x, ac[0]);
        ac[1] = fmaf(w, cv.y, ac[1]);
        ac[2] = fmaf(w, cv.z, ac[2]);
        ac[3] = fmaf(w, cv.w, ac[3]);
      }
      // blocked store: lane's 4 units live in group g = lane/2, half (lane&1)
      const size_t o = ((size_t)(tile * 16 + (lane >> 1)) * 128 + rr) * 8 + (lane & 1) * 4;
      *reinterpret_cast<uint2*>(mhb + o) = make_uint2(st_pack<F16>(ah[0], ah[1]), st_pack<F16>(ah[2], ah[3]));
      *reinterpret_cast<uint2*>(mcb + o) = make_uint2(st_pack<F16>(ac[0], ac[1]), st_pack<F16>(ac[2], ac[3]));
      __syncwarp();
    }
    __syncthreads();
  }
}

int launch_graph_aggregate_blocked(const float* pos, const uint8_t* valid, const void* hb, const float* c, int S, int N,
                                   float r2, float inv_2sigma2, void* mhb, void* mcb, int f16, cudaStream_t stream) {
  const int R = S * N, tiles = (R + 127) / 128;
  const size_t smem = 128 * GB_HROW * 2 + 128 * GB_CROW * 4 + 128 * 8 + (size_t)8 * N * 8 + 128 + 16;
  static DeviceMask smem_opted[2];   // per kernel: devices already opted in
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&graph_aggregate_blocked_kernel<false>), 112 * 1024, &smem_opted[0])) return rc;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&graph_aggregate_blocked_kernel<true>), 112 * 1024, &smem_opted[1])) return rc;
  const int grid = tiles < 2 * num_sms() ? tiles : 2 * num_sms();
  if (f16)
    graph_aggregate_blocked_kernel<true><<<grid, GB_THREADS, smem, stream>>>(
        pos, valid, reinterpret_cast<const __nv_bfloat16*>(hb), c, R, N, r2, inv_2sigma2,
        reinterpret_cast<__nv_bfloat16*>(mhb), reinterpret_cast<__nv_bfloat16*>(mcb), tiles);
  else
    graph_aggregate_blocked_kernel<false><<<grid, GB_THREADS, smem, stream>>>(
        pos, valid, reinterpret_cast<const __nv_bfloat16*>(hb), c, R, N, r2, inv_2sigma2,
        reinterpret_cast<__nv_bfloat16*>(mhb), reinterpret_cast<__nv_bfloat16*>(mcb), tiles);
  count_launch();
  return check_launch("graph_aggregate_blocked_kernel");
}

}  // namespace mmt
