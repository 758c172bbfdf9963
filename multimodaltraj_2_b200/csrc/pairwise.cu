// Pairwise pedestrian-distance kernel + adjacency (include/mmt.h).
//
// HBM-bound: per scene-frame it reads 9N bytes and writes 5N^2 (+4N) bytes.  One CTA stages
// the scene's positions in shared memory (float4 loads), then every thread produces quads of
// four consecutive j for one i so that each warp-wide store instruction is one fully coalesced
// 512-byte (kern, float4) or 128-byte (adj, uchar4) segment.  Row degrees are reduced with
// segmented warp shuffles.  The distance uses __fmul_rn/__fadd_rn so no FMA contraction can
// change the rounding: adj/deg are bit-identical to the fp32 oracle.
#include "mmt_common.cuh"

namespace mmt {

__global__ void __launch_bounds__(256) pairwise_adj_kernel(const float* __restrict__ pos,
                                                           const uint8_t* __restrict__ valid, int S, int N,
                                                           float r2, float inv_2sigma2, float* __restrict__ kern,
                                                           uint8_t* __restrict__ adj, int32_t* __restrict__ deg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sx = reinterpret_cast<float*>(smem_raw);  // [N]
  float* sy = sx + N;                               // [N]
  int* sdeg = reinterpret_cast<int*>(sy + N);       // [N]
  uint8_t* sv = reinterpret_cast<uint8_t*>(sdeg + N);  // [N]

  const int lpr = N >> 2;  // quads (lanes) per row
  const bool seg = (lpr & (lpr - 1)) == 0 && lpr <= 32;
  const int nquads = N * lpr;
  const int lane = threadIdx.x & 31;

  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    // ---- stage positions: one float4 = two agents (x0,y0,x1,y1)
    const float4* p4 = reinterpret_cast<const float4*>(pos + (size_t)s * N * 2);
    for (int a = threadIdx.x; a < (N >> 1); a += blockDim.x) {
      float4 v = __ldg(p4 + a);
      sx[2 * a] = v.x;
      sy[2 * a] = v.y;
      sx[2 * a + 1] = v.z;
      sy[2 * a + 1] = v.w;
    }
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
      sv[a] = valid[(size_t)s * N + a];
      sdeg[a] = 0;
    }
    __syncthreads();

    const size_t base = (size_t)s * N * N;
    const int iters = (nquads + blockDim.x - 1) / blockDim.x;
    for (int it = 0; it < iters; ++it) {
      const int q = it * blockDim.x + threadIdx.x;
      const bool active = q < nquads;
      int cnt = 0;
      int i = 0;
      if (active) {
        i = q / lpr;
        const int j0 = (q - i * lpr) << 2;
        const float xi = sx[i], yi = sy[i];
        const bool vi = sv[i] != 0;
        const float4 xj = *reinterpret_cast<const float4*>(sx + j0);
        const float4 yj = *reinterpret_cast<const float4*>(sy + j0);
        const uchar4 vj = *reinterpret_cast<const uchar4*>(sv + j0);
        float d2[4];
        {
          float dx, dy;
          dx = __fsub_rn(xi, xj.x); dy = __fsub_rn(yi, yj.x);
          d2[0] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(xi, xj.y); dy = __fsub_rn(yi, yj.y);
          d2[1] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(xi, xj.z); dy = __fsub_rn(yi, yj.z);
          d2[2] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(xi, xj.w); dy = __fsub_rn(yi, yj.w);
          d2[3] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        }
        const bool a0 = vi && vj.x && (j0 + 0 != i) && (d2[0] < r2);
        const bool a1 = vi && vj.y && (j0 + 1 != i) && (d2[1] < r2);
        const bool a2 = vi && vj.z && (j0 + 2 != i) && (d2[2] < r2);
        const bool a3 = vi && vj.w && (j0 + 3 != i) && (d2[3] < r2);
        cnt = (int)a0 + (int)a1 + (int)a2 + (int)a3;
        if (kern != nullptr) {
          float4 k;
          k.x = a0 ? expf(-(d2[0] * inv_2sigma2)) : 0.0f;
          k.y = a1 ? expf(-(d2[1] * inv_2sigma2)) : 0.0f;
          k.z = a2 ? expf(-(d2[2] * inv_2sigma2)) : 0.0f;
          k.w = a3 ? expf(-(d2[3] * inv_2sigma2)) : 0.0f;
          st_cs_f4(kern + base + ((size_t)q << 2), k);
        }
        if (adj != nullptr) {
          uchar4 m = make_uchar4(a0, a1, a2, a3);
          __stcs(reinterpret_cast<uchar4*>(adj + base + ((size_t)q << 2)), m);
        }
      }
      if (deg != nullptr) {
        if (seg) {
          // a row occupies lpr consecutive lanes of this warp instruction: segmented reduce
          for (int o = lpr >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
          if (active && (lane & (lpr - 1)) == 0) sdeg[i] = cnt;
        } else if (active && cnt) {
          atomicAdd(&sdeg[i], cnt);
        }
      }
    }
    __syncthreads();
    if (deg != nullptr)
      for (int a = threadIdx.x; a < N; a += blockDim.x) deg[(size_t)s * N + a] = sdeg[a];
    __syncthreads();
  }
}

// one warp per (scene,row): ordered compaction of the adjacency row into neighbour indices
__global__ void __launch_bounds__(256) neighbor_index_kernel(const uint8_t* __restrict__ adj, int rows, int N,
                                                             int max_nbr, int32_t* __restrict__ nbr,
                                                             int32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const uint8_t* row = adj + (size_t)r * N;
    int32_t* out = nbr + (size_t)r * max_nbr;
    int n = 0;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const bool a = j < N && row[j] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, a);
      const int pos = n + __popc(m & ((1u << lane) - 1u));
      if (a && pos < max_nbr) out[pos] = j;
      n += __popc(m);
    }
    const int w = n < max_nbr ? n : max_nbr;
    for (int k = w + lane; k < max_nbr; k += 32) out[k] = -1;
    if (lane == 0) cnt[r] = w;
  }
}

}  // namespace mmt

extern "C" int mmt_pairwise_adj_f32(const float* pos, const uint8_t* valid, int S, int N, float r2,
                                    float inv_2sigma2, float* kern, uint8_t* adj, int32_t* deg, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && N % 4 == 0 && N <= 1024, "need S >= 0, 0 < N <= 1024, N % 4 == 0");
  if (S == 0) return MMT_OK;  // empty batch: nothing to read or write
  MMT_REQUIRE(pos && valid, "pos/valid must not be NULL");
  MMT_ALIGNED(pos);
  MMT_ALIGNED(kern);
  if (adj && (reinterpret_cast<uintptr_t>(adj) & 3u)) {
    set_error("mmt_pairwise_adj_f32: adj not 4-byte aligned");
    return MMT_EALIGN;
  }
  if (S == 0) return MMT_OK;
  const size_t smem = (size_t)N * (4 + 4 + 4 + 1) + 16;
  // persistent-style grid: a multiple of the SM count, 8 resident CTAs per SM
  int grid = S < kNumSMs * 8 ? S : kNumSMs * 8;
  pairwise_adj_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(pos, valid, S, N, r2, inv_2sigma2, kern, adj, deg);
  count_launch();
  return check_launch("pairwise_adj_kernel");
}

extern "C" int mmt_neighbor_index_i32(const uint8_t* adj, int S, int N, int max_nbr, int32_t* nbr, int32_t* cnt,
                                      void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && max_nbr > 0, "need S >= 0, N > 0, max_nbr > 0");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(adj && nbr && cnt, "adj/nbr/cnt must not be NULL");
  if (S == 0) return MMT_OK;
  const long rows = (long)S * N;
  long blocks = (rows + 7) / 8;
  int grid = blocks < kNumSMs * 8 ? (int)blocks : kNumSMs * 8;
  neighbor_index_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(adj, (int)rows, N, max_nbr, nbr, cnt);
  count_launch();
  return check_launch("neighbor_index_kernel");
}
