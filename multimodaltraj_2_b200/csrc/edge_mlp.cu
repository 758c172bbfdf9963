// Relational edge MLP of g2k_lstm_mcr, fp32 parity mode (include/mmt.h).
//
//   a = h W1[:U], b = h W1[U:]                         node level: one SGEMM each  [R,U]x[U,He]
//   e1_ij = elu(a_i + b_j + b1); e2_ij = elu(e1_ij W2 + b2); score_ij = sigmoid(w_out.e2_ij + b_out)
// evaluated only on the edges of the adjacency mask (the crowd graphs are sparse: ~3 neighbours
// per agent), gathered per scene into tiles of 64 edges so W2 is staged in shared memory once per
// CTA and reused by every tile.
#include "mmt_common.cuh"

namespace mmt {

// generic C[M,N] = A[M,K] (lda) * B[K,N] (ldb), 64x64 tile, 16-deep chunks, 4x4 per thread
__global__ void __launch_bounds__(256) sgemm_nn_kernel(const float* __restrict__ A, int lda,
                                                       const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                       int ldc, int M, int N, int K) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int row0 = blockIdx.x * 64, col0 = blockIdx.y * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    {
      const int r = tid >> 2, kk = (tid & 3) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < M) v = *reinterpret_cast<const float4*>(A + (size_t)(row0 + r) * lda + k0 + kk);
      As[kk + 0][r] = v.x; As[kk + 1][r] = v.y; As[kk + 2][r] = v.z; As[kk + 3][r] = v.w;
      const int bk = tid >> 4, bc = (tid & 15) << 2;
      *reinterpret_cast<float4*>(&Bs[bk][bc]) =
          __ldg(reinterpret_cast<const float4*>(B + (size_t)(k0 + bk) * ldb + col0 + bc));
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r < M)
      *reinterpret_cast<float4*>(C + (size_t)r * ldc + col0 + tx * 4) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }

constexpr int ETILE = 64;     // edges per tile
constexpr int ECAP = 4096;    // edge-list capacity per row chunk

__global__ void __launch_bounds__(256) edge_mlp_kernel(const float* __restrict__ na, const float* __restrict__ nb,
                                                       const uint8_t* __restrict__ adj, const float* __restrict__ b1,
                                                       const float* __restrict__ W2, const float* __restrict__ b2,
                                                       const float* __restrict__ w_out, const float* __restrict__ b_out,
                                                       int S, int N, int He, float* __restrict__ score) {
  extern __shared__ __align__(16) float sm[];
  float* sW2 = sm;                                   // [He][He]
  float* sE1 = sW2 + He * He;                        // [ETILE][He+1]
  int* sList = reinterpret_cast<int*>(sE1 + ETILE * (He + 1));  // [ECAP]  (i << 16 | j)
  __shared__ int sCount;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < He * He; i += blockDim.x) sW2[i] = __ldg(W2 + i);
  const int nq = He >> 5;  // output columns per lane (He / 32) <= 4
  float wo[4], bb2[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    wo[q] = q < nq ? __ldg(w_out + lane + 32 * q) : 0.f;
    bb2[q] = q < nq ? __ldg(b2 + lane + 32 * q) : 0.f;
  }
  const float bo = __ldg(b_out);
  const int rows_per_chunk = ECAP / N > 0 ? ECAP / N : 1;

  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    float* sc = score + (size_t)s * N * N;
    for (int i = tid; i < (N * N) >> 2; i += blockDim.x) reinterpret_cast<float4*>(sc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < N; r0 += rows_per_chunk) {
      if (tid == 0) sCount = 0;
      __syncthreads();
      const int r1 = min(N, r0 + rows_per_chunk);
      const int tot = (r1 - r0) * N;
      for (int e0 = 0; e0 < tot; e0 += blockDim.x) {
        const int e = e0 + tid;
        const bool is = e < tot && adj[(size_t)s * N * N + (size_t)r0 * N + e] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, is);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&sCount, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is) {
          const int i = r0 + e / N, j = e % N;
          sList[base + __popc(m & ((1u << lane) - 1u))] = (i << 16) | j;
        }
      }
      __syncthreads();
      const int ne = sCount;
      for (int t0 = 0; t0 < ne; t0 += ETILE) {
        const int nt = min(ETILE, ne - t0);
        // e1 tile
        for (int idx = tid; idx < ETILE * He; idx += blockDim.x) {
          const int t = idx / He, m = idx - t * He;
          float v = 0.f;
          if (t < nt) {
            const int ij = sList[t0 + t];
            const int i = ij >> 16, j = ij & 0xffff;
            v = elu_f(na[((size_t)s * N + i) * He + m] + nb[((size_t)s * N + j) * He + m] + __ldg(b1 + m));
          }
          sE1[t * (He + 1) + m] = v;
        }
        __syncthreads();
        // e2 = elu(e1 W2 + b2); warp w owns edges w*8..w*8+7, lane owns columns lane + 32q
        float acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
        for (int k = 0; k < He; ++k) {
          float w2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) w2[q] = q < nq ? sW2[k * He + lane + 32 * q] : 0.f;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float ev = sE1[(warp * 8 + r) * (He + 1) + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(ev, w2[q], acc[r][q]);
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float part = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < nq) part = fmaf(elu_f(acc[r][q] + bb2[q]), wo[q], part);
          part = warp_sum(part);
          const int t = warp * 8 + r;
          if (lane == 0 && t < nt) {
            const int ij = sList[t0 + t];
            sc[(size_t)(ij >> 16) * N + (ij & 0xffff)] = sigmoid_acc(part + bo);
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

int launch_sgemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                 cudaStream_t stream) {
  dim3 grid((M + 63) / 64, N / 64);
  sgemm_nn_kernel<<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, M, N, K);
  count_launch();
  return check_launch("sgemm_nn_kernel");
}

int launch_edge_mlp_f32(const float* h, int ld_h, const uint8_t* adj, const mmt_edge_weights* w, int S, int N, int U,
                        float* score, float* work, cudaStream_t stream) {
  const int He = w->He;
  const long R = (long)S * N;
  float* na = work;
  float* nb = work + R * He;
  int rc = launch_sgemm(h, ld_h, w->W1, He, na, He, (int)R, He, U, stream);
  if (rc) return rc;
  rc = launch_sgemm(h, ld_h, w->W1 + (size_t)U * He, He, nb, He, (int)R, He, U, stream);
  if (rc) return rc;
  const size_t smem = sizeof(float) * ((size_t)He * He + (size_t)ETILE * (He + 1)) + sizeof(int) * ECAP;
  static DeviceMask smem_opted[1];   // per kernel: devices already opted in
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&edge_mlp_kernel), 160 * 1024, &smem_opted[0])) return rc;
  int grid = S < num_sms() ? S : num_sms();
  edge_mlp_kernel<<<grid, 256, smem, stream>>>(na, nb, adj, w->b1, w->W2, w->b2, w->w_out, w->b_out, S, N, He, score);
  count_launch();
  return check_launch("edge_mlp_kernel");
}

}  // namespace mmt

extern "C" int mmt_edge_mlp_f32(const float* h, const uint8_t* adj, const float* W1, const float* b1, const float* W2,
                                const float* b2, const float* w_out, const float* b_out, int S, int N, int U, int He,
                                float* score, float* work, size_t work_bytes, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(h && adj && W1 && b1 && W2 && b2 && w_out && b_out && score && work, "all pointers required");
  MMT_REQUIRE(S >= 0 && N > 0 && N % 4 == 0 && N <= 1024, "need 0 < N <= 1024, N % 4 == 0");
  MMT_REQUIRE(U > 0 && U % 16 == 0 && He >= 64 && He <= 128 && He % 64 == 0, "need U % 16 == 0, He in {64,128}");
  MMT_ALIGNED(h);
  MMT_ALIGNED(W1);
  MMT_ALIGNED(score);
  MMT_ALIGNED(work);
  if (work_bytes < sizeof(float) * 2 * (size_t)S * N * He) {
    set_error("mmt_edge_mlp_f32: workspace too small");
    return MMT_EWORKSPACE;
  }
  if (S == 0) return MMT_OK;
  mmt_edge_weights w{W1, b1, W2, b2, w_out, b_out, He};
  return launch_edge_mlp_f32(h, U, adj, &w, S, N, U, score, work, (cudaStream_t)stream);
}
