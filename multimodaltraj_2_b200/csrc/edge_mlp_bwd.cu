// Backward pass of the relational edge MLP of g2k_lstm_mcr (training, BASELINE configs[1]; SURVEY 8f / VERDICT r1 item 8):
// fp32 CUDA-core version (parity mode; any He in {64, 128}).  include/mmt.h: mmt_attention_score_grad_f32,
// mmt_edge_mlp_backward_f32.
//
// Forward (edge_mlp.cu; reference relational_inf_models/nri_learned.py:5-28, models/g2k_lstm_mcr.py:99-124):
//   a = h W1[:U], b = h W1[U:];  e1_ij = elu(a_i + b_j + b1);  e2_ij = elu(e1_ij W2 + b2);  s_ij = sigmoid(w_out.e2_ij + b_out)
//   logits_ij = kern_ij + s_ij on the edges;  att = masked softmax;  [mh | mc]_i = sum_j att_ij [h | c]_j
// Backward, given d loss / d [mh | mc]:
//   G_ij = d[mh|mc]_i . [h|c]_j;  d logit_ij = att_ij (G_ij - sum_k att_ik G_ik)              (attention_score_grad_kernel)
//   du = d logit * s (1 - s);  d e2 = du w_out;  d pre2 = d e2 * elu'(pre2);  d e1 = d pre2 W2^T;  d pre1 = d e1 * elu'(pre1)
//   g w_out += e2^T du, g b_out += sum du, g W2 += e1^T d pre2, g b2 += sum d pre2, g b1 += sum d pre1,
//   d a_i += d pre1_ij,  d b_j += d pre1_ij                                                    (edge_mlp_backward_kernel)
// Everything runs on the edges of the adjacency mask, compacted on the device per scene exactly like the forward kernel: no
// edge list ever goes to the host (the first version of Trainer._edge_backward called .nonzero(): one host sync per frame).
// The node level (g W1 = h^T [da | db], d h = da W1a^T + db W1b^T) is two GEMMs of the caller.
#include "mmt_common.cuh"

namespace mmt {

// ------------------------------------------------------------------------------------------------
// d logit of the attention softmax through the aggregated state.  One warp per agent row.
//   dm[R, C]: d loss / d [mh | mc];  v[R, C]: the aggregated features [h | c];  C % 128 == 0, C <= 512
__global__ void __launch_bounds__(256) attention_score_grad_kernel(const float* __restrict__ att, const uint8_t* __restrict__ adj,
                                                                   const float* __restrict__ dm, const float* __restrict__ v,
                                                                   long R, int N, int C, float* __restrict__ dlog) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const long s = row / N;
  const int nq = C >> 7;                       // float4 per lane
  float4 d4[4];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d4[q] = q < nq ? __ldg(reinterpret_cast<const float4*>(dm + row * C) + lane + 32 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const uint8_t* arow = adj + row * N;
  const float* trow = att + row * N;
  float* orow = dlog + row * N;
  float dot_part = 0.f;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    const bool is = j < N && arow[j] != 0;
    unsigned m = __ballot_sync(0xffffffffu, is);
    float mine = 0.f;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const float4* vj = reinterpret_cast<const float4*>(v + (s * N + j0 + b) * C);
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < nq) {
          const float4 x = __ldg(vj + lane + 32 * q);
          acc = fmaf(d4[q].x, x.x, acc);
          acc = fmaf(d4[q].y, x.y, acc);
          acc = fmaf(d4[q].z, x.z, acc);
          acc = fmaf(d4[q].w, x.w, acc);
        }
      acc = warp_sum(acc);
      if (lane == b) mine = acc;
    }
    if (j < N) {
      orow[j] = mine;                          // G_ij, rewritten below by the same thread
      if (is) dot_part = fmaf(trow[j], mine, dot_part);
    }
  }
  const float dot = warp_sum(dot_part);
  for (int j = lane; j < N; j += 32) {
    const float a = trow[j];
    orow[j] = arow[j] != 0 ? a * (orow[j] - dot) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float elu_b(float x) { return x > 0.f ? x : expm1f(x); }

constexpr int BT = 64;        // edges per tile
constexpr int BCAP = 4096;    // edge-list capacity per row chunk

// na, nb: node projections [R, He];  dscore: d loss / d logit [S,N,N] (read on the edges);  dab[R, 2 He] = [d a | d b] (zeroed
// by the caller, accumulated with red.global);  gW2[He,He], gb1, gb2, gw_out[He], gb_out[1]: accumulated.
// Thread roles per tile: warp w owns edges 8w..8w+7 and lane l the columns l + 32 q (as in the forward kernel) for the
// two per-edge products; for g W2 += e1^T d pre2 thread (ty, tx) = (tid / 16, tid % 16) keeps the (He/16)^2 entries
// (ty + 16 i, tx + 16 j) in registers over ALL tiles of the CTA.
template <int HE>
__global__ void __launch_bounds__(256) edge_mlp_backward_kernel(const float* __restrict__ na, const float* __restrict__ nb,
                                                                const uint8_t* __restrict__ adj, const float* __restrict__ dscore,
                                                                const float* __restrict__ b1, const float* __restrict__ W2,
                                                                const float* __restrict__ b2, const float* __restrict__ w_out,
                                                                const float* __restrict__ b_out, int S, int N,
                                                                float* __restrict__ dab, float* __restrict__ gW2,
                                                                float* __restrict__ gb1, float* __restrict__ gb2,
                                                                float* __restrict__ gw_out, float* __restrict__ gb_out) {
  constexpr int LD = HE + 1;                   // odd row stride: rows and columns both walk the banks
  constexpr int NQ = HE / 32;                  // columns per lane
  constexpr int BK = HE / 16;                  // g W2 block edge per thread
  extern __shared__ __align__(16) float sm[];
  float* sW2 = sm;                             // [HE][LD]
  float* sE1 = sW2 + HE * LD;                  // [BT][LD]
  float* sD2 = sE1 + BT * LD;                  // [BT][LD]  d pre2
  int* sList = reinterpret_cast<int*>(sD2 + BT * LD);   // [BCAP]  (i << 16 | j)
  __shared__ int sCount;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  for (int i = tid; i < HE * HE; i += 256) sW2[(i / HE) * LD + (i % HE)] = __ldg(W2 + i);
  float wo[NQ], bb2[NQ], bb1[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    wo[q] = __ldg(w_out + lane + 32 * q);
    bb2[q] = __ldg(b2 + lane + 32 * q);
    bb1[q] = __ldg(b1 + lane + 32 * q);
  }
  const float bo = __ldg(b_out);
  const int rows_per_chunk = BCAP / N > 0 ? BCAP / N : 1;

  float accW2[BK][BK];
#pragma unroll
  for (int i = 0; i < BK; ++i)
#pragma unroll
    for (int j = 0; j < BK; ++j) accW2[i][j] = 0.f;
  float a_b1[NQ], a_b2[NQ], a_wo[NQ], a_bo = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) a_b1[q] = a_b2[q] = a_wo[q] = 0.f;

  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    const float* ds = dscore + (size_t)s * N * N;
    for (int r0 = 0; r0 < N; r0 += rows_per_chunk) {
      __syncthreads();                         // the previous chunk's list is consumed (and sW2 is staged)
      if (tid == 0) sCount = 0;
      __syncthreads();
      const int r1 = min(N, r0 + rows_per_chunk);
      const int tot = (r1 - r0) * N;
      for (int e0 = 0; e0 < tot; e0 += 256) {
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
      for (int t0 = 0; t0 < ne; t0 += BT) {
        const int nt = min(BT, ne - t0);
        // ---- e1 tile (rows beyond nt: zero)
        for (int idx = tid; idx < BT * HE; idx += 256) {
          const int t = idx / HE, m = idx - t * HE;
          float v = 0.f;
          if (t < nt) {
            const int ij = sList[t0 + t];
            v = elu_b(na[((size_t)s * N + (ij >> 16)) * HE + m] + nb[((size_t)s * N + (ij & 0xffff)) * HE + m] + __ldg(b1 + m));
          }
          sE1[t * LD + m] = v;
        }
        __syncthreads();
        // ---- pre2 = e1 W2 + b2, e2, score, du, d pre2
        {
          float acc[8][NQ];
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[r][q] = 0.f;
          for (int k = 0; k < HE; ++k) {
            float w2[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) w2[q] = sW2[k * LD + lane + 32 * q];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float ev = sE1[(warp * 8 + r) * LD + k];
#pragma unroll
              for (int q = 0; q < NQ; ++q) acc[r][q] = fmaf(ev, w2[q], acc[r][q]);
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int t = warp * 8 + r;
            float e2[NQ], part = 0.f;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              acc[r][q] += bb2[q];
              e2[q] = elu_b(acc[r][q]);
              part = fmaf(e2[q], wo[q], part);
            }
            part = warp_sum(part);
            float du = 0.f;
            if (t < nt) {
              const int ij = sList[t0 + t];
              const float sc = sigmoid_acc(part + bo);
              du = __ldg(ds + (size_t)(ij >> 16) * N + (ij & 0xffff)) * sc * (1.0f - sc);
            }
            a_bo += du;                        // the same value in every lane; lane 0 reports it
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              a_wo[q] = fmaf(du, e2[q], a_wo[q]);
              const float d2 = du * wo[q] * (acc[r][q] > 0.f ? 1.0f : e2[q] + 1.0f);
              a_b2[q] += d2;
              sD2[t * LD + lane + 32 * q] = d2;
            }
          }
        }
        __syncthreads();
        // ---- d pre1 = (d pre2 W2^T) * elu'(pre1);  scatter into d a_i, d b_j
        {
          float acc[8][NQ];
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[r][q] = 0.f;
          for (int c = 0; c < HE; ++c) {
            float w2[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) w2[q] = sW2[(lane + 32 * q) * LD + c];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float dv = sD2[(warp * 8 + r) * LD + c];
#pragma unroll
              for (int q = 0; q < NQ; ++q) acc[r][q] = fmaf(dv, w2[q], acc[r][q]);
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int t = warp * 8 + r;
            if (t < nt) {                      // warp-uniform
              const int ij = sList[t0 + t];
              float* da = dab + ((size_t)s * N + (ij >> 16)) * (2 * HE);
              float* db = dab + ((size_t)s * N + (ij & 0xffff)) * (2 * HE) + HE;
#pragma unroll
              for (int q = 0; q < NQ; ++q) {
                const float e1 = sE1[t * LD + lane + 32 * q];
                const float d1 = acc[r][q] * (e1 > 0.f ? 1.0f : e1 + 1.0f);
                a_b1[q] += d1;
                atomicAdd(da + lane + 32 * q, d1);
                atomicAdd(db + lane + 32 * q, d1);
              }
            }
          }
        }
        // ---- g W2 += e1^T d pre2 over the tile's edges (zero rows beyond nt add nothing)
        for (int e = 0; e < BT; ++e) {
          float av[BK], dv[BK];
#pragma unroll
          for (int i = 0; i < BK; ++i) av[i] = sE1[e * LD + ty + 16 * i];
#pragma unroll
          for (int j = 0; j < BK; ++j) dv[j] = sD2[e * LD + tx + 16 * j];
#pragma unroll
          for (int i = 0; i < BK; ++i)
#pragma unroll
            for (int j = 0; j < BK; ++j) accW2[i][j] = fmaf(av[i], dv[j], accW2[i][j]);
        }
        __syncthreads();                       // sE1 / sD2 are rebuilt by the next tile
      }
    }
  }
  // ---- this CTA's partial sums -> global
#pragma unroll
  for (int i = 0; i < BK; ++i)
#pragma unroll
    for (int j = 0; j < BK; ++j) atomicAdd(gW2 + (size_t)(ty + 16 * i) * HE + tx + 16 * j, accW2[i][j]);
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    atomicAdd(gb1 + lane + 32 * q, a_b1[q]);
    atomicAdd(gb2 + lane + 32 * q, a_b2[q]);
    atomicAdd(gw_out + lane + 32 * q, a_wo[q]);
  }
  if (lane == 0) atomicAdd(gb_out, a_bo);
}

int launch_sgemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K, cudaStream_t stream);

template <int HE>
static int launch_bwd(const float* na, const float* nb, const uint8_t* adj, const float* dscore, const float* b1, const float* W2,
                      const float* b2, const float* w_out, const float* b_out, int S, int N, float* dab, float* gW2, float* gb1,
                      float* gb2, float* gw_out, float* gb_out, cudaStream_t stream) {
  const size_t smem = sizeof(float) * ((size_t)HE * (HE + 1) + 2 * (size_t)BT * (HE + 1)) + sizeof(int) * BCAP;
  static DeviceMask smem_opted[1];
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&edge_mlp_backward_kernel<HE>), (int)smem, &smem_opted[0])) return rc;
  const int grid = S < num_sms() ? S : num_sms();
  edge_mlp_backward_kernel<HE><<<grid, 256, smem, stream>>>(na, nb, adj, dscore, b1, W2, b2, w_out, b_out, S, N, dab, gW2, gb1,
                                                           gb2, gw_out, gb_out);
  count_launch();
  return check_launch("edge_mlp_backward_kernel");
}

}  // namespace mmt

extern "C" int mmt_attention_score_grad_f32(const float* att, const uint8_t* adj, const float* dm, const float* v, int S, int N,
                                            int C, float* dlogit, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && N <= 65535, "need S >= 0, 0 < N < 65536");
  MMT_REQUIRE(C > 0 && C % 128 == 0 && C <= 512, "need C % 128 == 0, C <= 512");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(att && adj && dm && v && dlogit, "all pointers required");
  MMT_ALIGNED(dm);
  MMT_ALIGNED(v);
  const long R = (long)S * N;
  attention_score_grad_kernel<<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(att, adj, dm, v, R, N, C, dlogit);
  count_launch();
  return check_launch("attention_score_grad_kernel");
}

extern "C" int mmt_edge_mlp_backward_f32(const float* h, const uint8_t* adj, const float* dlogit, const float* W1, const float* b1,
                                         const float* W2, const float* b2, const float* w_out, const float* b_out, int S, int N,
                                         int U, int He, float* dab, float* gW2, float* gb1, float* gb2, float* gw_out,
                                         float* gb_out, float* work, size_t work_bytes, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && N % 4 == 0 && N <= 1024, "need 0 < N <= 1024, N % 4 == 0");
  MMT_REQUIRE(U > 0 && U % 16 == 0 && (He == 64 || He == 128), "need U % 16 == 0, He in {64,128}");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(h && adj && dlogit && W1 && b1 && W2 && b2 && w_out && b_out && dab && gW2 && gb1 && gb2 && gw_out && gb_out && work,
              "all pointers required");
  MMT_ALIGNED(h);
  MMT_ALIGNED(W1);
  MMT_ALIGNED(work);
  const long R = (long)S * N;
  if (work_bytes < sizeof(float) * 2 * (size_t)R * He) {
    set_error("mmt_edge_mlp_backward_f32: workspace too small");
    return MMT_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* na = work;
  float* nb = work + R * He;
  if (int rc = launch_sgemm(h, U, W1, He, na, He, (int)R, He, U, st)) return rc;
  if (int rc = launch_sgemm(h, U, W1 + (size_t)U * He, He, nb, He, (int)R, He, U, st)) return rc;
  if (cudaMemsetAsync(dab, 0, sizeof(float) * 2 * (size_t)R * He, st) != cudaSuccess) {
    set_error("mmt_edge_mlp_backward_f32: cudaMemsetAsync failed");
    return MMT_ECUDA;
  }
  return He == 64 ? launch_bwd<64>(na, nb, adj, dlogit, b1, W2, b2, w_out, b_out, S, N, dab, gW2, gb1, gb2, gw_out, gb_out, st)
                  : launch_bwd<128>(na, nb, adj, dlogit, b1, W2, b2, w_out, b_out, S, N, dab, gW2, gb1, gb2, gw_out, gb_out, st);
}
