// The whole hot path: obs T -> pred P rollout + K-sample decode + ADE/FDE (include/mmt.h
// mmt_forecast_f32).  Host-side orchestration of the per-step kernels on the caller's stream;
// no host synchronisation, no allocation -- everything lives in the caller's workspace.
#include "mmt_common.cuh"

namespace mmt {

int launch_aggregate(const float* logits, const float* logits2, const uint8_t* adj, const float* feat, int S, int N,
                     int C, int ld_feat, float* attn, float* out, int ld_out, cudaStream_t stream);
int launch_cell_f32(const float* x, const float* h, const float* c, const float* mh, const float* mc,
                    const uint8_t* valid, const mmt_cell_weights* w, int R, int ld, float* h_out, float* c_out,
                    float* mf_out, int ld_mf, cudaStream_t stream);
int launch_head(const float* mt, int ld, const float* mf, int ld_mf, const uint8_t* valid, const mmt_cell_weights* w, int R,
                const float* cur_pos, float* params_out, int params_stride, float* next_pos, cudaStream_t stream);
int launch_edge_mlp_f32(const float* h, int ld_h, const uint8_t* adj, const mmt_edge_weights* w, int S, int N, int U,
                        float* score, float* work, cudaStream_t stream);
int launch_cell_tc(const float* x, const float* h, const float* c, const float* mh, const float* mc, int ld,
                   const uint8_t* valid, const mmt_cell_weights* w, int R, float* h_out, float* c_out, float* mf_out,
                   int ld_mf, const float* cur_pos, float* params_out, int params_stride, float* next_pos, int x3,
                   cudaStream_t stream);

int launch_decode(const float* params, const float* eps, uint64_t seed, uint64_t agent_offset, const float* last_obs,
                  int lo_stride, const float* gt, int gt_stride, const uint8_t* valid, int A, int P, int K, float* ade,
                  float* fde, int32_t* best_k, float* best_ade, float* best_fde, float* best_traj, float* eps_out,
                  cudaStream_t stream);

int launch_cell_tc_bf16(const float* x, const void* hb, const float* c, const void* mhb, const void* mcb,
                        const uint8_t* valid, const mmt_cell_weights* w, int R, void* hb_out, float* c_out,
                        const float* cur_pos, float* params_out, int params_stride, float* next_pos, int blocked, int f16,
                        cudaStream_t stream);
int launch_graph_aggregate_mma(const float* pos, const uint8_t* valid, const void* hb, const float* c, const float* score,
                               int S, int N,
                               float r2, float inv_2sigma2, void* mhb, void* mcb, int f16, cudaStream_t stream);
int launch_edge_mlp_tc(const float* h, int ld_h, const uint8_t* adj, const void* packed, const mmt_edge_weights* w, int S,
                       int N, float* score, float* nab, int zero_fill, int f16, cudaStream_t stream);
int launch_pack_edge_weights(const float* W1, const float* W2, void* packed, int f16, cudaStream_t stream);
int launch_rollout_tc(const float* pos, const float* vis, const uint8_t* valid, const mmt_cell_weights* w, int S, int N,
                      int T, int P, float r2, float inv_2sigma2, float* params, long long* dbg, int f16, cudaStream_t stream);
int launch_graph_aggregate_blocked(const float* pos, const uint8_t* valid, const void* hb, const float* c, int S, int N,
                                   float r2, float inv_2sigma2, void* mhb, void* mcb, int f16, cudaStream_t stream);
int launch_graph_aggregate_bf16(const float* pos, const uint8_t* valid, const void* hb, const float* c, int S, int N,
                                int U, float r2, float inv_2sigma2, void* mhb, void* mcb, int f16, cudaStream_t stream);

// x = [cur - prev | vis_t]; for observed frames cur is gathered from pos[:, :, t]
__global__ void __launch_bounds__(256) prep_step_kernel(const float* __restrict__ pos, const float* __restrict__ vis,
                                                        int R, int F, int T, int t, float* __restrict__ cur,
                                                        const float* __restrict__ prev, float* __restrict__ x) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float2 c;
  if (t < T) {
    c = __ldg(reinterpret_cast<const float2*>(pos) + (size_t)r * F + t);
    reinterpret_cast<float2*>(cur)[r] = c;
  } else {
    c = reinterpret_cast<const float2*>(cur)[r];
  }
  float2 d = make_float2(0.f, 0.f);
  if (t > 0) {
    const float2 p = reinterpret_cast<const float2*>(prev)[r];
    d = make_float2(__fsub_rn(c.x, p.x), __fsub_rn(c.y, p.y));
  }
  const int tv = t < T ? t : T - 1;
  const float2 v = __ldg(reinterpret_cast<const float2*>(vis) + (size_t)r * T + tv);
  reinterpret_cast<float4*>(x)[r] = make_float4(d.x, d.y, v.x, v.y);
}


struct Workspace {
  float *pbuf[3], *x, *hc[2], *mhc, *mf, *kern, *score, *ework, *params;
  void* epacked;        // bf16 operand images of the edge-MLP weights (relational bf16 modes)
  uint8_t* adj;
  // bf16-state fast path (non-relational bf16 mode): h, mh, mc bf16 [R,U]; c fp32 [R,U]
  void *hb[2], *mhb, *mcb;
  float* cf[2];
  bool fast, blocked;   // blocked: tile-blocked state layout (128 % N == 0, or N % 128 == 0: a scene spans whole tiles)
  size_t bytes;
};

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static Workspace carve(char* base, const mmt_forecast_cfg* cfg, int U, int He) {
  Workspace w;
  const size_t R = (size_t)cfg->S * cfg->N, NN = (size_t)cfg->S * cfg->N * cfg->N;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  // the relational variant takes the bf16-state path when its edge MLP runs on the tensor cores and the graph step is the
  // MMA kernel (blocked layout, N >= 16): edge scores are added to the logits inside graph_aggregate_mma_kernel
  const bool tileable = 128 % cfg->N == 0 || (cfg->N % 128 == 0 && cfg->N <= 1024);
  const bool rel_fast = cfg->relational && U == 128 && He == 128 && tileable && cfg->N >= 16;
  w.fast = (cfg->prec == MMT_PREC_BF16 || cfg->prec == MMT_PREC_BF16_STEPWISE || cfg->prec == MMT_PREC_F16) &&
           (!cfg->relational || rel_fast);
  for (int i = 0; i < 3; ++i) w.pbuf[i] = (float*)take(R * 2 * 4);
  w.x = (float*)take(R * 4 * 4);
  w.blocked = w.fast && tileable;
  if (w.fast) {
    const size_t Rp = (R + 127) / 128 * 128;   // state rows padded to whole 128-row tiles
    for (int i = 0; i < 2; ++i) {
      w.hb[i] = take(Rp * U * 2);
      w.cf[i] = (float*)take(Rp * U * 4);
    }
    w.mhb = take(Rp * U * 2);
    w.mcb = take(Rp * U * 2);
    w.hc[0] = w.hc[1] = w.mhc = w.mf = w.kern = nullptr;
    w.adj = cfg->relational ? (uint8_t*)take(NN) : nullptr;   // the edge kernel walks the adjacency mask
  } else {
    w.hc[0] = (float*)take(R * 2 * U * 4);
    w.hc[1] = (float*)take(R * 2 * U * 4);
    w.mhc = (float*)take(R * 2 * U * 4);
    w.mf = (float*)take(R * U * 4);
    w.kern = (float*)take(NN * 4);
    w.adj = (uint8_t*)take(NN);
  }
  w.score = (float*)take(cfg->relational ? NN * 4 : 0);
  w.ework = (float*)take(cfg->relational ? 2 * R * He * 4 : 0);
  w.epacked = take(cfg->relational && cfg->prec != MMT_PREC_F32 && cfg->prec != MMT_PREC_BF16X3 && U == 128 && He == 128 ? 96 * 1024 : 0);
  w.params = (float*)take(R * cfg->P * 5 * 4);
  w.bytes = off;
  return w;
}

}  // namespace mmt

extern "C" size_t mmt_forecast_workspace_bytes(const mmt_forecast_cfg* cfg, int U, int He) {
  if (!cfg) return 0;
  return mmt::carve(nullptr, cfg, U, He).bytes;
}

extern "C" int mmt_forecast_f32(const float* pos, const float* vis, const uint8_t* valid, const mmt_cell_weights* cw,
                                const mmt_edge_weights* ew, const mmt_forecast_cfg* cfg, const float* eps,
                                float* params, float* ade, float* fde, int32_t* best_k, float* best_ade,
                                float* best_fde, float* best_traj, void* work, size_t work_bytes, void* stream_) {
  using namespace mmt;
  MMT_REQUIRE(pos && vis && valid && cw && cfg && best_k && work, "pos/vis/valid/weights/cfg/best_k/work required");
  MMT_REQUIRE(cfg->S >= 0 && cfg->N > 0 && cfg->N % 4 == 0 && cfg->N <= 1024, "need 0 < N <= 1024, N % 4 == 0");
  MMT_REQUIRE(cfg->T >= 1 && cfg->P >= 1 && cfg->P <= 32 && cfg->K >= 1 && cfg->K <= 32, "need T >= 1, P, K in [1,32]");
  MMT_REQUIRE(cw->U == 128 && cw->E == 64, "cell is built for U = 128, E = 64");
  MMT_REQUIRE(cw->W_h && cw->b_h, "head weights required");
  MMT_REQUIRE(!cfg->relational || (ew && ew->He > 0), "relational mode needs edge weights");
  MMT_REQUIRE(cfg->prec == MMT_PREC_F32 || cfg->prec == MMT_PREC_BF16 || cfg->prec == MMT_PREC_BF16_STEPWISE ||
                  cfg->prec == MMT_PREC_BF16X3 || cfg->prec == MMT_PREC_F16,
              "unknown precision mode");
  const bool f16 = cfg->prec == MMT_PREC_F16;
  MMT_REQUIRE(!f16 || cw->W_packed_f16, "f16 mode needs W_packed_f16");
  MMT_REQUIRE(cfg->prec == MMT_PREC_F32 || cfg->prec == MMT_PREC_BF16X3 || f16 || cw->W_packed_bf16, "bf16 mode needs W_packed_bf16");
  MMT_REQUIRE(cfg->prec != MMT_PREC_BF16X3 || cw->W_packed_bf16x3, "bf16x3 mode needs W_packed_bf16x3");
  MMT_ALIGNED(pos);
  MMT_ALIGNED(vis);
  MMT_ALIGNED(work);
  const int U = cw->U, He = ew ? ew->He : 0;
  Workspace w = carve((char*)work, cfg, U, He);
  if (work_bytes < w.bytes) {
    set_error("mmt_forecast_f32: workspace too small (%zu < %zu)", work_bytes, w.bytes);
    return MMT_EWORKSPACE;
  }
  if (cfg->S == 0) return MMT_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int S = cfg->S, N = cfg->N, T = cfg->T, P = cfg->P, F = T + P;
  const int R = S * N;
  float* par = params ? params : w.params;
  int rc0;

  // bf16, non-relational, whole scenes per 128-row tile: the entire recurrence is one persistent kernel with the
  // state on chip (rollout_tc.cu).  MMT_PREC_BF16_STEPWISE keeps the per-step kernels (A/B checks, other N).
  const bool fused = w.fast && w.blocked && N >= 8 && N <= 128 && (cfg->prec == MMT_PREC_BF16 || f16) && !cfg->relational;
  if (fused) {
    if ((rc0 = launch_rollout_tc(pos, vis, valid, cw, S, N, T, P, cfg->r2, cfg->inv_2sigma2, par, nullptr, f16, stream)))
      return rc0;
  } else if (w.fast) {
    const size_t Rp = ((size_t)R + 127) / 128 * 128;
    cudaMemsetAsync(w.hb[0], 0, Rp * U * 2, stream);
    cudaMemsetAsync(w.cf[0], 0, Rp * U * 4, stream);
  } else {
    cudaMemsetAsync(w.hc[0], 0, (size_t)R * 2 * U * 4, stream);
  }
  int ic = 0, ip = 1, in = 2, hb = 0;
  int rc;
  // relational bf16 modes: the edge MLP runs on the tensor cores (edge_mlp_tc.cu); weights packed once per call
  const bool edge_tc = cfg->relational && cfg->prec != MMT_PREC_F32 && cfg->prec != MMT_PREC_BF16X3 && U == 128 && He == 128;
  if (edge_tc && (rc = launch_pack_edge_weights(ew->W1, ew->W2, w.epacked, f16, stream))) return rc;
  for (int t = 0; t < (fused ? 0 : T + P - 1); ++t) {
    prep_step_kernel<<<(R + 255) / 256, 256, 0, stream>>>(pos, vis, R, F, T, t, w.pbuf[ic], w.pbuf[ip], w.x);
    count_launch();
    if ((rc = check_launch("prep_step_kernel"))) return rc;
    if (w.fast) {
      // bf16 fast path: [pairwise + softmax + aggregation] and [gate GEMM + gates + head], two kernels per step
      const bool emit = t >= T - 1;
      float* po = par + (size_t)(t - (T - 1)) * 5;
      const float* esc = nullptr;
      if (cfg->relational) {
        // adjacency mask only (no kernel matrix) -> node projections from the blocked bf16 state -> edge scores
        if ((rc = mmt_pairwise_adj_f32(w.pbuf[ic], valid, S, N, cfg->r2, cfg->inv_2sigma2, nullptr, w.adj, nullptr, stream)))
          return rc;
        if ((rc = launch_edge_mlp_tc(reinterpret_cast<const float*>(w.hb[hb]), -1, w.adj, w.epacked, ew, S, N, w.score,
                                     w.ework, 0, f16, stream)))
          return rc;
        esc = w.score;
      }
      if (w.blocked && N >= 16)
        rc = launch_graph_aggregate_mma(w.pbuf[ic], valid, w.hb[hb], w.cf[hb], esc, S, N, cfg->r2, cfg->inv_2sigma2, w.mhb,
                                        w.mcb, f16, stream);
      else if (w.blocked)
        rc = launch_graph_aggregate_blocked(w.pbuf[ic], valid, w.hb[hb], w.cf[hb], S, N, cfg->r2, cfg->inv_2sigma2,
                                            w.mhb, w.mcb, f16, stream);
      else
        rc = launch_graph_aggregate_bf16(w.pbuf[ic], valid, w.hb[hb], w.cf[hb], S, N, U, cfg->r2, cfg->inv_2sigma2,
                                         w.mhb, w.mcb, f16, stream);
      if (rc) return rc;
      if ((rc = launch_cell_tc_bf16(w.x, w.hb[hb], w.cf[hb], w.mhb, w.mcb, valid, cw, R, w.hb[hb ^ 1], w.cf[hb ^ 1],
                                    w.pbuf[ic], emit ? po : nullptr, P * 5, emit ? w.pbuf[in] : nullptr,
                                    w.blocked ? 1 : 0, f16, stream)))
        return rc;
      hb ^= 1;
      const int old_p = ip;
      ip = ic;
      if (emit) {
        ic = in;
        in = old_p;
      } else {
        ic = old_p;
      }
      continue;
    }
    if ((rc = mmt_pairwise_adj_f32(w.pbuf[ic], valid, S, N, cfg->r2, cfg->inv_2sigma2, w.kern, w.adj, nullptr, stream)))
      return rc;
    const float* l2 = nullptr;
    if (cfg->relational) {
      // (the aggregation reads the scores only on the edges: no zero fill needed here)
      if (edge_tc) rc = launch_edge_mlp_tc(w.hc[hb], 2 * U, w.adj, w.epacked, ew, S, N, w.score, w.ework, 0, f16, stream);
      else rc = launch_edge_mlp_f32(w.hc[hb], 2 * U, w.adj, ew, S, N, U, w.score, w.ework, stream);
      if (rc) return rc;
      l2 = w.score;
    }
    if ((rc = launch_aggregate(w.kern, l2, w.adj, w.hc[hb], S, N, 2 * U, 2 * U, nullptr, w.mhc, 2 * U, stream)))
      return rc;
    const bool emit = t >= T - 1;
    float* po = par + (size_t)(t - (T - 1)) * 5;
    if (cfg->prec == MMT_PREC_F32) {
      const float* hc = w.hc[hb];
      float* hco = w.hc[hb ^ 1];
      if ((rc = launch_cell_f32(w.x, hc, hc + U, w.mhc, w.mhc + U, valid, cw, R, 2 * U, hco, hco + U, w.mf, U, stream)))
        return rc;
      if (emit && (rc = launch_head(hco, 2 * U, w.mf, U, valid, cw, R, w.pbuf[ic], po, P * 5, w.pbuf[in], stream)))
        return rc;
    } else {
      const float* hc = w.hc[hb];
      float* hco = w.hc[hb ^ 1];
      if ((rc = launch_cell_tc(w.x, hc, hc + U, w.mhc, w.mhc + U, 2 * U, valid, cw, R, hco, hco + U, nullptr, 0,
                               w.pbuf[ic], emit ? po : nullptr, P * 5, emit ? w.pbuf[in] : nullptr,
                               cfg->prec == MMT_PREC_BF16X3 ? 1 : f16 ? 2 : 0, stream)))
        return rc;
    }
    hb ^= 1;
    // rotate position buffers: prev <- cur; cur <- next (predicted) or a fresh buffer (observed)
    const int old_p = ip;
    ip = ic;
    if (emit) {
      ic = in;
      in = old_p;
    } else {
      ic = old_p;
    }
  }
  // the fused epilogue reads the last observed point and the ground truth straight out of pos[S,N,F,2]
  MMT_ALIGNED(eps);
  return launch_decode(par, eps, cfg->seed, cfg->agent_offset, pos + (size_t)(T - 1) * 2, 2 * F, pos + (size_t)T * 2, 2 * F,
                       valid, R, P, cfg->K, ade, fde, best_k, best_ade, best_fde, best_traj, nullptr, stream);
}

extern "C" int mmt_rollout_bf16(const float* pos, const float* vis, const uint8_t* valid, const mmt_cell_weights* cw,
                                int S, int N, int T, int P, float r2, float inv_2sigma2, float* params,
                                int64_t* timeline, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N >= 8 && N <= 128 && 128 % N == 0, "fused rollout needs N in {8,16,32,64,128}");
  MMT_REQUIRE(T >= 1 && P >= 1, "need T >= 1, P >= 1");
  if (S == 0) return MMT_OK;   // empty batch: nothing to read or write (the buffers may be NULL)
  MMT_REQUIRE(pos && vis && valid && cw && params, "pos/vis/valid/weights/params required");
  MMT_REQUIRE(cw->U == 128 && cw->E == 64 && cw->W_h && cw->b_h && cw->W_packed_bf16,
              "needs U = 128, E = 64, head weights and W_packed_bf16");
  MMT_ALIGNED(pos);
  MMT_ALIGNED(vis);
  MMT_ALIGNED(params);
  if (S == 0) return MMT_OK;
  return launch_rollout_tc(pos, vis, valid, cw, S, N, T, P, r2, inv_2sigma2, params,
                           reinterpret_cast<long long*>(timeline), 0, (cudaStream_t)stream);
}

extern "C" int mmt_rollout_f16(const float* pos, const float* vis, const uint8_t* valid, const mmt_cell_weights* cw,
                               int S, int N, int T, int P, float r2, float inv_2sigma2, float* params,
                               int64_t* timeline, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N >= 8 && N <= 128 && 128 % N == 0, "fused rollout needs N in {8,16,32,64,128}");
  MMT_REQUIRE(T >= 1 && P >= 1, "need T >= 1, P >= 1");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(pos && vis && valid && cw && params, "pos/vis/valid/weights/params required");
  MMT_REQUIRE(cw->U == 128 && cw->E == 64 && cw->W_h && cw->b_h && cw->W_packed_f16,
              "needs U = 128, E = 64, head weights and W_packed_f16");
  MMT_ALIGNED(pos);
  MMT_ALIGNED(vis);
  MMT_ALIGNED(params);
  return launch_rollout_tc(pos, vis, valid, cw, S, N, T, P, r2, inv_2sigma2, params,
                           reinterpret_cast<long long*>(timeline), 1, (cudaStream_t)stream);
}
