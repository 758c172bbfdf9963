/* libmmt -- C-ABI of the B200-native multimodaltraj forecasting hot path.
 *
 * Drop-in boundary for serenetech90/multimodaltraj_2's per-timestep g2k_lstm_mc / g2k_lstm_mcr
 * step.  The reference is pure Python/TensorFlow-1.14 and has no FFI; each entry
 * point below names the reference lines whose arithmetic it replaces (paths relative to the
 * reference root).  Conventions:
 *   - every pointer is a DEVICE pointer to caller-owned, contiguous, row-major memory, 16-byte
 *     aligned; the library never allocates or frees device memory and keeps no global state
 *     except a thread-local error string;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the device;
 *   - return 0 on success; <0 on error: -1 bad argument/shape, -2 misaligned pointer,
 *     -3 workspace too small, -4 CUDA error (text via mmt_last_error()), -5 NCCL error;
 *   - nothing throws or exits across the ABI.
 * Symbols: S scenes, N agents per scene (padded; valid[S,N]), T obs frames, P pred frames,
 * U hidden units (128), E embedding (64), K samples, D = neighborhood_size/grid_size.
 */
#ifndef MMT_H_
#define MMT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMT_VERSION 100
#define MMT_OK 0
#define MMT_EARG (-1)
#define MMT_EALIGN (-2)
#define MMT_EWORKSPACE (-3)
#define MMT_ECUDA (-4)
#define MMT_ENCCL (-5)

/* precision modes of the gate / aggregation contractions */
#define MMT_PREC_F32 0   /* fp32 CUDA-core FMA, parity mode (1e-4 rel vs oracle)             */
#define MMT_PREC_BF16 1  /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM   */
#define MMT_PREC_BF16_STEPWISE 2 /* mmt_forecast_f32 only: bf16 as above but one kernel pair per step
                                    (state in HBM) instead of the fused persistent rollout           */
#define MMT_PREC_BF16X3 3 /* tolerance-meeting tensor-core mode: fp32 state and aggregation, gate GEMM as split bf16
                             (a_hi w_hi + a_lo w_hi + a_hi w_lo, fp32 accumulate in TMEM: 16 operand mantissa bits),
                             tanhf / expf in the epilogue -- the north_star's 1e-4 / 1e-3 tolerance on tcgen05      */

#define MMT_PREC_F16 4 /* the kernels of MMT_PREC_BF16 with fp16 operands and fp16 state words (10 stored mantissa bits
                          instead of 7; fp32 accumulate in TMEM): same kernels and speed, ADE / FDE inside the 1e-3 bar that
                          bf16 misses.  Needs W_packed_f16; operands must stay inside the fp16 range (|x| < 65504)          */

int mmt_version(void);
const char* mmt_last_error(void);
/* SM count of the current device (grids are sized from it; queried once per device). */
int mmt_num_sms(void);
/* Diagnostics of a device-side trap.  The tcgen05 kernels bound every mbarrier wait and check the alignment of
 * their shared-memory base; a violated bound ends the launch with cudaErrorLaunchFailure (719), which by itself
 * names nothing.  Before trapping, the kernel writes (site, CTA, thread, barrier address, parity) into a block
 * of host-mapped pinned memory owned by the library (64 bytes, allocated at the first tensor-core launch, the one
 * exception to "never allocates": it is HOST memory and survives the dead context).  Returns 1 and fills out[8]
 * = {site (kernel << 8 | wait), CTA, thread, barrier smem address, parity, extra, 0, 0} if a trap was recorded,
 * else 0.  The same text is appended to mmt_last_error() of the first failing launch check after the trap. */
int mmt_last_trap(uint32_t out[8]);

/* ---- pairwise distance kernel + adjacency ------------------------------------------------
 * Replaces networkx_graph.py:71 (dist_mat, never filled) and :83-85 (L2 edge norm) with the
 * N x N kernel the north_star names.  Per scene-frame s:
 *   d2 = (xi-xj)^2 + (yi-yj)^2 (fp32, no FMA); adj = d2 < r2 && i != j && valid_i && valid_j;
 *   kern = adj ? exp(-d2 * inv_2sigma2) : 0; deg_i = sum_j adj_ij.
 * pos[S,N,2] f32, valid[S,N] u8; outputs kern[S,N,N] f32 (or NULL), adj[S,N,N] u8 (or NULL),
 * deg[S,N] i32 (or NULL).  N % 4 == 0, N <= 1024. */
int mmt_pairwise_adj_f32(const float* pos, const uint8_t* valid, int S, int N, float r2,
                         float inv_2sigma2, float* kern, uint8_t* adj, int32_t* deg, void* stream);

/* Neighbour index lists from the adjacency mask: nbr[S,N,max_nbr] i32 ascending j, padded -1;
 * cnt[S,N] = min(deg, max_nbr).  (north_star: "neighbour indexing ... bit-exact") */
int mmt_neighbor_index_i32(const uint8_t* adj, int S, int N, int max_nbr, int32_t* nbr,
                           int32_t* cnt, void* stream);

/* ---- graph aggregation --------------------------------------------------------------------
 * Generalises train.py:240-247 (attn = row softmax; Hs = attn @ Hs).  a = softmax over
 * {j : adj_ij} of logits_ij (rows without neighbours -> 0); out = a @ feat.
 * logits[S,N,N] f32, adj[S,N,N] u8, feat[S,N,C] f32 -> attn[S,N,N] (or NULL), out[S,N,C]. */
int mmt_aggregate_f32(const float* logits, const uint8_t* adj, const float* feat, int S, int N,
                      int C, float* attn, float* out, void* stream);

/* Adjoint of the aggregation (backward pass of the training step): out[S,N,C] = attn^T d per scene,
 * out_j = sum_i attn_ij d_i, attn[S,N,N] as written by mmt_aggregate_f32 (zero off the edges), d[S,N,C].
 * Deterministic (a gather in ascending i, no atomics). */
int mmt_aggregate_transpose_f32(const float* attn, const float* d, int S, int N, int C, float* out, void* stream);

/* ---- relational edge MLP (g2k_lstm_mcr only) ------------------------------------------------
 * Replaces relational_inf_models/nri_learned.py:5-28 (infer_rlns sigmoid gate) with the fNRI
 * node2edge -> 2-layer ELU MLP -> score it stubs out.
 *   score_ij = adj_ij ? sigmoid(w_out . elu(W2^T elu(W1a^T h_i + W1b^T h_j + b1) + b2) + b_out) : 0
 * h[S,N,U]; W1[2U,He] b1[He] W2[He,He] b2[He] w_out[He] b_out[1] -> score[S,N,N] f32.
 * work: >= 2*S*N*He floats. */
int mmt_edge_mlp_f32(const float* h, const uint8_t* adj, const float* W1, const float* b1,
                     const float* W2, const float* b2, const float* w_out, const float* b_out,
                     int S, int N, int U, int He, float* score, float* work, size_t work_bytes,
                     void* stream);

/* The same edge MLP on the tcgen05 tensor cores (bf16 operands, fp32 accumulation; U = He = 128): node projections
 * [a | b] = h [W1a | W1b] as one GEMM, then per tile of 128 EDGES of the adjacency mask
 * e1 = elu(a_i + b_j + b1) -> e2 = elu(e1 W2 + b2) (MMA) -> sigmoid(w_out . e2 + b_out).  `packed` comes from
 * mmt_pack_edge_weights_bf16 (mmt_edge_weights_packed_bytes bytes).  work: >= 2*S*N*He floats.
 * Replaces relational_inf_models/nri_learned.py:5-28 like mmt_edge_mlp_f32; tolerance 2e-2 vs the fp32 oracle. */
size_t mmt_edge_weights_packed_bytes(int U, int He);
int mmt_pack_edge_weights_bf16(const float* W1, const float* W2, int U, int He, void* packed, void* stream);
int mmt_edge_mlp_bf16(const float* h, const uint8_t* adj, const void* packed, const float* b1, const float* b2,
                      const float* w_out, const float* b_out, int S, int N, int U, int He, float* score,
                      float* work, size_t work_bytes, void* stream);

/* Backward pass of the relational scores (training of g2k_lstm_mcr, BASELINE configs[1]; the reference differentiates
 * models/g2k_lstm_mcr.py:99-124 / relational_inf_models/nri_learned.py:5-28 through TF autodiff, train.py:240-254).
 * Everything runs on the edges of adj, compacted on the device: no edge list goes to the host.
 *
 * mmt_attention_score_grad_f32: gradient of the loss w.r.t. the attention LOGITS through the aggregated state,
 *   G_ij = dm_i . v_j,  dlogit_ij = attn_ij (G_ij - sum_k attn_ik G_ik) on the edges, 0 elsewhere;
 *   attn[S,N,N] as written by mmt_aggregate_f32, dm[S*N,C] = d loss / d [mh | mc], v[S*N,C] = [h | c]; C % 128 == 0, C <= 512.
 * mmt_edge_mlp_backward_f32 (fp32 CUDA cores, He in {64,128}) / _bf16 (tcgen05, bf16 operands, U = He = 128; tolerance
 *   2e-2 of the largest gradient entry): from dlogit (the scores are added to the logits), per edge
 *   du = dlogit s (1 - s), d pre2 = du w_out elu'(pre2), d pre1 = (d pre2 W2^T) elu'(pre1); outputs
 *   dab[S*N, 2 He] = [sum_j d pre1_ij | sum_i d pre1_ij] (overwritten), and ACCUMULATED into: gW2[He,He] += e1^T d pre2,
 *   gb2 += sum d pre2, gw_out += e2^T du, gb_out += sum du, gb1 += sum d pre1 (fp32 version only: the column sums of
 *   dab[:, :He] give the same).  The node level (g W1 = h^T [da | db], d h = da W1a^T + db W1b^T) is two GEMMs of the
 *   caller (mmt_gemm_tf32).  work: >= 2*S*N*He floats. */
int mmt_attention_score_grad_f32(const float* attn, const uint8_t* adj, const float* dm, const float* v, int S, int N, int C,
                                 float* dlogit, void* stream);
int mmt_edge_mlp_backward_f32(const float* h, const uint8_t* adj, const float* dlogit, const float* W1, const float* b1,
                              const float* W2, const float* b2, const float* w_out, const float* b_out, int S, int N, int U,
                              int He, float* dab, float* gW2, float* gb1, float* gb2, float* gw_out, float* gb_out,
                              float* work, size_t work_bytes, void* stream);
int mmt_edge_mlp_backward_bf16(const float* h, const uint8_t* adj, const float* dlogit, const void* packed, const float* b1,
                               const float* b2, const float* w_out, const float* b_out, int S, int N, int U, int He,
                               float* dab, float* gW2, float* gb2, float* gw_out, float* gb_out, float* work,
                               size_t work_bytes, void* stream);

/* ---- gsk_lstm_cell: fused gate update ---------------------------------------------------------
 * Replaces models/gsk_lstm_cell.py:4-65 (dead Hadamard stub) with the GridLSTMCell gate
 * equations of helper.py:31-39 at U units over the graph neighbourhood:
 *   e = relu(x W_e + b_e); z = [e|h|mh] W + b; (i,j,o) = split(z)
 *   g = sig(i + w_If*mc + w_It*c); c_f = (1-g)mc + g tanh j; c_t = (1-g)c + g tanh j
 *   q = sig(o + w_Of*c_f + w_Ot*c_t); m_f = q tanh c_f; m_t = q tanh c_t
 * x[R,4] h,c,mh,mc[R,U] valid[R] (R = S*N rows) -> h_out=m_t, c_out=c_t, mf_out=m_f [R,U].
 * If W_h != NULL also the head: y = [m_t|m_f] W_h + b_h -> params_out[R*params_stride .. +5]
 * = (mu_x, mu_y, exp(.), exp(.), tanh(.)) and next_pos[R,2] = cur_pos + (mu_x, mu_y).
 * prec = MMT_PREC_F32 (W fp32 [E+2U,3U]) or MMT_PREC_BF16 (W_packed from mmt_pack_gate_weights_bf16).
 * h_out / c_out must not alias h / c (other CTAs re-read the input rows). */
typedef struct mmt_cell_weights {
  const float* W_e;   /* [4,E]      */
  const float* b_e;   /* [E]        */
  const float* W;     /* [E+2U,3U] fp32 row-major                                     */
  const float* b;     /* [3U]       */
  const float* w_If;  /* [U] peephole diagonals                                        */
  const float* w_It;
  const float* w_Of;
  const float* w_Ot;
  const float* W_h;   /* [2U,5] or NULL */
  const float* b_h;   /* [5]   or NULL */
  const void* W_packed_bf16; /* tcgen05 operand image of W (mmt_pack_gate_weights_bf16) or NULL */
  int E;
  int U;
  const void* W_packed_bf16x3; /* split-bf16 image [W_hi ; W_hi ; W_lo] (mmt_pack_gate_weights_bf16x3) or NULL */
  const void* W_packed_f16;    /* the W_packed_bf16 image with fp16 entries (mmt_pack_gate_weights_f16) or NULL */
} mmt_cell_weights;

int mmt_gsk_cell(const float* x, const float* h, const float* c, const float* mh, const float* mc,
                 const uint8_t* valid, const mmt_cell_weights* w, int R, int prec, float* h_out,
                 float* c_out, float* mf_out, const float* cur_pos, float* params_out,
                 int params_stride, float* next_pos, void* stream);

/* bytes of the packed bf16 operand image for W[E+2U,3U] */
size_t mmt_gate_weights_packed_bytes(int E, int U);
/* the split-bf16 image of MMT_PREC_BF16X3 (three times the size) */
size_t mmt_gate_weights_packed_x3_bytes(int E, int U);
int mmt_pack_gate_weights_bf16x3(const float* W, int E, int U, void* packed, void* stream);
int mmt_pack_gate_weights_bf16(const float* W, int E, int U, void* packed, void* stream);
/* fp16 entries instead of bf16 (MMT_PREC_F16); mmt_gate_weights_packed_bytes bytes */
int mmt_pack_gate_weights_f16(const float* W, int E, int U, void* packed, void* stream);

/* ---- GridLSTMCell exactly as helper.py:31-39 / :131-139 instantiate it -------
 * inputs[B, in_stride] (first 4F columns used), state[B, st_stride] (first 2UF used),
 * W_f[4+2U,3U], B_f[3U], peephole diagonals [U] (ignored when peepholes == 0)
 * -> m_out[B,2UF], state_out[B,2UF].  U <= 16. */
int mmt_gridlstm_step_f32(const float* inputs, int in_stride, const float* state, int st_stride,
                          const float* W_f, const float* B_f, const float* w_If, const float* w_It,
                          const float* w_Of, const float* w_Ot, int B, int U, int F, int peepholes,
                          float* m_out, float* state_out, void* stream);

/* ---- Track-A batched g2k_lstm_mcr step -----------------------------------------------------------
 * Replaces models/g2k_lstm_mcr.py:99-124 + train.py:178-183,194-195,240-254 for S scenes at once:
 *   I = W_ii (X W_i); vemb = V W_i; vrel = vemb_prev * vemb; outputs = [I; vemb]
 *   ngh = (lam C) stat_mask; ngh' = lam ngh; Eo = W_v outputs + b_v
 *   attn = ngh' (Eo * (W_r vrel)); cost = Eo ngh'; band = reshape((W_c cost) W_o, (2,P,n))
 *   a = softmax(exp(attn)/cumsum0(exp(attn))); Hs = a softmax(Hs); adj = rowsum(softmax(Hs)); Hs *= adj
 * variant 0 = g2k_lstm_mcr, 1 = g2k_lstm_mc (cost := 0 -> band == 0, models/g2k_lstm_mc.py:59-69).
 * Per scene: X[T,n] V[2,n] C[D,D] Hs[D,H] vemb_prev[2,D] (or NULL: first step).
 * Shared: W_i[n,D] W_ii[D,T] W_v[T,D+2] b_v[D] W_r[T,2] W_c[2P,T] W_o[T,n].
 * Outputs per scene: attn[D,D] cost[T,T] band[2,P,n] Hs_out[D,H] adj[D] vemb_out[2,D] (any may be NULL
 * except Hs_out).  D <= 16, T <= 16, n <= 64, H <= 128, 2P <= 32. */
typedef struct mmt_mcr_weights {
  const float *W_i, *W_ii, *W_v, *b_v, *W_r, *W_c, *W_o;
} mmt_mcr_weights;

int mmt_mcr_step_f32(const float* X, const float* V, const float* C, const float* Hs,
                     const float* vemb_prev, const mmt_mcr_weights* w, int S, int n, int D, int T,
                     int P, int H, float lam, int variant, float* attn, float* cost, float* band,
                     float* Hs_out, float* adj, float* vemb_out, void* stream);

/* g2k_lstm_mcr.forward alone (models/g2k_lstm_mcr.py:99-124) on caller-built placeholders, S scenes:
 * outputs[S,D+2,D] rel[S,2,D] ngh[S,D,T] -> attn[S,D,D] cost[S,T,T] band[S,2,P,n] (any may be NULL).
 * variant 1 = g2k_lstm_mc (cost := 0). */
int mmt_mcr_forward_f32(const float* outputs, const float* rel, const float* ngh, const mmt_mcr_weights* w,
                        int S, int n, int D, int T, int P, float lam, int variant, float* attn, float* cost,
                        float* band, void* stream);

/* ---- reference-compatible scores -----------------------------------------------------------------------
 * mmt_mean_error_f32: sample.get_mean_error (sample.py:21-82) value for value.  predicted/truth [n,L,2]
 * agent-major; out3 = (ade, fde, counter): signed errors are summed over the first maxNumPeds agents per
 * step before the norm (reference defect F-8), steps observed_length..L-1 only.
 * mmt_train_val_scores_f32: train.py:639-674 per agent: euc_i = sigma_max(pred_i[:L_i]-tgt_i[:L_i])/12
 * (np.linalg.norm(M, ord=2) of a matrix is the spectral norm; short tracks also / n_targets),
 * err_i = last row of the difference.  pred/tgt [n,P,2], len[n]. */
int mmt_mean_error_f32(const float* predicted, const float* truth, int n, int L, int observed_length,
                       int maxNumPeds, float* out3, void* stream);
int mmt_train_val_scores_f32(const float* pred, const float* tgt, const int32_t* len, int n, int P,
                             int n_targets, float* euc, float* err, void* stream);

/* ADE / FDE in metres (data/eth/univ/getPixelCoordinates.m:8-30): positions are pixel coordinates
 * divided by (scale0, scale1) = (480, 640); world = Hm [pos0*scale0, pos1*scale1, 1]^T, divided by its third
 * component.  pred/gt [n,P,2], valid[n] u8 or NULL, Hm[9] row-major (device).  ade[n], fde[n] (0 for invalid
 * agents); sums[3] or NULL = (sum ade, sum fde, number of valid agents). */
int mmt_ade_fde_world_f32(const float* pred, const float* gt, const uint8_t* valid, int n, int P, const float* Hm,
                          float scale0, float scale1, float* ade, float* fde, float* sums, void* stream);

/* ---- static-context branch ------------------------------------------------------------
 * Replaces train.py:93-110 (scene image x one huge random filter -> _2dconv[D,D]) and train.py:154-158
 * (stat_mask, _2dconv_in = _2dconv x stat_mask -> the ngh[D,T] input of g2k_lstm_mc(r)).
 * img[H,W,C] f32; the image is zero-padded by one row on top and bottom and one column on the right
 * (train.py:97-99); filt[FH,FW,C] with FH = H+2-D+1, FW = W+1-D+1 (the reference draws it unseeded, so it is an
 * input); conv[D,D] = lam * VALID-correlation; ngh[D,T][i,t] = (sum_j conv[i,j]) * t/T.
 * Workspace: mmt_static_context_workspace_bytes(H, D). */
size_t mmt_static_context_workspace_bytes(int H, int D);
int mmt_static_context_f32(const float* img, int H, int W, int C, const float* filt, int D, int T, float lam,
                           float* conv, float* ngh, void* workspace, size_t workspace_bytes, void* stream);

/* nri_learned.infer_rlns (sigmoid, nri_learned.py:16-21) and eval_rln_ngh (row softmax, :23-28) */
int mmt_sigmoid_f32(const float* x, float* y, size_t n, void* stream);
int mmt_rowsoftmax_f32(const float* x, float* y, int rows, int cols, void* stream);

/* ---- K-sample bivariate-Gaussian decode + ADE/FDE + best-of-K (one fused epilogue) ---------------
 * Replaces the scoring of train.py:639-674 / sample.py:21-82 (which score tf.random_normal,
 * SURVEY F4) with the decode the north_star names.  params[S,N,P,5] activated
 * (mu_x,mu_y,sig_x,sig_y,rho); eps[S,N,K,P,2] or NULL (then Philox4x32-10 keyed
 * (seed, agent_offset + s*N+i, k, t) + Box-Muller in-kernel); last_obs[S,N,2]; gt[S,N,P,2];
 * valid[S,N] -> ade[S,N,K], fde[S,N,K] (either may be NULL), best_k[S,N] i32 (argmin ADE, ties ->
 * lowest k, -1 invalid), best_ade[S,N], best_fde[S,N] (may be NULL), best_traj[S,N,P,2] (may be NULL).
 * P <= 32, K <= 32. */
int mmt_decode_score_f32(const float* params, const float* eps, uint64_t seed, uint64_t agent_offset,
                         const float* last_obs, const float* gt, const uint8_t* valid, int S, int N,
                         int P, int K, float* ade, float* fde, int32_t* best_k, float* best_ade,
                         float* best_fde, float* best_traj, void* stream);

/* Diagnostic variant of the above in Philox mode that also writes the noise it used to
 * eps_out[S,N,K,P,2] (tests compare it with the oracle's Philox/Box-Muller restatement). */
int mmt_decode_score_dump_eps_f32(const float* params, uint64_t seed, uint64_t agent_offset,
                                  const float* last_obs, const float* gt, const uint8_t* valid, int S, int N,
                                  int P, int K, int32_t* best_k, float* best_ade, float* eps_out, void* stream);

/* ---- device-side padded scene batching ---------------------------------------------------------------
 * Replaces load_traj.py:153-224 (next_step dict walking) + networkx_graph.py:30-73,114-129
 * (ConstructGraph / setNodes) for window extraction: rows sorted by (frame, ped) as
 * frame_id[M] i32, ped_id[M] i32, xy[M,2] f32 (+ vis[M,2] or NULL); windows start at
 * win_start[S] (frame ids) and span F frames with stride `fstride`.  A pedestrian is given a
 * slot (ascending ped id order of first appearance in the window's first frame) if present in
 * ALL F frames and the scene has < N slots used.  Outputs pos[S,N,F,2], visout[S,N,F,2] (or NULL),
 * valid[S,N] u8, ped_of_slot[S,N] i32 (-1 empty).  frame_row_start[n_frames+1] is the CSR offset of
 * each distinct frame id in frame_ids_sorted[n_frames]. */
int mmt_scene_batch_f32(const int32_t* frame_ids_sorted, const int32_t* frame_row_start, int n_frames,
                        const int32_t* ped_id, const float* xy, const float* vis,
                        const int32_t* win_start, int S, int N, int F, int fstride, float* pos,
                        float* visout, uint8_t* valid, int32_t* ped_of_slot, void* stream);

/* ---- the whole hot path: obs T -> pred P rollout + decode + score -------------------------------------
 * One call = one batch of scenes through T+P-1 cell steps (pairwise -> [edge MLP] -> aggregation ->
 * gate update -> head), then decode/score.  pos[S,N,T+P,2], vis[S,N,T,2], valid[S,N].
 * relational = 0 (g2k_lstm_mc: logits = kern) or 1 (g2k_lstm_mcr: logits = kern + edge score).
 * Outputs as mmt_decode_score_f32 plus params[S,N,P,5] (may be NULL -> uses workspace).
 * Workspace: mmt_forecast_workspace_bytes(). */
typedef struct mmt_edge_weights {
  const float *W1, *b1, *W2, *b2, *w_out, *b_out;
  int He;
} mmt_edge_weights;

typedef struct mmt_forecast_cfg {
  int S, N, T, P, K;
  float r2, inv_2sigma2;
  int relational;
  int prec;            /* MMT_PREC_* */
  uint64_t seed;       /* Philox seed when eps == NULL */
  uint64_t agent_offset;
} mmt_forecast_cfg;

size_t mmt_forecast_workspace_bytes(const mmt_forecast_cfg* cfg, int U, int He);

int mmt_forecast_f32(const float* pos, const float* vis, const uint8_t* valid,
                     const mmt_cell_weights* cw, const mmt_edge_weights* ew, const mmt_forecast_cfg* cfg,
                     const float* eps, float* params, float* ade, float* fde, int32_t* best_k,
                     float* best_ade, float* best_fde, float* best_traj, void* work, size_t work_bytes,
                     void* stream);

/* The recurrence alone in bf16 mode as ONE persistent kernel with the state on chip (what mmt_forecast_f32
 * runs for g2k_lstm_mc when prec = MMT_PREC_BF16 and 128 % N == 0): per step pairwise kernel + adjacency +
 * masked softmax -> aggregation (tcgen05) -> gate GEMM (tcgen05) -> gate update -> head -> next position.
 * Replaces the per-frame loop of train.py:161-254 for a whole batch.  pos[S,N,T+P,2] (only the T observed
 * frames are read), vis[S,N,T,2], valid[S,N] -> params[S,N,P,5].  N in {8,16,32,64,128}.
 * timeline: NULL, or [64][32] int64 clock64() phase stamps of CTA 0 (diagnostics). */
int mmt_rollout_bf16(const float* pos, const float* vis, const uint8_t* valid, const mmt_cell_weights* cw,
                     int S, int N, int T, int P, float r2, float inv_2sigma2, float* params,
                     int64_t* timeline, void* stream);
/* The same kernel with fp16 instead of bf16 operands (MMT_PREC_F16; cw->W_packed_f16): 10 stored mantissa bits in the
 * gate-GEMM operands, the state images and the attention numerators, fp32 accumulation; same speed, and inside the
 * 1e-3 bar on ADE / FDE that bf16 misses (measured: profiles/, bench.py modes{}).  Values must stay below 65504. */
int mmt_rollout_f16(const float* pos, const float* vis, const uint8_t* valid, const mmt_cell_weights* cw,
                    int S, int N, int T, int P, float r2, float inv_2sigma2, float* params,
                    int64_t* timeline, void* stream);

/* Forward gate update from pre-activations z[R,3U] = [e|h|mh] W + b (columns i | j | o) computed by a library GEMM:
 * the gate equations of helper.py:31-39 -> h'[R,U], c'[R,U], m_f[R,U]; invalid rows give zeros.  Used by
 * the training forward (Trainer), which keeps z for mmt_gsk_cell_backward_f32 instead of recomputing it. */
int mmt_gsk_gates_f32(const float* z, const float* c, const float* mc, const uint8_t* valid, const float* w_If,
                      const float* w_It, const float* w_Of, const float* w_Ot, int R, int U, float* h_out,
                      float* c_out, float* mf_out, void* stream);

/* ---- training step (SURVEY App. C.5 "Training loss (fills F2)", section 8e) ------------------------------------
 * The reference has no loss / optimiser (train.py:23-366 logs raw errors); the loss defined for it is the teacher-
 * forced mean bivariate-Gaussian NLL of the next displacement.  These two kernels are the element-wise work of one
 * step of back-propagation through time; multimodaltraj_2_b200/train.py (Trainer) orchestrates them with the forward
 * kernels above, library GEMMs for the weight / input gradients and one NCCL all-reduce of the flat gradient bucket.
 *
 * mmt_head_nll_f32: y = [m_t | m_f] W_h + b_h (raw head outputs), nll of target[R,2] under mu = y0:2,
 * sigma = exp(y2:4), rho = tanh(y4); loss_sum[0] += scale * sum over valid rows; dy[R,5] = scale * d nll / d y
 * (0 for invalid rows).
 * mmt_gsk_cell_backward_f32: backward of mmt_gsk_cell given the pre-activations z[R,3U] = [e|h|mh] W + b, the
 * previous cell state c, mc, and the upstream gradients d_mt (w.r.t. h' = m_t), d_mf, d_ct (either may be NULL = 0):
 * dz[R,3U], dc[R,U] (w.r.t. the previous c), dmc[R,U]; dpeep[4,U] += gradients of (w_If, w_It, w_Of, w_Ot);
 * db[3U] += the gate-bias gradient (column sums of dz), fused so that no separate pass re-reads dz. */
int mmt_head_nll_f32(const float* m_t, const float* m_f, const uint8_t* valid, const float* W_h, const float* b_h,
                     const float* target, int R, int U, float scale, float* loss_sum, float* dy, void* stream);
int mmt_gsk_cell_backward_f32(const float* z, const float* c, const float* mc, const uint8_t* valid,
                              const float* w_If, const float* w_It, const float* w_Of, const float* w_Ot,
                              const float* d_mt, const float* d_mf, const float* d_ct, int R, int U, float* dz,
                              float* dc, float* dmc, float* dpeep, float* db /* [3U] += column sums of dz, or NULL */,
                              void* stream);

/* Packed training path (Trainer(gemm="tc")): the same two kernels on PACKED rows, and the element-wise glue between this
 * library's kernels of one frame of the teacher-forced step (what the per-frame loop of train.py:161-254 does in Python /
 * TensorFlow ops, plus its adjoint), so that no library slicing / concatenation / broadcast kernel runs between them.
 *   hc[R,2U] = [h | c], mhc[R,2U] = [mh | mc] (mmt_aggregate_f32 on hc), A[R,E+2U] = [e | h | mh], z[R,3U] = A W (NO bias).
 * mmt_train_frame_inputs_f32:   cur[R,2] = pos[:, t]; x[R,4] = [cur - pos[:, t-1] (0 at t = 0) | vis[:, min(t, T-1)]];
 *                               target[R,2] = pos[:, t+1] - cur (target may be NULL).  pos[R,F,2], vis[R,T,2].
 * mmt_train_gate_input_f32:     A[r] = [relu(x_r W_e + b_e) | hc[r,:U] | mhc[r,:U]].
 * mmt_gsk_gates_packed_f32:     mmt_gsk_gates_f32 on z + b with c = hc[:,U:], mc = mhc[:,U:]; writes hc_out[R,2U] = [h' | c'],
 *                               h_out[R,U] = h' (contiguous copy for the head) and mf_out[R,U].
 * mmt_gsk_cell_backward_packed_f32: mmt_gsk_cell_backward_f32 on z + b and the packed c / mc; d_head[R,2U] (NULL or the
 *                               head's gradient w.r.t. [m_t | m_f]) is added to d_mt / taken as d_mf; d mc goes to dmhc[:,U:].
 * mmt_train_backward_split_f32: from dA[R,E+2U] = dz W^T: dpre[R,E] = dA[:,:E] * (A[:,:E] > 0), gbe[E] += its column sums
 *                               (gbe may be NULL), dmhc[:,:U] = dA[:,E+U:].
 * mmt_train_backward_merge_f32: from back[R,2U] = att^T dmhc: Gh[R,U] = dA[:,E:E+U] + back[:,:U], Gc[R,U] = dc + back[:,U:]. */
int mmt_train_frame_inputs_f32(const float* pos, const float* vis, int R, int F, int T, int t, float* cur, float* x,
                               float* target, void* stream);
int mmt_train_gate_input_f32(const float* x, const float* hc, const float* mhc, const float* W_e, const float* b_e, int R,
                             int E, int U, float* A, void* stream);
int mmt_gsk_gates_packed_f32(const float* z, const float* b, const float* hc, const float* mhc, const uint8_t* valid,
                             const float* w_If, const float* w_It, const float* w_Of, const float* w_Ot, int R, int U,
                             float* hc_out, float* h_out, float* mf_out, void* stream);
int mmt_gsk_cell_backward_packed_f32(const float* z, const float* b, const float* hc, const float* mhc, const uint8_t* valid,
                                     const float* w_If, const float* w_It, const float* w_Of, const float* w_Ot,
                                     const float* d_mt, const float* d_head, const float* d_ct, int R, int U, float* dz,
                                     float* dc, float* dmhc, float* dpeep, float* db, void* stream);
int mmt_train_backward_split_f32(const float* dA, const float* A, int R, int E, int U, float* dpre, float* dmhc, float* gbe,
                                 void* stream);
int mmt_train_backward_merge_f32(const float* dA, const float* back, const float* dc, int R, int E, int U, float* Gh,
                                 float* Gc, void* stream);

/* number of kernel launches issued by this process through the library (bench's gpu_launches) */
uint64_t mmt_launch_count(void);

/* ---- tensor-core GEMM of the training step (TMA tensor maps -> tcgen05.mma.kind::tf32 -> TMEM) -------------------
 * The reference has no backward pass (the flags of argParser.py:40-47 are never read); the training step
 * defined for it contracts  dW += [e|h|mh]^T dz  (K = all agent rows),  dA = dz W^T  and their
 * smaller relatives.  C[M,N] = alpha * op(A) op(B) (+ C if accumulate), fp32 in memory, operands rounded to tf32 by
 * the tensor core, fp32 accumulation.  op(A) is [M,K]: transA = 0 -> A is [M,K] row-major (lda >= K), transA = 1 -> A
 * is [K,M] row-major (lda >= M); op(B) is [K,N]: transB = 0 -> B is [K,N] row-major (ldb >= N), transB = 1 -> B is
 * [N,K] row-major (ldb >= K).  lda, ldb multiples of 4; A, B 16-byte aligned.  Small outputs are split over K and
 * accumulated with red.global.add (summation order not fixed). */
int mmt_gemm_tf32(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc, int M,
                  int N, int K, float alpha, int accumulate, void* stream);

/* ---- the one collective of the path: gradient all-reduce of data-parallel training -----------
 * The reference has no parallelism (train.py:28-41 is a serial loop over datasets and batches); the north_star asks
 * for "an NCCL-over-NVLink gradient allreduce".  In-place SUM (MAX for timings) of n floats over the ranks of `comm`
 * (an ncclComm_t owned by the caller, e.g. torch.distributed's ProcessGroupNCCL._comm_ptr()) on `stream`.  libmmt does
 * not link NCCL: the calls are resolved from the NCCL instance already loaded in the process, the only one in which
 * the caller's communicator is valid.  Returns MMT_ENCCL (text via mmt_last_error) on any NCCL failure. */
int mmt_allreduce_f32(void* comm, float* buf, size_t n, void* stream);
int mmt_allreduce_max_f32(void* comm, float* buf, size_t n, void* stream);

/* ---- generic workspace / packed-image size query ---------------------------------------------------------------- */
typedef struct mmt_shape {
  int S, N, T, P, K;     /* scenes, agents per scene, observed / predicted frames, samples */
  int U, E, He, D;       /* hidden units, embedding, edge-MLP width, static-context grid   */
  int relational, prec;  /* MMT_PREC_*                                                     */
  int img_h;             /* MMT_OP_STATIC_CONTEXT: image height                            */
} mmt_shape;
#define MMT_OP_FORECAST 0            /* mmt_forecast_f32 workspace                          */
#define MMT_OP_EDGE_MLP 1            /* mmt_edge_mlp_{f32,bf16} work                        */
#define MMT_OP_STATIC_CONTEXT 2      /* mmt_static_context_f32 workspace                    */
#define MMT_OP_GATE_WEIGHTS_BF16 3   /* mmt_pack_gate_weights_bf16 image                    */
#define MMT_OP_GATE_WEIGHTS_BF16X3 4 /* mmt_pack_gate_weights_bf16x3 image                  */
#define MMT_OP_EDGE_WEIGHTS_BF16 5   /* mmt_pack_edge_weights_bf16 image                    */
int mmt_workspace_bytes(int op, const mmt_shape* shape, size_t* out);

#ifdef __cplusplus
}
#endif
#endif /* MMT_H_ */
