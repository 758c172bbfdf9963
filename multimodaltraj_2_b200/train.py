"""Mirror of ``train.py`` (reference :17-698): ``main()`` / ``train(args)`` keep their signatures.

The reference's "training" has no loss, optimiser or gradient step: per batch it builds
the model, runs the per-frame step and logs raw errors, then validates.  ``train(args)`` here does
the same work on the batched B200 path: leave-one-out over the datasets, device-side scene
batching, scene-sharded rollout + best-of-K scoring on every rank, ADE/FDE partial sums combined
with one 3-float all-reduce, TensorFlow-bundle checkpoints (``--save_dir``; ``tf_bundle.py``) the reference's
``Saver`` can restore.
"""
from __future__ import annotations

import os
import time

import torch

from . import argParser as argsParser
from . import load_traj as load
from . import ops, synth


def main():
    args = argsParser.ArgsParser().parser.parse_args()
    return train(args)


def shard_range(n, rank, world):
    """Contiguous scene range of ``rank``: [rank*n/world, (rank+1)*n/world)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def train(args, params=None, datasets=(2, 3, 4), rank=None, world=None, device=None):
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", getattr(args, "world_size", 1))) if world is None else world
    device = device or torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    dist = torch.distributed if (world > 1 and torch.distributed.is_initialized()) else None
    prec = ops.prec_from_name(getattr(args, "precision", "fp16"))
    p = params if params is not None else ops.CellParams.from_numpy(
        synth.init_params(seed=0, E=args.embedding_size, U=args.rnn_size), device)
    save_dir = getattr(args, "save_dir", None)
    if params is None and save_dir and os.path.exists(os.path.join(save_dir, "checkpoint")):
        p = load_checkpoint(save_dir, device)                        # train.py:383-402 (get_checkpoint_state + restore)
    N, T, P, K = args.max_agents, args.obs_len, args.pred_len, args.K
    results = {}
    for l in {args.leaveDataset}:                                   # train.py:28
        for d in sorted(set(datasets) - {l}):                       # train.py:38-41
            t0 = time.time()
            dl = load.DataLoader(args, datasets=[0, 1, 2, 3, 4, 5, 6], sel=0, start=d)
            table = dl.device_table(device)
            pos, vis, valid, _ = dl.scene_batch(table, N, T + P)
            lo, hi = shard_range(pos.shape[0], rank, world)          # scenes shard with no data-path collective
            pos, vis, valid = pos[lo:hi].contiguous(), vis[lo:hi, :, :T].contiguous(), valid[lo:hi].contiguous()
            sums = torch.zeros(3, device=device)
            if hi > lo:
                # the reference's driver builds g2k_lstm_mcr (train.py:6,161); --variant mc selects g2k_lstm_mc, whose bf16
                # mode is the fused persistent rollout kernel
                fc = ops.Forecaster(p, hi - lo, N, T, P, K, relational=getattr(args, "variant", "mcr") != "mc", prec=prec,
                                    seed=d, agent_offset=lo * N, device=device)
                o = fc(pos, vis, valid)
                sums = torch.stack([o["best_ade"].sum(), o["best_fde"].sum(), valid.sum().float()])
            if dist is not None:
                dist.all_reduce(sums)
            ade, fde, n = (float(x) for x in sums.cpu())
            results[d] = dict(ade=ade / max(n, 1), fde=fde / max(n, 1), n_agents=int(n), seconds=time.time() - t0)
            if rank == 0:
                print('dataset {0}: ADE = {1:.4f}  FDE = {2:.4f}  agents = {3}  ({4:.2f} s)'.format(
                    d, results[d]["ade"], results[d]["fde"], int(n), results[d]["seconds"]))
                if save_dir:                                         # train.py:330-343 (same file naming)
                    os.makedirs(save_dir, exist_ok=True)
                    path = save_checkpoint(os.path.join(save_dir, 'g2k_MPC_model_kfold_train_{1}_{0}_{2}.ckpt'.format(0, d, 0)),
                                           p, global_step=len(results) - 1)
                    print("model saved to {}".format(path))
    return results


# --------------------------------------------------------------------------------------------
def track_a_batch(args, batch, target_traj, graph, fp, sess_g=None, C=None, hidden_state=None, frame=1, device="cuda",
                  vislet=None, vemb_prev=None):
    """One batch of the reference's per-frame loop as written (train.py:423-674, the same lines as :61-276 of the
    training branch): ConstructGraph -> batch_v (node-axis slice, n <= obs_len: defect F-4, kept) -> model ->
    per-frame state step with the carried hidden state -> the per-agent scores of :639-662.
    Returns None for a batch the reference skips (:428-429, :434-435, :442-444), else
    dict(pred[n,P,2], euc[n], err[n,2], num_nodes, hidden_state, frames)."""
    import numpy as np
    from .models import g2k_lstm_mcr as mcr
    if len(batch) == 0:
        return None
    graph_t = graph.ConstructGraph(current_batch=batch, framenum=fp, future_traj=target_traj)
    batch_v = list(graph_t.get_node_attr(param='node_pos_list').values())
    if len(batch_v) == 0 or np.array(batch_v).ndim <= 1:
        return None
    batch_v = np.array(batch_v)[frame:frame + args.obs_len]             # :438 slices the NODE axis
    if batch_v.shape[0] == 0:
        return None
    X = np.linalg.norm(batch_v, axis=2).T                                # [T, n]  (:439-444)
    T, n = X.shape
    dim = int(args.neighborhood_size / args.grid_size)
    H = args.rnn_size
    dev = torch.device(device)
    f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(dev)  # noqa: E731
    m = mcr.g2k_lstm_mcr(in_features=torch.zeros((dim, dim)), num_nodes=n, obs_len=T, hidden_size=H,
                         lambda_reg=args.lambda_param, sess_g=sess_g, pred_len=args.pred_len, device=dev)
    rng = np.random.default_rng(0)                                       # init_w: N(0,1), seed 0 (:446, :514-521)
    W_i, W_ii = rng.standard_normal((n, dim)), rng.standard_normal((dim, T))
    # vislet = first two rows of the batch's vislet table cut to n columns (:182-183 / :529-530)
    V = np.zeros((2, n))
    if vislet is not None and vislet.shape[0] == 2 and vislet.shape[1] >= n:
        V = np.asarray(vislet[:, :n], np.float64)
    if C is None:                                                        # ctxt.png is absent upstream: no static context
        C = np.zeros((dim, dim))
    Hs = torch.zeros((1, dim, H), device=dev) if hidden_state is None else hidden_state
    out = None
    for _ in batch:                                                      # :556: every frame of the batch, state carried
        out = m.forward_batched(f32(X[None]), f32(V[None]), f32(C[None]), Hs, f32(W_i), f32(W_ii), vemb_prev)
        Hs = out["Hs"]
    pred = out["pred"][0].contiguous()                                   # [n, P, 2]  (:636)
    ids = list(target_traj)[:n]                                          # zip(range(num_nodes), iter(target_traj)) :641
    P = args.pred_len
    tgt = np.zeros((len(ids), P, 2), np.float32)
    lens = np.zeros(len(ids), np.int32)
    for i, k in enumerate(ids):
        t = np.asarray(target_traj[k], np.float32)[:P]
        lens[i] = len(t)
        tgt[i, :len(t)] = t
    euc, err = ops.train_val_scores(pred[:len(ids)].contiguous(), f32(tgt), torch.as_tensor(lens).to(dev), len(target_traj))
    return dict(pred=pred, euc=euc, err=err, num_nodes=n, hidden_state=Hs, frames=len(batch), vemb=out["vemb"])


def validate(args, params=None, l=None, rank=None, world=None, device=None, sess_g=None, max_batches=None):
    """The validation branch (train.py:371-695) on the leave-out split ``l`` (default ``args.leaveDataset``).

    Two scores come back, labelled:
      * ``cv_ade`` / ``cv_fde`` -- the reference's own reductions (:639-674, :688-689: spectral norm / 12 per agent,
        ||stack(err)||_F / len(batch) per batch, means over batches) of the AS-WRITTEN Track-A model on the batches
        ``DataLoader.next_step`` yields, scene at a time as the reference runs them (rank 0; these are a few hundred tiny
        matrix products).  The reference resets ``frame_pointer`` to 0 here (:377) and gets an empty first batch
        (SURVEY F12); the pointer starts at the split's first frame instead.
      * ``best_of_k`` -- the north_star metric: best-of-K ADE / FDE of the batched forecaster over every obs+pred window
        of the validation columns (the 30 % of load_traj.py:125-134), scenes sharded over the ranks
        (``realdata.evaluate_split``).
    """
    import numpy as np
    from . import networkx_graph as nx_g
    from . import realdata
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", getattr(args, "world_size", 1))) if world is None else world
    device = device or torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    l = args.leaveDataset if l is None else l
    t0 = time.time()
    graph = nx_g.online_graph(args)
    dataloader = load.DataLoader(args=args, datasets=[0, 1, 2, 3, 4, 5], start=l, sel=0)
    dataloader.reset_data_pointer(valid=True, frame_pointer=dataloader.seed)          # :377 (see docstring)
    dataloader.valid_frame_pointer = int((dataloader.len - int(dataloader.max * .7)) / dataloader.val_max)   # :409-410
    dataloader.valid_num_batches = int(dataloader.val_max / dataloader.batch_size)     # :412
    cv_ade_err, cv_fde_err, n_scored = [], [], 0
    hidden = None
    nb = dataloader.valid_num_batches if max_batches is None else min(max_batches, dataloader.valid_num_batches)
    for vb in range(nb):
        batch, target_traj, fp = dataloader.next_step()
        vfp = int(dataloader.valid_frame_pointer)                                       # :527-528
        r = track_a_batch(args, batch, target_traj, graph, fp, sess_g=sess_g, hidden_state=hidden, device=device,
                          vislet=dataloader.vislet[:, vfp:vfp + args.obs_len] if dataloader.vislet.shape[0] == 2 else None)
        if r is None:
            break
        hidden = r["hidden_state"]
        euc, err = r["euc"].double().cpu().numpy(), r["err"].double().cpu().numpy()
        if len(euc):
            cv_ade_err.append(float(np.mean(euc)))                                         # :668-669
            cv_fde_err.append(float(np.linalg.norm(err) / (r["num_nodes"] if l == 5 else r["frames"])))   # :670-674
            n_scored += len(euc)
    res = dict(dataset=l, cv_ade=float(np.mean(cv_ade_err)) if cv_ade_err else float("nan"),
               cv_fde=float(np.mean(cv_fde_err)) if cv_fde_err else float("nan"), batches=len(cv_ade_err),
               agents_scored=n_scored)
    if rank == 0:
        print('Cross-Validation total mean error (ADE) for dataset {0} = '.format(l), res["cv_ade"])       # :688
        print('Cross-Validation total final error (FDE) for dataset {0} = '.format(l), res["cv_fde"])      # :689
    prec = ops.prec_from_name(getattr(args, "precision", "fp16"))
    p = params if params is not None else ops.CellParams.from_numpy(
        synth.init_params(seed=0, E=args.embedding_size, U=args.rnn_size), device)
    res["best_of_k"] = realdata.public(realdata.evaluate_split(args, l, p, part="val", prec=prec, rank=rank, world=world,
                                                              device=device))
    res["seconds"] = time.time() - t0
    return res


# --------------------------------------------------------------------------------------------
# training step (SURVEY App. C.5 "Training loss (fills F2)", section 8e): the reference has no loss or optimiser
# (train.py:23-366 only logs raw errors); the step defined for it is
#   teacher-forced rollout -> mean bivariate-Gaussian NLL of the next displacement + (lambda/2) ||W||^2
#   -> back-propagation through time -> ONE all-reduce of the flat gradient bucket -> RMSProp (lr 0.005, decay 0.95,
#   global-norm clip 10: the unused flags of argParser.py:40-47,72).
TRAIN_KEYS = ("W_e", "b_e", "W", "b", "w_If", "w_It", "w_Of", "w_Ot", "W_h", "b_h")
EDGE_KEYS = ("W1", "b1", "W2", "b2", "w_out", "b_out")     # + the relational edge MLP of g2k_lstm_mcr


def _bucket_keys(d: dict):
    return [k for k in TRAIN_KEYS + EDGE_KEYS if k in d]


def flatten_bucket(grads: dict):
    """One contiguous fp32 bucket (< 1 MB) in TRAIN_KEYS (+ EDGE_KEYS) order."""
    return torch.cat([grads[k].reshape(-1) for k in _bucket_keys(grads)])


def unflatten_bucket(flat, like: dict):
    out, o = {}, 0
    for k in _bucket_keys(like):
        n = like[k].numel()
        out[k] = flat[o:o + n].view_as(like[k])
        o += n
    return out


def allreduce_mean_(flat, counts=None):
    """Sum-all-reduce of the gradient bucket over the data-parallel ranks (NCCL on GPUs, gloo in the CPU tests).
    Every rank's gradient is the SUM over its own scenes of d nll; dividing by the global number of valid
    agent-steps turns the reduced sum into the gradient of the global mean loss, whatever the sharding."""
    dist = torch.distributed
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if flat.is_cuda and ops.nccl_comm_ptr() is not None:
            # the C-ABI collective (mmt_allreduce_f32): NCCL on the CURRENT stream, right behind the kernels that produced
            # the bucket -- no hop to the process group's internal stream, no extra event pair per step
            ops.allreduce_(flat)
            if counts is not None:
                ops.allreduce_(counts)
        else:                                    # gloo (the CPU tests of the host logic)
            dist.all_reduce(flat)
            if counts is not None:
                dist.all_reduce(counts)
    return flat


def rmsprop_update_(params: dict, grads: dict, ms: dict, lr=0.005, decay=0.95, eps=1e-10, clip=10.0):
    """In-place RMSProp with global-norm clipping; returns the pre-clip gradient norm."""
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    s = torch.clamp(clip / torch.clamp(gn, min=1e-30), max=1.0)
    for k, g in grads.items():
        g = g * s
        ms[k].mul_(decay).addcmul_(g, g, value=1 - decay)
        params[k].addcdiv_(g, ms[k].sqrt() + eps, value=-lr)
    return gn


class Trainer:
    """Data-parallel training of the g2k_lstm_mc cell on the B200 path (fp32 kernels of libmmt + library GEMMs).

    forward  per step: mmt_pairwise_adj_f32 -> mmt_aggregate_f32 (attention kept) -> mmt_gsk_cell (fp32)
    loss     emitting steps: mmt_head_nll_f32 (raw head, NLL, d nll / d y)
    backward per step, reversed: head GEMMs -> mmt_gsk_cell_backward_f32 -> dW += A^T dz, dA = dz W^T (GEMMs)
             -> relu / embedding -> att^T (d mh, d mc) back to the previous h, c
    step()   flat bucket -> all-reduce (SUM) -> / global valid count -> + lambda W -> clip -> RMSProp.
    """

    def __init__(self, params: ops.CellParams, T=8, P=12, r2=4.0, inv_2sigma2=0.5, lam=0.0005, lr=0.005, decay=0.95,
                 clip=10.0, gemm="fp32", relational=False, graph=False):
        # gemm: arithmetic of the library contractions of the backward pass (A^T dz, dz W^T, att^T d mh, head):
        #   "fp32" CUDA-core SGEMM (parity mode: gradients within 2e-3 of the fp64 autograd oracle),
        #   "tf32" tensor cores, operands rounded to 10 mantissa bits, fp32 accumulation (stated separately: 2e-2); the
        #          forward gate GEMM then is a library tensor-core GEMM + mmt_gsk_gates_f32 as well (z kept, not recomputed)
        #          instead of the fused fp32 CUDA-core cell kernel.
        #   "tc"   this library's own contractions: mmt_gemm_tf32 (TMA tensor maps -> tcgen05.mma.kind::tf32 -> TMEM; split-K
        #          with red.add for the weight gradients) for the gate GEMM forward, A^T dz and dz W^T -- 97 % of the step's
        #          FLOPs -- the head-weight and embedding-weight gradients ([hn|mf]^T dy, x^T dpre: K = all agent rows too),
        #          mmt_aggregate_transpose_f32 for att^T [d mh | d mc], the bias gradients inside the backward kernels, and the
        #          element-wise glue between them on packed rows (_packed_step); same tolerance as "tf32".  Left as library
        #          calls: dy W_h^T (inner dimension 5) and two [R,5] reductions.
        if gemm not in ("fp32", "tf32", "tc"):
            raise ValueError("gemm must be 'fp32', 'tf32' or 'tc'")
        self.gemm = gemm
        # relational: g2k_lstm_mcr -- the attention logits are kern + the edge-MLP score (mmt_edge_mlp_f32); its backward
        # (softmax -> per-edge two-layer ELU MLP -> node projections) runs on the device-side edge list of
        # mmt_attention_score_grad_f32 + mmt_edge_mlp_backward_f32 (gemm = "tc", U = He = 128: _bf16 on tcgen05), no host sync
        self.relational = bool(relational)
        if self.relational and params.W1 is None:
            raise ValueError("relational training needs the edge-MLP weights (W1 .. b_out)")
        # graph: step() replays the forward + BPTT of one shard (~1000 launches, no host synchronisation, fixed shapes) as ONE
        # CUDA graph, captured on the first step of each input shape; all-reduce, clipping and RMSProp stay outside it.
        self.graph = bool(graph)
        self._captured = {}                                            # input shapes -> (CUDAGraph, static inputs, outputs, launches)
        self.keys = TRAIN_KEYS + (EDGE_KEYS if self.relational else ())
        self.p, self.T, self.P, self.r2, self.inv = params, T, P, r2, inv_2sigma2
        self.lam, self.lr, self.decay, self.clip = lam, lr, decay, clip
        self.ms = {k: torch.zeros_like(getattr(params, k)) for k in self.keys}

    def _tensors(self):
        return {k: getattr(self.p, k) for k in self.keys}

    def loss_and_grad_sums(self, pos, vis, valid):
        """Teacher-forced forward + BPTT on this rank's scenes.  Returns (sum of nll over valid agent-steps [1],
        number of valid agent-steps, {name: SUM-gradient}) -- sums, so that ranks combine by plain addition."""
        tf32_was = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.gemm == "tf32"
        try:
            return self._loss_and_grad_sums(pos, vis, valid)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32_was

    def _packed_step(self, pos, vis, valid):
        """gemm = "tc": forward + BPTT of the shard on PACKED rows (hc = [h | c], mhc = [mh | mc], A = [e | h | mh], z without
        bias) -- every kernel between the frame's inputs and its gradients is this library's, the element-wise glue included
        (mmt_train_frame_inputs_f32, mmt_train_gate_input_f32, mmt_gsk_gates_packed_f32, mmt_gsk_cell_backward_packed_f32,
        mmt_train_backward_split_f32 / _merge_f32).  Left to the library: dy W_h^T (inner dimension 5), two [R,5] reductions,
        and, for g2k_lstm_mcr, the addition of the edge scores to the logits."""
        p, T, P = self.p, self.T, self.P
        S, N = valid.shape
        R, U, E = S * N, p.U, p.E
        dev = pos.device
        vflat = valid.reshape(-1).contiguous()
        hc = torch.zeros((R, 2 * U), device=dev)
        saved = []
        loss_sum = torch.zeros((1,), device=dev)
        edge_packed = None
        if self.relational and U == 128 and p.W2.shape[0] == 128:
            edge_packed = ops.pack_edge_weights(p.W1, p.W2)
        for t in range(T + P - 1):
            cur, x, target = ops.train_frame_inputs(pos, vis, t, t >= T - 1)
            kern, adj, _ = ops.pairwise_adj(cur, valid, self.r2, self.inv, want_deg=False)
            rec = dict(x=x, hc=hc)
            if self.relational:
                rec["h"] = hc[:, :U].contiguous()
                kern = kern + ops.edge_mlp(rec["h"].view(S, N, U), adj, p.W1, p.b1, p.W2, p.b2, p.w_out, p.b_out,
                                           ops.PREC_BF16 if edge_packed is not None else ops.PREC_F32, edge_packed)
                rec["adj"] = adj
            att, mhc = ops.aggregate(kern, adj, hc.view(S, N, 2 * U))
            mhc = mhc.view(R, 2 * U)
            A = ops.train_gate_input(x, hc, mhc, p)
            z = ops.gemm_tf32(A, p.W)                                   # pre-activations without the bias (added in the kernels)
            hc_next, hn, mf = ops.gsk_gates_packed(z, hc, mhc, vflat, p)
            rec.update(att=att, mhc=mhc, A=A, z=z, hn=hn, mf=mf)
            if t >= T - 1:
                rec["dy"] = ops.head_nll(hn, mf, vflat, p.W_h, p.b_h, target, 1.0, loss_sum)
            saved.append(rec)
            hc = hc_next
        # ---- back-propagation through time
        g = {k: torch.zeros_like(getattr(p, k)) for k in self.keys}
        dpeep = torch.zeros((4, U), device=dev)
        Gh, Gc = torch.zeros((R, U), device=dev), None
        gWh8 = torch.zeros((2 * U, 8), device=dev)                      # head-weight gradient, 8-column rows (ld % 4 == 0)
        node = None
        if self.relational:
            He = p.W2.shape[0]
            node = dict(W1cat=torch.cat([p.W1[:U], p.W1[U:]], 1).contiguous(), gW1=torch.zeros((U, 2 * He), device=dev),
                        packed=edge_packed)
        for t in reversed(range(T + P - 1)):
            r = saved[t]
            d_head = None
            if "dy" in r:
                dy = r["dy"]
                dy8 = torch.nn.functional.pad(dy, (0, 3))
                ops.gemm_tf32(r["hn"], dy8, transA=True, out=gWh8[:U], accumulate=True)
                ops.gemm_tf32(r["mf"], dy8, transA=True, out=gWh8[U:], accumulate=True)
                g["b_h"] += dy.sum(0)
                d_head = dy @ p.W_h.t()                                 # [R, 2U]: gradient w.r.t. [m_t | m_f]
            A, z = r.pop("A"), r.pop("z")
            dz, dc, dmhc = ops.gsk_cell_backward_packed(z, r["hc"], r["mhc"], vflat, p, Gh, d_head, Gc, dpeep, g["b"])
            ops.gemm_tf32(A, dz, transA=True, out=g["W"], accumulate=True)      # dW += A^T dz  (K = all agent rows)
            dA = ops.gemm_tf32(dz, p.W, transB=True)                           # dA  = dz W^T
            dpre = ops.train_backward_split(dA, A, dmhc, g["b_e"], p)          # relu mask, b_e gradient, d mh into dmhc
            ops.gemm_tf32(r["x"], dpre, transA=True, out=g["W_e"], accumulate=True)
            back = ops.aggregate_transpose(r["att"], dmhc.view(S, N, 2 * U)).view(R, 2 * U)
            Gh, Gc = ops.train_backward_merge(dA, back, dc, p)
            if self.relational:
                Gh = Gh + self._edge_backward(r, dmhc.view(S, N, 2 * U), g, S, N, node)
        g["W_h"] += gWh8[:, :5]
        if self.relational:
            He = p.W2.shape[0]
            g["W1"][:U] += node["gW1"][:, :He]
            g["W1"][U:] += node["gW1"][:, He:]
        g["w_If"], g["w_It"], g["w_Of"], g["w_Ot"] = dpeep[0], dpeep[1], dpeep[2], dpeep[3]
        return loss_sum, valid.sum().float() * P, g

    def _loss_and_grad_sums(self, pos, vis, valid):
        if self.gemm == "tc":
            return self._packed_step(pos, vis, valid)
        p, T, P = self.p, self.T, self.P
        S, N = valid.shape
        R, U, E = S * N, p.U, p.E
        dev = pos.device
        vflat = valid.reshape(-1).contiguous()
        h = torch.zeros((S, N, U), device=dev)
        c = torch.zeros((S, N, U), device=dev)
        saved = []
        loss_sum = torch.zeros((1,), device=dev)
        for t in range(T + P - 1):                                     # fp32 / tf32 (library GEMM) modes: tensors per quantity
            cur = pos[:, :, t].contiguous()
            disp = cur - pos[:, :, t - 1] if t > 0 else torch.zeros_like(cur)
            x = torch.cat([disp, vis[:, :, min(t, T - 1)]], -1).reshape(R, 4).contiguous()
            kern, adj, _ = ops.pairwise_adj(cur, valid, self.r2, self.inv, want_deg=False)
            if self.relational:
                kern = kern + ops.edge_mlp(h.contiguous(), adj, p.W1, p.b1, p.W2, p.b2, p.w_out, p.b_out)
            att, mhc = ops.aggregate(kern, adj, torch.cat([h, c], -1).contiguous())
            mh, mc = mhc[..., :U].reshape(R, U).contiguous(), mhc[..., U:].reshape(R, U).contiguous()
            rec = dict(x=x, h=h.reshape(R, U), c=c.reshape(R, U), att=att, mh=mh, mc=mc)
            if self.relational:
                rec["adj"] = adj
            if self.gemm == "tf32":
                # gate GEMM on the tensor cores (library GEMM) + mmt_gsk_gates_f32; z and e are kept for the backward
                e = torch.relu(torch.addmm(p.b_e, x, p.W_e))
                z = torch.addmm(p.b, torch.cat([e, rec["h"], mh], -1), p.W)
                hn, cn, mf = ops.gsk_gates(z, rec["c"], mc, vflat, p)
                rec["e"], rec["z"] = e, z
            else:
                hn, cn, mf = ops.gsk_cell(x, rec["h"], rec["c"], mh, mc, vflat, p, ops.PREC_F32)
            rec["hn"], rec["mf"] = hn, mf
            if t >= T - 1:
                rec["dy"] = ops.head_nll(hn, mf, vflat, p.W_h, p.b_h, (pos[:, :, t + 1] - cur).reshape(R, 2).contiguous(),
                                         1.0, loss_sum)
            saved.append(rec)
            h, c = hn.view(S, N, U), cn.view(S, N, U)
        # ---- back-propagation through time
        return self._backward(saved, loss_sum, valid, vflat, S, N)

    def _backward(self, saved, loss_sum, valid, vflat, S, N):
        p, T, P = self.p, self.T, self.P
        R, U, E = S * N, p.U, p.E
        dev = vflat.device
        g = {k: torch.zeros_like(getattr(p, k)) for k in self.keys}
        dpeep = torch.zeros((4, U), device=dev)
        Gh = torch.zeros((R, U), device=dev)
        Gc = None
        node = None
        if self.relational:                                            # node level of the edge MLP: [a | b] = h [W1a | W1b]
            He = p.W2.shape[0]
            node = dict(W1cat=torch.cat([p.W1[:U], p.W1[U:]], 1).contiguous(), gW1=torch.zeros((U, 2 * He), device=dev),
                        packed=None)
        for t in reversed(range(T + P - 1)):
            r = saved[t]
            d_mf = None
            if "dy" in r:
                dy = r["dy"]
                g["W_h"] += torch.cat([r["hn"], r["mf"]], -1).t() @ dy
                g["b_h"] += dy.sum(0)
                dhm = dy @ p.W_h.t()
                Gh = Gh + dhm[:, :U]
                d_mf = dhm[:, U:].contiguous()
            e = r["e"] if "e" in r else torch.relu(r["x"] @ p.W_e + p.b_e)
            A = torch.cat([e, r["h"], r["mh"]], -1)
            z = r.pop("z") if "z" in r else torch.addmm(p.b, A, p.W)
            dz, dc, dmc = ops.gsk_cell_backward(z, r["c"], r["mc"], vflat, p, Gh.contiguous(), d_mf, Gc, dpeep)
            g["W"] += A.t() @ dz
            dA = dz @ p.W.t()
            g["b"] += dz.sum(0)
            dpre = dA[:, :E] * (e > 0)
            g["W_e"] += r["x"].t() @ dpre
            g["b_e"] += dpre.sum(0)
            d_mh = dA[:, E + U:].reshape(S, N, U)
            attT = r["att"].transpose(1, 2)
            Gh = dA[:, E:E + U] + torch.bmm(attT, d_mh).reshape(R, U)
            Gc = (dc + torch.bmm(attT, dmc.view(S, N, U)).reshape(R, U)).contiguous()
            if self.relational:
                dmhc = torch.cat([d_mh, dmc.view(S, N, U)], -1).contiguous()
                Gh = Gh + self._edge_backward(r, dmhc, g, S, N, node)
        if self.relational:
            He = p.W2.shape[0]
            g["W1"][:U] += node["gW1"][:, :He]
            g["W1"][U:] += node["gW1"][:, He:]
        g["w_If"], g["w_It"], g["w_Of"], g["w_Ot"] = dpeep[0], dpeep[1], dpeep[2], dpeep[3]
        return loss_sum, valid.sum().float() * P, g

    def _edge_backward(self, r, dmhc, g, S, N, node):
        """Back through the attention softmax and the relational edge MLP of one step (oracle: train_b.edge_scores):
        accumulates the edge-weight gradients into ``g`` / ``node`` and returns d loss / d h[R,U] through the scores.
        dmhc[S,N,2U] = d loss / d [mh | mc].  All of it on the device-side edge list of the kernels: no host sync.
          mmt_attention_score_grad_f32   d logit_ij = att_ij (G_ij - sum_k att_ik G_ik), G_ij = d mh_i . h_j + d mc_i . c_j
          mmt_edge_mlp_backward_f32/bf16 per edge: sigmoid -> ELU layer 2 -> ELU layer 1; g W2, g b*, g w_out; [d a | d b] rows
          node level                     g W1 += h^T [d a | d b],  d h = [d a | d b] [W1a | W1b]^T   (two GEMMs)."""
        p = self.p
        U, He = p.U, p.W2.shape[0]
        h = r["h"]
        hc = (r["hc"] if "hc" in r else torch.cat([h, r["c"]], -1)).view(S, N, 2 * U)
        dlog = ops.attention_score_grad(r["att"], r["adj"], dmhc, hc)
        packed = node["packed"]
        dab = ops.edge_mlp_backward(h.view(S, N, U), r["adj"], dlog, p, g, ops.PREC_BF16 if packed is not None else ops.PREC_F32,
                                    packed)
        if self.gemm == "tc":
            ops.gemm_tf32(h, dab, transA=True, out=node["gW1"], accumulate=True)
            return ops.gemm_tf32(dab, node["W1cat"], transB=True)
        node["gW1"] += h.t() @ dab
        return dab @ node["W1cat"].t()

    def loss_and_grads(self, pos, vis, valid):
        """Single-process view: (mean loss incl. weight decay, {name: gradient of it})."""
        loss_sum, n, g = self.loss_and_grad_sums(pos, vis, valid)
        g = {k: v / n for k, v in g.items()}
        g["W"] = g["W"] + self.lam * self.p.W
        return loss_sum[0] / n + 0.5 * self.lam * (self.p.W * self.p.W).sum(), g

    def _replay(self, pos, vis, valid):
        """loss_and_grad_sums through a CUDA graph (captured once per input shape; the weights are updated in place, so a
        replay reads the current ones).  The outputs are the graph's own buffers: consumed before the next replay."""
        key = (tuple(pos.shape), tuple(vis.shape), tuple(valid.shape), pos.device)
        if key not in self._captured:
            if len(self._captured) >= 8:                               # a handful of batch shapes (one per training table)
                self._captured.pop(next(iter(self._captured)))
            static = tuple(t.clone() for t in (pos, vis, valid))
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):                              # warm-up off the capture: opt-ins, library handles
                self.loss_and_grad_sums(*static)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            n0 = ops.launch_count()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                out = self.loss_and_grad_sums(*static)
            self.graph_launches = ops.launch_count() - n0              # this library's kernels inside one replay (last capture)
            self._captured[key] = (cg, static, out, self.graph_launches)
        cg, static, out, n_launches = self._captured[key]
        for dst, src in zip(static, (pos, vis, valid)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
        cg.replay()
        ops._graph_launches += n_launches                             # replayed launches, for ops.launch_count()
        return out

    def step(self, pos, vis, valid):
        """One data-parallel training step on this rank's scene shard; returns the global mean loss."""
        loss_sum, n, g = self._replay(pos, vis, valid) if self.graph else self.loss_and_grad_sums(pos, vis, valid)
        flat = flatten_bucket(g)
        counts = torch.stack([loss_sum[0], n])
        allreduce_mean_(flat, counts)                                  # ONE all-reduce of the < 1 MB bucket (+ 2 floats)
        g = unflatten_bucket(flat / counts[1], g)
        g["W"] = g["W"] + self.lam * self.p.W
        rmsprop_update_(self._tensors(), g, self.ms, self.lr, self.decay, clip=self.clip)
        self.p.repack()                                                # the bf16 operand images are stale now: refreshed in place
        return counts[0] / counts[1] + 0.5 * self.lam * (self.p.W * self.p.W).sum()


def _param_names(params):
    return [k for k in params.__dataclass_fields__ if isinstance(getattr(params, k), torch.Tensor) and k not in ("W_packed", "W_packed_x3", "W_packed_f16")]


def save_checkpoint(path, params: ops.CellParams, trainer: "Trainer" = None, global_step=None):
    """``saver.save(sess, checkpoint_path, global_step=...)`` of train.py:330-343.

    ``path`` ending in ``.pt``: flat torch ``{name: tensor}`` file.  Otherwise ``path`` is a checkpoint PREFIX and the
    tensors are written as a TensorFlow bundle (``prefix[-step].index`` + ``.data-00000-of-00001`` + the ``checkpoint``
    state file; multimodaltraj_2_b200/tf_bundle.py) that ``tf.train.Saver`` of the reference restores.  With ``trainer``
    the RMSProp accumulators go in as ``<name>/RMSProp`` (TF's slot naming), which is what resuming needs."""
    tensors = {k: getattr(params, k).detach().cpu() for k in _param_names(params)}
    if trainer is not None:
        tensors.update({f"{k}/RMSProp": v.detach().cpu() for k, v in trainer.ms.items()})
    if str(path).endswith(".pt"):
        torch.save(tensors, path)
        return str(path)
    from pathlib import Path
    from . import tf_bundle
    prefix = f"{path}-{int(global_step)}" if global_step is not None else str(path)
    tf_bundle.write_checkpoint(prefix, {k: v.numpy() for k, v in tensors.items()})
    tf_bundle.save_state(Path(prefix).parent, Path(prefix).name)
    return prefix


def load_checkpoint(path, device="cuda", trainer: "Trainer" = None):
    """Restore what ``save_checkpoint`` wrote (``saver.restore``, train.py:383-402); ``path`` may also be a directory
    holding a ``checkpoint`` state file (``tf.train.get_checkpoint_state``).  Fills ``trainer.ms`` when given."""
    import os
    if str(path).endswith(".pt"):
        d = torch.load(path, map_location=device)
    else:
        from . import tf_bundle
        prefix = tf_bundle.latest_checkpoint(path) if os.path.isdir(path) else str(path)
        if prefix is None:
            raise FileNotFoundError(f"no checkpoint state in {path}")
        d = {k: torch.from_numpy(v.copy()).to(device) for k, v in tf_bundle.read_checkpoint(prefix).items()}
    slots = {k[:-len("/RMSProp")]: v for k, v in d.items() if k.endswith("/RMSProp")}
    params = ops.CellParams(**{k: v for k, v in d.items() if not k.endswith("/RMSProp")})
    if trainer is not None:
        trainer.p = params
        for k, v in slots.items():
            trainer.ms[k] = v.to(device)
    return params


if __name__ == '__main__':
    main()
