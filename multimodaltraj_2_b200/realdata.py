"""Real-data evaluation on the ETH / UCY tables (BASELINE configs C1 and C5).

What the reference does per batch with Python dicts (``load_traj.DataLoader.next_step`` :153-224 ->
``online_graph.ConstructGraph`` networkx_graph.py:30-73 -> per-frame ``sess.run`` train.py:556-636 -> error lists
:639-674) is here one device-side pass per split: the (frame, ped)-sorted table goes to HBM once, ``mmt_scene_batch_f32``
cuts every obs+pred window into padded scenes, the scenes shard over the ranks with no data-path collective, the
forecaster (pairwise kernel -> aggregation -> [edge MLP] -> gate update -> K-sample decode -> best-of-K) runs on the
shard, and three floats per split (sum ADE, sum FDE, agents) are all-reduced.

Units: the tables hold normalised pixels (ETH) or z-scores (UCY); ADE/FDE are reported in those units, and for the two
ETH scenes also in metres through the scene homography (``mmt_ade_fde_world_f32``, getPixelCoordinates.m:8-30 run
backwards).  The homographies are the BIWI dataset's ``H.txt`` (image -> world), which ``getPixelCoordinates.m:8`` reads
but the reference tree does not ship; UCY needs its z-score constants, which are not in the tree either -> no metres.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import load_traj as load
from . import ops

DATASET_NAMES = {0: "eth/hotel", 1: "eth/univ", 2: "ucy/zara01", 3: "ucy/zara02", 4: "ucy/univ"}
# BIWI walking-pedestrians dataset, seq_hotel/H.txt and seq_eth/H.txt: [x y w]^T = H [u v 1]^T
HOMOGRAPHY = {
    0: [[1.1048200e-02, 6.6958900e-04, -3.3295300e+00],
        [-1.5966000e-03, 1.1632400e-02, -5.3951400e+00],
        [1.1190700e-04, 1.3617400e-05, 5.4276600e-01]],
    1: [[2.8128700e-02, 2.0091900e-03, -4.6693600e+00],
        [8.0625700e-04, 2.5195500e-02, -5.0608800e+00],
        [3.4555400e-04, 9.2512200e-05, 4.6255300e-01]],
}


def shard_range(n, rank, world):
    """Contiguous scene range of ``rank``: [rank*n/world, (rank+1)*n/world)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def pad_agents(max_present):
    """Agent slots per scene: the fused rollout takes N in {8, 16, 32, 64, 128} (128 % N == 0)."""
    for n in (8, 16, 32, 64, 128):
        if max_present <= n:
            return n
    return (max_present + 127) // 128 * 128


def scene_windows(args, d, part="val", device="cuda", N=None, hop=1, data_root=None):
    """Every obs+pred window of split ``d`` (``load_traj.py:25-33`` index) as padded scenes on the device.
    part: 'train' (first 70 % of the columns, load_traj.py:125-134), 'val' (the next 30 %) or 'all'.
    Returns dict(pos[S,N,F,2], vis[S,N,T,2] (zeros for the 4-row ETH tables), valid[S,N], slot_ped[S,N], N, loader)."""
    dl = load.DataLoader(args, datasets=[0, 1, 2, 3, 4, 5, 6], sel=0, start=d, parent_dir=data_root)
    if part == "all":
        dl.tr_data = dl.raw_data
    table = dl.device_table(device, val=(part == "val"))
    if N is None:
        counts = (table["row_start"][1:] - table["row_start"][:-1])
        N = pad_agents(int(counts.max()) if counts.numel() else 1)
    T, F = args.obs_len, args.obs_len + args.pred_len
    pos, vis, valid, slot = dl.scene_batch(table, N, F, hop=hop)
    vis = torch.zeros((pos.shape[0], N, T, 2), device=pos.device) if vis is None else vis[:, :, :T].contiguous()
    keep = valid.sum(1) > 0                                        # windows with no pedestrian present throughout
    return dict(pos=pos[keep].contiguous(), vis=vis[keep].contiguous(), valid=valid[keep].contiguous(),
                slot_ped=slot[keep].contiguous(), N=N, loader=dl, windows=int(pos.shape[0]))


def evaluate_split(args, d, params, part="val", prec=ops.PREC_F16, relational=False, rank=0, world=1, device="cuda",
                   seed=None, eps=None, data_root=None, scenes=None):
    """Best-of-K ADE / FDE of one split on this rank's scene shard, combined over the ranks.
    eps: optional fed noise [S,N,K,P,2] for the WHOLE split (parity runs); otherwise in-kernel Philox keyed by the global
    agent index, so the result does not depend on the sharding.  Returns dict(ade, fde, n_agents, scenes, N, ...)."""
    t0 = time.time()
    sc = scenes if scenes is not None else scene_windows(args, d, part, device, data_root=data_root)
    S, N, T, P, K = sc["pos"].shape[0], sc["N"], args.obs_len, args.pred_len, args.K
    lo, hi = shard_range(S, rank, world)
    sums = torch.zeros(5, device=device, dtype=torch.float64)
    out = None
    if hi > lo:
        fc = ops.Forecaster(params, hi - lo, N, T, P, K, relational=relational, prec=prec, seed=d if seed is None else seed,
                            agent_offset=lo * N, device=device)
        pos, vis, valid = (sc[k][lo:hi].contiguous() for k in ("pos", "vis", "valid"))
        out = fc(pos, vis, valid, eps=None if eps is None else eps[lo:hi].contiguous())
        sums[:3] = torch.stack([out["best_ade"].sum().double(), out["best_fde"].sum().double(), valid.sum().double()])
        if d in HOMOGRAPHY:                                          # metres (ETH scenes)
            H = torch.tensor(HOMOGRAPHY[d], dtype=torch.float32, device=device)
            gt = pos[:, :, T:].reshape(-1, P, 2).contiguous()
            _, _, s3 = ops.ade_fde_world(out["best_traj"].reshape(-1, P, 2), gt, H, valid.reshape(-1).contiguous())
            sums[3:5] = s3[:2].double()
    dist = torch.distributed
    if world > 1 and dist.is_available() and dist.is_initialized():
        dist.all_reduce(sums)                                        # the only collective of the evaluation
    a, f, n, am, fm = (float(x) for x in sums.cpu())
    res = dict(dataset=DATASET_NAMES.get(d, str(d)), ade=a / max(n, 1), fde=f / max(n, 1), n_agents=int(n), scenes=S,
               agents_per_scene=N, part=part, seconds=time.time() - t0)
    if d in HOMOGRAPHY:
        res.update(ade_m=am / max(n, 1), fde_m=fm / max(n, 1))
    res["_out"], res["_scenes"] = out, sc
    return res


def public(res):
    """The JSON-able part of an evaluate_split result."""
    return {k: v for k, v in res.items() if not k.startswith("_")}
