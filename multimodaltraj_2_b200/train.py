"""Mirror of ``train.py`` (reference :17-698): ``main()`` / ``train(args)`` keep their signatures.

The reference's "training" has no loss, optimiser or gradient step: per batch it builds
the model, runs the per-frame step and logs raw errors, then validates.  ``train(args)`` here does
the same work on the batched B200 path: leave-one-out over the datasets, device-side scene
batching, scene-sharded rollout + best-of-K scoring on every rank, ADE/FDE partial sums combined
with one 3-float all-reduce, flat ``{name: tensor}`` checkpoints every ``save_every`` batches.
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
    prec = ops.PREC_BF16 if getattr(args, "precision", "bf16") == "bf16" else ops.PREC_F32
    p = params if params is not None else ops.CellParams.from_numpy(
        synth.init_params(seed=0, E=args.embedding_size, U=args.rnn_size), device)
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
                fc = ops.Forecaster(p, hi - lo, N, T, P, K, relational=True, prec=prec, seed=d,
                                    agent_offset=lo * N, device=device)
                o = fc(pos, vis, valid)
                sums = torch.stack([o["best_ade"].sum(), o["best_fde"].sum(), valid.sum().float()])
            if dist is not None:
                dist.all_reduce(sums)
            ade, fde, n = (float(x) for x in sums.cpu())
            results[d] = dict(ade=ade / max(n, 1), fde=fde / max(n, 1), n_agents=int(n), seconds=time.time() - t0)
            if rank == 0:
                print('dataset {0}: ADE = {1:.4f}  FDE = {2:.4f}  agents = {3}  ({4:.2f} s)'.format(
                    d, results[d]["ade"], results[d]["fde"], int(n), results[d]["seconds"]))
    return results


def save_checkpoint(path, params: ops.CellParams):
    """Flat {name: tensor} checkpoint (replaces tf.train.Saver, train.py:330-343)."""
    torch.save({k: getattr(params, k).cpu() for k in params.__dataclass_fields__
                if isinstance(getattr(params, k), torch.Tensor) and k != "W_packed"}, path)


def load_checkpoint(path, device="cuda"):
    d = torch.load(path, map_location=device)
    return ops.CellParams(**d)


if __name__ == '__main__':
    main()
