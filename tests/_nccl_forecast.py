"""Worker of tests/test_gpu_multi.py: one process per GPU under an initialised NCCL group.  Every rank forecasts its
own scene shard (eagerly and through a captured CUDA graph) and the whole batch; the shard must equal the matching
slice of the whole batch bit for bit (noise keyed by the global agent index), with collectives between the launches."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from multimodaltraj_2_b200 import ops, synth  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S, N, T, P, K = 296 * world, 64, 8, 12, 20                 # 148 tiles per rank: every SM of the rollout kernel busy
    pos, vis, valid = (torch.from_numpy(a).to(dev) for a in synth.make_crowd(S, N, seed=3, ragged=True))
    p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
    whole = ops.Forecaster(p, S, N, T, P, K, prec=ops.PREC_BF16, seed=7, device=dev)
    ref = {k: v.clone() for k, v in whole(pos, vis, valid).items() if v is not None}
    lo, hi = rank * S // world, (rank + 1) * S // world
    sh = [t[lo:hi].contiguous() for t in (pos, vis, valid)]
    for use_graph in (False, True):
        fc = ops.Forecaster(p, hi - lo, N, T, P, K, prec=ops.PREC_BF16, seed=7, agent_offset=lo * N, device=dev,
                            use_graph=use_graph)
        for it in range(6):
            out = fc(*sh)
            t = torch.stack([out["best_ade"].sum(), out["best_fde"].sum()])
            dist.all_reduce(t)                                  # a collective between consecutive launches
            dist.barrier()
            for k in ("params", "best_k", "best_ade", "best_fde", "best_traj"):
                assert torch.equal(out[k], ref[k][lo:hi]), (rank, use_graph, it, k)
        tot = torch.stack([ref["best_ade"].sum(), ref["best_fde"].sum()]) if world == 1 else None
        if tot is not None:
            assert torch.allclose(t, tot, rtol=1e-5)
    # the C-ABI collective (mmt_allreduce_f32 over the process group's own communicator) == torch's all_reduce
    g = torch.arange(131072, device=dev, dtype=torch.float32) * (rank + 1)
    want = g.clone()
    dist.all_reduce(want)
    got = ops.allreduce_(g.clone())
    assert torch.equal(got, want), (rank, "sum")
    m = torch.tensor([float(rank), -float(rank)], device=dev)
    assert torch.equal(ops.allreduce_(m, op="max"), torch.tensor([float(world - 1), 0.0], device=dev))
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_FORECAST_OK", world)


if __name__ == "__main__":
    main()
