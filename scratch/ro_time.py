"""Kernel time of rollout_tc_kernel at C3 for the library named by MMT_LIB (default: in-tree): CUDA events, 3 + 20 launches,
inputs cycled over 4 device copies.  `python scratch/ro_time.py all` runs every variant under build/ in a subprocess each."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def one():
    import torch
    from multimodaltraj_2_b200 import ops, synth
    dev = torch.device("cuda")
    S, N = 4096, 64
    p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
    base = [torch.from_numpy(a).to(dev) for a in synth.make_crowd(S, N, seed=synth.SEED)]
    sets = [base] + [[t.clone() for t in base] for _ in range(3)]
    out = torch.empty((S, N, 12, 5), device=dev)
    for i in range(3):
        ops.rollout_bf16(*sets[i % 4], p, out=out)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            ops.rollout_bf16(*sets[i % 4], p, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
    print(f"{os.environ.get('MMT_LIB', 'in-tree'):40s} rollout_tc_kernel ms/launch: " + " ".join(f"{t:.4f}" for t in ts)
          + f"  checksum {float(out.double().abs().sum()):.6f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        libs = [None] + sorted(str(p) for p in (ROOT / "build").glob("libmmt_*.so"))
        for _ in range(2):
            for lib in libs:
                env = dict(os.environ)
                env.pop("MMT_LIB", None)
                if lib:
                    env["MMT_LIB"] = lib
                subprocess.run([sys.executable, __file__], env=env)
    else:
        one()
