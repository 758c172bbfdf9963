"""Kernels of one training step (Trainer, gemm = "tc"), counted and timed with torch.profiler (CUPTI), no ncu:
    python scratch/train_kernel_count.py [mc|mcr] [scenes]"""
import collections
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from multimodaltraj_2_b200 import ops, synth  # noqa: E402
from multimodaltraj_2_b200.train import Trainer  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
variant = args[0] if args else "mc"
S = int(args[1]) if len(args) > 1 else (1024 if variant == "mc" else 512)
dev = torch.device("cuda")
p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
pos, vis, valid = (torch.from_numpy(a).to(dev) for a in synth.make_crowd(S, 64, seed=synth.SEED))
tr = Trainer(p, 8, 12, 4.0, 0.5, lr=1e-3, gemm="tc", relational=(variant == "mcr"))
for _ in range(3):
    tr.step(pos, vis, valid)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(pos, vis, valid)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower():
        k = e.name.split("(")[0].replace("void ", "")[:72]
        agg[k][0] += 1
        agg[k][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
n = sum(v[0] for v in agg.values())
tot = sum(v[1] for v in agg.values())
own = sum(v[0] for k, v in agg.items() if k.startswith("mmt::"))
own_t = sum(v[1] for k, v in agg.items() if k.startswith("mmt::"))
print(f"{variant} {S} scenes: {n} kernels per step, {tot / 1e3:.2f} ms of kernel time; this library's: {own} kernels, {own_t / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{v[0]:5d} {v[1] / v[0]:8.1f} us {v[1] / 1e3:7.2f} ms {v[1] / tot * 100:5.1f} %  {k}")
if "--gemm-order" in sys.argv:
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "gemm_tf32" in e.name),
                key=lambda e: e.time_range.start)
    d = [round(e.device_time if hasattr(e, "device_time") else e.cuda_time, 1) for e in ev]
    print("gemm_tf32 launches in order (us): forward x19, then per frame backwards [head hn^T dy, mf^T dy]? dW dA W_e")
    print(d[:19])
    print(d[19:19 + 30])
    print(d[-12:])
