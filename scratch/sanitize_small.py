"""Small invocation of every tcgen05 / TMA kernel for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python scratch/sanitize_small.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from multimodaltraj_2_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda")
p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
for S, N in ((5, 64), (3, 16), (2, 128)):
    pos, vis, valid = (torch.from_numpy(a).to(dev) for a in synth.make_crowd(S, N, seed=3, half_extent=4.0, ragged=True))
    for prec, rel in ((ops.PREC_BF16, False), (ops.PREC_BF16X3, False), (ops.PREC_BF16, True), (ops.PREC_F32, False)):
        fc = ops.Forecaster(p, S, N, 8, 12, 20, relational=rel, prec=prec, seed=1, device=dev)
        o = fc(pos, vis, valid)
        torch.cuda.synchronize()
        print("forecast", S, N, prec, rel, float(o["best_ade"].sum()), flush=True)
A = torch.randn((300, 200), device=dev)
B = torch.randn((200, 136), device=dev)
print("gemm", float(ops.gemm_tf32(A, B).sum()), float(ops.gemm_tf32(A.t().contiguous(), B, transA=True).sum()), flush=True)
torch.cuda.synchronize()
print("SANITIZE_DONE")
