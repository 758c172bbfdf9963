"""One hot-path step of every mode, launched eagerly (no CUDA graph), for ncu:
  python scratch/step_profile.py [f16] [bf16] [bf16x3] [f32] [train] [train-mcr]
C3 batch (4096 scenes x 64 agents) for the inference modes (f32: 1024 scenes), 512 scenes for the training step."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from multimodaltraj_2_b200 import ops, synth  # noqa: E402
from multimodaltraj_2_b200.train import Trainer  # noqa: E402

what = sys.argv[1:] or ["bf16", "bf16x3", "train"]
dev = torch.device("cuda")
p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
PREC = {"bf16": ops.PREC_BF16, "bf16x3": ops.PREC_BF16X3, "f32": ops.PREC_F32, "f16": ops.PREC_F16}
for mode in what:
    if mode in ("train", "train-mcr"):
        S = 512
        pos, vis, valid = (torch.from_numpy(a).to(dev) for a in synth.make_crowd(S, 64, seed=synth.SEED))
        tr = Trainer(p, 8, 12, 4.0, 0.5, lr=1e-3, gemm="tc", relational=(mode == "train-mcr"))
        for _ in range(2):
            tr.step(pos, vis, valid)
    else:
        S = 1024 if mode == "f32" else 4096
        pos, vis, valid = (torch.from_numpy(a).to(dev) for a in synth.make_crowd(S, 64, seed=synth.SEED))
        fc = ops.Forecaster(p, S, 64, 8, 12, 20, prec=PREC[mode], seed=1, device=dev)
        for _ in range(3):
            fc(pos, vis, valid)
    torch.cuda.synchronize()
    print("done", mode, flush=True)
