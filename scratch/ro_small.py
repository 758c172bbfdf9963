import sys, torch, numpy as np
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import ops, synth
dev = torch.device('cuda')
for S, N in ((5, 64), (9, 16), (3, 128), (20, 8)):
    pos, vis, valid = synth.make_crowd(S, N, seed=3, half_extent=4.0, ragged=True)
    p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
    out = ops.rollout_bf16(*(torch.from_numpy(a).to(dev) for a in (pos, vis, valid)), p)
    torch.cuda.synchronize()
    print(S, N, float(out.abs().sum()))
