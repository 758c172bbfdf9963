"""A/B of two builds of the library: the rollout kernel's parameters on the same inputs (run once per MMT_LIB, compare dumps)."""
import sys, os, torch, numpy as np
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import ops, synth
dev = torch.device('cuda')
tag = sys.argv[1]
outs = {}
for S, N in ((512, 64), (37, 16), (6, 128), (50, 8), (9, 32)):
    pos, vis, valid = synth.make_crowd(S, N, seed=3, half_extent=4.0, ragged=True)
    p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
    out = ops.rollout_bf16(*(torch.from_numpy(a).to(dev) for a in (pos, vis, valid)), p)
    torch.cuda.synchronize()
    outs[f'{S}x{N}'] = out.cpu().numpy()
np.savez(f'gpurun_out/ro_ab_{tag}.npz', **outs)
if len(sys.argv) > 2:
    ref = np.load(f'gpurun_out/ro_ab_{sys.argv[2]}.npz')
    for k, v in outs.items():
        d = np.abs(v - ref[k])
        print(k, 'max|diff|', float(d.max()), 'max|ref|', float(np.abs(ref[k]).max()), 'nan', int(np.isnan(v).sum()))
