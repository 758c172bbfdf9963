import sys, torch
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import ops, synth
from multimodaltraj_2_b200.train import Trainer
dev = torch.device('cuda')
for lr in (0.005, 0.001):
    pos, vis, valid = (torch.from_numpy(a).to(dev) for a in synth.make_crowd(512, 64, seed=1))
    tr = Trainer(ops.CellParams.from_numpy(synth.init_params(seed=0), dev), lr=lr)
    ls = [float(tr.step(pos, vis, valid)) for _ in range(40)]
    print(lr, ' '.join(f'{x:.3f}' for x in ls[::3]))
