"""Phase timeline of the fused rollout kernel (CTA 0, worker thread 0) + kernel time alone."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import ops, synth
S, N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device('cuda')
pos, vis, valid = synth.make_crowd(S, N, seed=synth.SEED)
p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
pos, vis, valid = (torch.from_numpy(a).to(dev) for a in (pos, vis, valid))
tl = torch.zeros((64, 32), dtype=torch.int64, device=dev)
out = ops.rollout_bf16(pos, vis, valid, p, timeline=tl)
torch.cuda.synchronize()
t = tl.cpu().numpy()
names = {1: 'att', 2: 'e+sync', 3: 'wait_agg', 4: 'mh_conv', 5: 'wait_acc0', 7: 'epi0+wait1', 9: 'epi1+wait2', 11: 'epi2+wait3', 13: 'epi3', 14: 'head'}
idx = [0, 1, 2, 3, 4, 5, 7, 9, 11, 13, 14]
print('step ' + ' '.join(f'{names[i]:>11s}' for i in idx[1:]) + '       total')
for s in range(40):
    r = t[s]
    if r[0] == 0: break
    d = [r[idx[k]] - r[idx[k - 1]] for k in range(1, len(idx))]
    print(f'{s:4d} ' + ' '.join(f'{x:11d}' for x in d) + f' {r[14] - r[0]:11d}')
print('MMA thread (relative to worker step start): att_rdy agg_issued e_rdy p0_start mh_rdy p0_commit p1_start p1_commit p2_start p2_commit p3_start p3_commit | W wait')
for s in range(24):
    r = t[s]
    if r[0] == 0: break
    m = r[16:]
    order = [0, 1, 2, 3, 12, 4, 5, 6, 7, 8, 9, 10]
    print(f'{s:4d} ' + ' '.join(f'{m[k] - r[0]:7d}' for k in order) + f' | Wwait {m[13]:5d} {m[14]:5d} |   worker: att_arr {r[1]-r[0]} mh_arr {r[4]-r[0]} acc0 {r[5]-r[0]} acc1 {r[7]-r[0]} acc2 {r[9]-r[0]} acc3 {r[11]-r[0]} end {r[14]-r[0]}')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): ops.rollout_bf16(pos, vis, valid, p, out=out)
torch.cuda.synchronize()
e0.record()
for _ in range(10): ops.rollout_bf16(pos, vis, valid, p, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
R = S * N
fl = R * 19 * (2 * 320 * 384 + 2 * N * 256)
print(f'rollout kernel {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s algorithmic  {R / ms / 1e3:.1f} M agent-traj/s')
