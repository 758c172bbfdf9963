import sys, torch
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import ops
S, N, P, K = 4096, 64, 12, 20
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(0)
par = torch.randn((S, N, P, 5), device=dev, generator=g) * 0.3
par[..., 2:4] = par[..., 2:4].exp(); par[..., 4] = par[..., 4].tanh()
lo = torch.randn((S, N, 2), device=dev, generator=g); gt = torch.randn((S, N, P, 2), device=dev, generator=g)
valid = torch.ones((S, N), dtype=torch.uint8, device=dev)
for want_traj in (True, False):
    for _ in range(3): ops.decode_score(par, lo, gt, valid, K, seed=1, want_all=False, want_traj=want_traj)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.decode_score(par, lo, gt, valid, K, seed=1, want_all=False, want_traj=want_traj)
    e1.record(); torch.cuda.synchronize()
    print(f'decode want_traj={want_traj}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us')
