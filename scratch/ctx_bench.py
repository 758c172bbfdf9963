"""Timing of the static-context kernels at the reference's image size (576 x 720 x 3, D = 16) and of the metric scores."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import ops
dev = torch.device('cuda')
H, W, C, D, T = 576, 720, 3, 16, 8
g = np.random.Generator(np.random.Philox(0))
img = torch.from_numpy(g.integers(0, 256, (H, W, C)).astype(np.float32)).to(dev)
filt = torch.from_numpy(g.standard_normal((H + 3 - D, W + 2 - D, C)).astype(np.float32)).to(dev)
for _ in range(3): ops.static_context(img, filt, D, T, 0.0005)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): ops.static_context(img, filt, D, T, 0.0005)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
fl = 2.0 * D * D * (H + 3 - D) * (W + 2 - D) * C
print(f'static_context {H}x{W}x{C} D={D}: {ms*1e3:.1f} us per call, {fl/ms/1e9:.2f} TFLOP/s fp32 FMA ({fl/1e9:.2f} GFLOP), bytes read once {(img.numel()+filt.numel())*4/1e6:.1f} MB')
n, P = 262144, 12
pred = torch.rand((n, P, 2), device=dev); gt = pred + 0.02 * torch.randn((n, P, 2), device=dev)
Hm = torch.tensor([[0.028, 0.002, -3.1], [-0.001, 0.023, -2.2], [0.0003, -0.0001, 1.0]], device=dev)
for _ in range(3): ops.ade_fde_world(pred, gt, Hm)
torch.cuda.synchronize(); e0.record()
for _ in range(20): ops.ade_fde_world(pred, gt, Hm)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
by = n * (P * 16 + 8)
print(f'ade_fde_world n={n} P={P}: {ms*1e3:.1f} us per call, {by/ms/1e6:.0f} GB/s of {by/1e6:.1f} MB algorithmic')
