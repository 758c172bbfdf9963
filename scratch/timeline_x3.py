"""Per-tile phase times (clock64 of worker thread 0) of the fp32-state tcgen05 cell kernel: `python scratch/timeline_x3.py [x3|bf16]`."""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import _lib, ops, synth
lib = _lib.load()
raw = C.CDLL(str(_lib.lib_path()))
dev = torch.device('cuda')
R = 4096 * 64
x3 = 0 if (len(sys.argv) > 1 and sys.argv[1] == 'bf16') else 1
p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev).pack().pack_x3()
x = torch.randn((R, 4), device=dev) * 0.3
h, c, mh, mc = (torch.randn((R, 128), device=dev) * 0.5 for _ in range(4))
valid = torch.ones(R, dtype=torch.uint8, device=dev)
ho, co = torch.empty_like(h), torch.empty_like(c)
cur = torch.randn((R, 2), device=dev); par = torch.empty((R, 5), device=dev); nxt = torch.empty((R, 2), device=dev)
dbg = torch.zeros((296, 16, 16), dtype=torch.int64, device=dev)
w = p.c_cell()
vp = lambda t: C.c_void_p(t.data_ptr())
fn = raw.mmt_debug_cell_tc_f32state_timeline; fn.restype = C.c_int
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(4):
    if i == 3: e0.record()
    rc = fn(vp(x), vp(h), vp(c), vp(mh), vp(mc), vp(valid), C.byref(w), R, vp(ho), vp(co), vp(cur), vp(par), vp(nxt), x3, vp(dbg), None)
    assert rc == 0, lib.mmt_last_error()
e1.record(); torch.cuda.synchronize()
print('x3' if x3 else 'bf16', 'kernel ms', e0.elapsed_time(e1))
d = dbg.cpu().numpy()
tot = []
ncta = 148 if x3 else 296
for cta in range(ncta):
    for ti in range(12):
        if d[cta, ti, 0] and d[cta, ti + 1, 0]:
            t = d[cta, ti]
            tot.append([t[1] - t[0]] + [t[3 + 3 * p] - t[2 + 3 * p] for p in range(4)] + [t[4 + 3 * p] - t[3 + 3 * p] for p in range(4)] + [t[14] - t[4 + 9], d[cta, ti + 1, 0] - t[14], d[cta, ti + 1, 0] - t[0]])
tot = np.array(tot)
print('tiles sampled', len(tot))
print('mean clk: build', int(tot[:, 0].mean()), '| wait for acc p0..p3', tot[:, 1:5].mean(0).astype(int), '| epilogue p0..p3', tot[:, 5:9].mean(0).astype(int),
      '| after last epi', int(tot[:, 9].mean()), '| head', int(tot[:, 10].mean()), '| tile total', int(tot[:, 11].mean()))
