import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import _lib, ops, synth
lib = _lib.load()
raw = C.CDLL(str(_lib.lib_path()))
dev = torch.device('cuda')
R = 4096 * 64
emit = len(sys.argv) > 1 and sys.argv[1] == 'emit'
p = ops.CellParams.from_numpy(synth.init_params(seed=0), dev).pack()
x = torch.randn((R, 4), device=dev) * 0.3
hb = (torch.randn((R * 128,), device=dev) * 0.5).to(torch.bfloat16)
mhb, mcb = hb.clone(), hb.clone()
c = torch.randn((R * 128,), device=dev) * 0.5
valid = torch.ones(R, dtype=torch.uint8, device=dev)
hbo, co = torch.empty_like(hb), torch.empty_like(c)
cur = torch.randn((R, 2), device=dev); par = torch.empty((R, 5), device=dev); nxt = torch.empty((R, 2), device=dev)
dbg = torch.zeros((296, 16, 16), dtype=torch.int64, device=dev)
w = p.c_cell()
vp = lambda t: C.c_void_p(t.data_ptr())
fn = raw.mmt_debug_cell_tc_timeline; fn.restype = C.c_int
for _ in range(3):
    rc = fn(vp(x), vp(hb), vp(c), vp(mhb), vp(mcb), vp(valid), C.byref(w), R, vp(hbo), vp(co), vp(cur),
            vp(par) if emit else None, vp(nxt) if emit else None, vp(dbg), None)
    assert rc == 0, lib.mmt_last_error()
torch.cuda.synchronize()
d = dbg.cpu().numpy()
names = ['start', 'built'] + [f'{n}{p}' for p in range(4) for n in ('wait', 'full', 'done')] + ['epi_end']
for cta in (0, 1, 150):
    for ti in (0, 3, 5):
        t = d[cta, ti]
        if t[0] == 0: continue
        nxt_start = d[cta, ti + 1, 0] if d[cta, ti + 1, 0] else t[14]
        print(f'cta {cta} tile_iter {ti}: build {t[1]-t[0]:6d} |', ' '.join(f'p{p}: wait {t[3+3*p]-t[2+3*p]:6d} epi {t[4+3*p]-t[3+3*p]:6d}' for p in range(4)), f'| head {nxt_start - t[14]:6d} total {nxt_start - t[0]:7d}')
tot = []
for cta in range(296):
    for ti in range(6):
        if d[cta, ti, 0] and d[cta, ti + 1, 0]: tot.append([d[cta, ti, 1] - d[cta, ti, 0]] + [d[cta, ti, 3 + 3 * p] - d[cta, ti, 2 + 3 * p] for p in range(4)] + [d[cta, ti, 4 + 3 * p] - d[cta, ti, 3 + 3 * p] for p in range(4)] + [d[cta, ti + 1, 0] - d[cta, ti, 0]])
tot = np.array(tot); print('mean: build', tot[:, 0].mean(), 'waits', tot[:, 1:5].mean(0), 'epis', tot[:, 5:9].mean(0), 'tile total', tot[:, 9].mean())
