import sys, torch, ctypes as C
sys.path.insert(0, '.')
from multimodaltraj_2_b200 import _lib, ops, synth
lib = _lib.load()
S, N = 32768, int(sys.argv[1]) if len(sys.argv) > 1 else 64
if N != 64: S = 32768 * 64 * 64 // (N * N)
dev = torch.device('cuda')
pos = (torch.rand((S, N, 2), device=dev) * 16 - 8)
valid = torch.ones((S, N), dtype=torch.uint8, device=dev)
kern = torch.empty((S, N, N), device=dev); adj = torch.empty((S, N, N), dtype=torch.uint8, device=dev)
deg = torch.empty((S, N), dtype=torch.int32, device=dev)
def run(k=True, a=True, d=False):
    _lib.check(lib.mmt_pairwise_adj_f32(C.c_void_p(pos.data_ptr()), C.c_void_p(valid.data_ptr()), S, N, 4.0, 0.5,
        C.c_void_p(kern.data_ptr()) if k else None, C.c_void_p(adj.data_ptr()) if a else None, C.c_void_p(deg.data_ptr()) if d else None, None))
for cfg in ((True, True, False), (True, True, True), (True, False, False), (False, True, False)):
    for _ in range(3): run(*cfg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run(*cfg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    b = S * ((4 * N * N if cfg[0] else 0) + (N * N if cfg[1] else 0) + 9 * N + (4 * N if cfg[2] else 0))
    print(f'N={N} S={S} kern={cfg[0]} adj={cfg[1]} deg={cfg[2]}: {ms*1e3:.1f} us  {b/ms/1e6:.0f} GB/s')
# reference: torch fill (pure write) and copy
x = torch.empty((S * N * N,), device=dev)
for name, fn, nb in (('fill', lambda: x.fill_(1.0), 4), ('copy', lambda: x.copy_(kern.view(-1)), 8)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    print(f'torch {name}: {ms*1e3:.1f} us {x.numel()*nb/ms/1e6:.0f} GB/s')
