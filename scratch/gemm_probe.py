"""Layout probe for mmt_gemm_tf32: identity / one-hot operands show where every element lands."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from multimodaltraj_2_b200 import ops

torch.manual_seed(0)
for transA in (False, True):
    for transB in (False, True):
        M = N = K = 128
        A = torch.eye(128, device="cuda")
        Bm = (torch.arange(K, device="cuda")[:, None] * 8 + torch.arange(N, device="cuda")[None, :] % 8).float()  # op(B)[k,n] = 1000k+n
        B = Bm.t().contiguous() if transB else Bm
        C = ops.gemm_tf32(A, B, transA=transA, transB=transB)
        torch.cuda.synchronize()
        ok = torch.equal(C, Bm)
        print(f"transA={transA} transB={transB}: identity x B == B: {ok}")
        if not ok:
            c = C.cpu().numpy()
            print("   C[0,:10] =", c[0, :10].tolist(), " C[1,:4] =", c[1, :4].tolist(), " C[:6,0] =", c[:6, 0].tolist())
            print("   C[0,32:36] =", c[0, 32:36].tolist(), "C[8,:4] =", c[8, :4].tolist(), "nonzero frac", float((c != 0).mean()))
        # A = arange, B = identity
        Am = (torch.arange(M, device="cuda")[:, None] * 8 + torch.arange(K, device="cuda")[None, :] % 8).float()   # op(A)[m,k]
        A2 = Am.t().contiguous() if transA else Am
        I = torch.eye(128, device="cuda")
        C = ops.gemm_tf32(A2, I, transA=transA, transB=transB)
        torch.cuda.synchronize()
        ok = torch.equal(C, Am)
        print(f"transA={transA} transB={transB}: A x identity == A: {ok}")
        if not ok:
            c = C.cpu().numpy()
            print("   C[0,:10] =", c[0, :10].tolist(), " C[1,:4] =", c[1, :4].tolist(), " C[:6,0] =", c[:6, 0].tolist())
            print("   C[0,32:36] =", c[0, 32:36].tolist(), "C[8,:4] =", c[8, :4].tolist(), "nonzero frac", float((c != 0).mean()))
for (M, N, K) in ((128, 128, 32), (128, 128, 256), (256, 256, 64)):
    for transA in (False, True):
        for transB in (False, True):
            a = torch.randn((K, M) if transA else (M, K), device="cuda")
            b = torch.randn((N, K) if transB else (K, N), device="cuda")
            C = ops.gemm_tf32(a, b, transA=transA, transB=transB)
            want = (a.t() if transA else a).double() @ (b.t() if transB else b).double()
            print(f"M{M} N{N} K{K} tA={int(transA)} tB={int(transB)}: max err {float((C.double() - want).abs().max()):.4f}")
