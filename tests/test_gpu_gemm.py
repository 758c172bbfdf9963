"""mmt_gemm_tf32 (TMA tensor maps -> tcgen05.mma.kind::tf32 -> TMEM) against fp64 matmul: every operand layout, ragged
tile edges, split-K accumulation, the shapes of the training step's contractions."""
import numpy as np
import pytest
import torch

from multimodaltraj_2_b200 import ops

pytestmark = pytest.mark.gpu


def _mat(rng, rows, cols, pad=0):
    """[rows, cols] view of a [rows, cols + pad] buffer: a leading dimension that is not the width"""
    ld = (cols + pad + 3) // 4 * 4
    buf = torch.as_tensor(rng.standard_normal((rows, ld)).astype(np.float32)).cuda()
    return buf[:, :cols]


@pytest.mark.parametrize("transA", [False, True])
@pytest.mark.parametrize("transB", [False, True])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (130, 72, 100), (320, 384, 5000), (1000, 320, 384), (5, 8, 7), (257, 129, 33)])
def test_gemm_tf32_matches_fp64(cuda, M, N, K, transA, transB):
    rng = np.random.default_rng(M + 3 * N + 7 * K)
    A = _mat(rng, K, M, pad=4) if transA else _mat(rng, M, K, pad=8)
    B = _mat(rng, N, K, pad=4) if transB else _mat(rng, K, N)
    C = ops.gemm_tf32(A, B, transA=transA, transB=transB)
    torch.cuda.synchronize()
    a, b = A.double().cpu().numpy(), B.double().cpu().numpy()
    want = (a.T if transA else a) @ (b.T if transB else b)
    # tf32 operands keep 10 explicit mantissa bits (the tensor core drops the rest): up to 2^-10 relative per operand,
    # accumulated in fp32 over K -> compare against the sum of |a||b| that bounds it
    bound = (np.abs(a.T if transA else a) @ np.abs(b.T if transB else b)) * 2.0 ** -9 + 1e-6
    err = np.abs(C.cpu().numpy() - want)
    assert (err <= bound).all(), (float(err.max()), float((err / bound).max()))
    assert err.max() / np.abs(want).max() < 3e-3


def test_gemm_tf32_accumulates_and_scales(cuda):
    rng = np.random.default_rng(3)
    A, B = _mat(rng, 4096, 320), _mat(rng, 4096, 384)          # dW += alpha * A^T B, split over K
    C0 = torch.as_tensor(rng.standard_normal((320, 384)).astype(np.float32)).cuda()
    C = C0.clone()
    ops.gemm_tf32(A, B, transA=True, out=C, alpha=0.5, accumulate=True)
    want = C0.double().cpu().numpy() + 0.5 * A.double().cpu().numpy().T @ B.double().cpu().numpy()
    assert np.abs(C.cpu().numpy() - want).max() / np.abs(want).max() < 3e-3
    # no accumulate with split-K: the output is cleared first, whatever it held
    C2 = torch.full((320, 384), 1e6, device="cuda")
    ops.gemm_tf32(A, B, transA=True, out=C2)
    full = (want - C0.double().cpu().numpy()) * 2
    assert np.abs(C2.cpu().numpy() - full).max() / np.abs(full).max() < 3e-3
    # exactly representable operands (small integers) -> exact result, any layout
    Ai = torch.randint(-4, 5, (200, 64), device="cuda").float()
    Bi = torch.randint(-4, 5, (64, 96), device="cuda").float()
    assert torch.equal(ops.gemm_tf32(Ai, Bi), Ai @ Bi)
    assert torch.equal(ops.gemm_tf32(Ai.t().contiguous(), Bi.t().contiguous(), transA=True, transB=True), Ai @ Bi)


def test_gemm_tf32_rejects_bad_arguments(cuda):
    A = torch.zeros((8, 6), device="cuda")        # row stride 6: not a multiple of 4 floats
    with pytest.raises(RuntimeError, match="leading dimensions"):
        ops.gemm_tf32(A, torch.zeros((6, 8), device="cuda"))
    with pytest.raises(ValueError):
        ops.gemm_tf32(torch.zeros((8, 8), device="cuda"), torch.zeros((4, 8), device="cuda"))


def test_aggregate_transpose_is_the_adjoint(cuda):
    rng = np.random.default_rng(5)
    S, N, Cc = 7, 24, 256
    att = rng.random((S, N, N)).astype(np.float32) * (rng.random((S, N, N)) < 0.2)
    d = rng.standard_normal((S, N, Cc)).astype(np.float32)
    got = ops.aggregate_transpose(torch.as_tensor(att).cuda(), torch.as_tensor(d).cuda()).cpu().numpy()
    want = np.einsum("sij,sic->sjc", att.astype(np.float64), d.astype(np.float64))
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
