"""CPU estimate: how much of the bf16 mode's ADE/FDE error is operand rounding (and what fp16 / tf32-like operands would give).
The oracle's rollout with a rounding function applied where the fused kernel rounds: the gate-GEMM operands ([e | h | mh], W),
the aggregation operands (un-normalised attention numerators, h, c images).  Transcendentals stay exact fp32, so the
difference to the measured GPU error is what the approximate tanh / ex2 contribute.
    python scratch/operand_rounding_emulate.py [scenes]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import track_b as o_b  # noqa: E402
from multimodaltraj_2_b200 import synth  # noqa: E402

f32 = np.float32


def r_bf16(x):
    u = np.ascontiguousarray(x, f32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(f32)


def r_f16(x):
    return np.asarray(x, f32).astype(np.float16).astype(f32)


def r_tf32(x):
    u = np.ascontiguousarray(x, f32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(f32)


def rollout(pos, vis, valid, p, rnd, T=8, P=12, r2=4.0, inv=0.5):
    S, N = valid.shape
    U = p["w_If"].shape[0]
    h = np.zeros((S, N, U), f32); c = np.zeros((S, N, U), f32)
    params = np.zeros((S, N, P, 5), f32)
    cur = pos[:, :, 0].astype(f32); prev = cur
    W = rnd(p["W"])
    for t in range(T + P - 1):
        if t < T:
            cur = pos[:, :, t].astype(f32); v_t = vis[:, :, t].astype(f32)
        else:
            v_t = vis[:, :, T - 1].astype(f32)
        disp = (cur - prev).astype(f32) if t > 0 else np.zeros_like(cur)
        x = np.concatenate([disp, v_t], -1).astype(f32)
        kern, adj, _ = o_b.pairwise_adj(cur, valid, r2, inv)
        m = adj.astype(bool)
        num = rnd(np.where(m, np.exp(kern, dtype=f32), f32(0)))           # the kernel's operand: un-normalised numerators
        den = np.where(m, np.exp(kern, dtype=f32), f32(0)).sum(-1, keepdims=True, dtype=f32)
        inv_den = np.where(den > 0, f32(1) / np.where(den > 0, den, f32(1)), f32(0))
        mh = (np.matmul(num, rnd(h)) * inv_den).astype(f32)
        mc = (np.matmul(num, rnd(c)) * inv_den).astype(f32)
        e = np.maximum(np.matmul(x, p["W_e"]) + p["b_e"], f32(0)).astype(f32)
        u = rnd(np.concatenate([e, h, mh], -1))
        z = (np.matmul(u, W) + p["b"]).astype(f32)
        i, j, o = z[..., :U], z[..., U:2 * U], z[..., 2 * U:]
        g = o_b.sigmoid(i + p["w_If"] * mc + p["w_It"] * c)
        tj = np.tanh(j, dtype=f32)
        c_f = ((f32(1) - g) * mc + g * tj).astype(f32)
        c_t = ((f32(1) - g) * c + g * tj).astype(f32)
        q = o_b.sigmoid(o + p["w_Of"] * c_f + p["w_Ot"] * c_t)
        m_f = (q * np.tanh(c_f, dtype=f32)).astype(f32)
        m_t = (q * np.tanh(c_t, dtype=f32)).astype(f32)
        v = valid.astype(bool)[..., None]
        h, c, m_f = np.where(v, m_t, f32(0)), np.where(v, c_t, f32(0)), np.where(v, m_f, f32(0))
        prev = cur
        if t >= T - 1:
            y = o_b.head(h, m_f, p)
            y = np.where(v, y, f32(0)).astype(f32)
            params[:, :, t - (T - 1)] = y
            cur = (cur + y[..., :2]).astype(f32)
    return params


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    N, T, P, K = 64, 8, 12, 20
    p = synth.init_params(seed=0)
    pos, vis, valid = synth.make_crowd(S, N, seed=synth.SEED)
    eps = o_b.philox_eps(0xB200, S, N, K, P)
    want = o_b.forecast(pos, vis, valid, p, eps, T, P)
    v = valid.astype(bool)
    for name, rnd in (("none", lambda x: np.asarray(x, f32)), ("bf16", r_bf16), ("tf32 (10-bit)", r_tf32), ("fp16", r_f16)):
        par = rollout(pos, vis, valid, p, rnd)
        ade, fde, *_ = o_b.decode_score(par, eps, pos[:, :, T - 1], pos[:, :, T:T + P], valid)
        print(f"{name:14s} max|dADE| {np.abs(ade[v] - want['ade'][v]).max():.3e}  max|dFDE| {np.abs(fde[v] - want['fde'][v]).max():.3e}")


if __name__ == "__main__":
    main()
