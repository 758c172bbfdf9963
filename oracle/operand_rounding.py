"""Test infrastructure (oracle): the Track-B rollout of oracle/track_b.py with a rounding function applied where the fused tensor-core
kernel rounds its operands -- the gate-GEMM operands ([e | h | mh], W) and the aggregation operands (un-normalised attention
numerators, the h and c images).  Transcendentals stay exact fp32.  It answers, on the CPU, how much of a reduced-precision mode's
ADE / FDE error is operand rounding (bf16: all of it) and what another operand format would leave.
Only tests/ and scratch/ import this module."""
import numpy as np

import track_b as o_b

f32 = np.float32


def r_none(x):
    return np.asarray(x, f32)


def r_bf16(x):
    """round to nearest even at 8 significand bits (bfloat16)"""
    u = np.ascontiguousarray(x, f32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(f32)


def r_f16(x):
    """IEEE half: 11 significand bits, narrow exponent"""
    return np.asarray(x, f32).astype(np.float16).astype(f32)


def r_tf32(x):
    """11 significand bits with the fp32 exponent"""
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
