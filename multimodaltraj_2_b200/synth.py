"""Synthetic ETH/UCY-shaped crowds and random-init weights.

Generator law (numpy Philox, seed 0xB200, float32): per-agent start ~ U([-1,1]^2) * half_extent
(normalised-std units), constant velocity ~ N(0, 0.16^2) per axis per frame (UCY step norm
0.135-0.164) + N(0, 0.02^2) jitter, T+P frames; vislet = unit heading of the
velocity + N(0, 0.1^2).  ``ragged`` draws the number of valid agents per scene from a clipped
Poisson around the UCY peds/frame means so padded slots exercise the ``valid`` mask.
"""
from __future__ import annotations

import numpy as np

SEED = 0xB200


def make_crowd(S, N, T=8, P=12, seed=SEED, half_extent=8.0, ragged=False):
    rng = np.random.Generator(np.random.Philox(seed))
    F = T + P
    start = (rng.random((S, N, 1, 2), dtype=np.float32) * 2 - 1) * np.float32(half_extent)
    vel = rng.standard_normal((S, N, 1, 2), dtype=np.float32) * np.float32(0.16)
    t = np.arange(F, dtype=np.float32).reshape(1, 1, F, 1)
    pos = start + vel * t + rng.standard_normal((S, N, F, 2), dtype=np.float32) * np.float32(0.02)
    head = vel / np.maximum(np.linalg.norm(vel, axis=-1, keepdims=True), np.float32(1e-6))
    vis = head + rng.standard_normal((S, N, T, 2), dtype=np.float32) * np.float32(0.1)
    valid = np.ones((S, N), np.uint8)
    if ragged:
        n = np.clip(rng.poisson(min(9.0, N / 2), size=S), 1, N)
        valid = (np.arange(N)[None, :] < n[:, None]).astype(np.uint8)
    return pos.astype(np.float32), vis.astype(np.float32), valid


def init_params(seed=0, E=64, U=128, He=128, scale=1.0):
    """Random-init weights N(0,1) * fan_in^-1/2 (biases / peepholes small), float32 numpy dict."""
    rng = np.random.Generator(np.random.Philox(seed))

    def w(*shape, fan_in=None):
        fan_in = shape[0] if fan_in is None else fan_in
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(scale / np.sqrt(fan_in))).astype(np.float32)

    K = E + 2 * U
    p = dict(
        W_e=w(4, E), b_e=w(E, fan_in=16),
        W=w(K, 3 * U), b=w(3 * U, fan_in=16),
        w_If=w(U, fan_in=4), w_It=w(U, fan_in=4), w_Of=w(U, fan_in=4), w_Ot=w(U, fan_in=4),
        W_h=w(2 * U, 5), b_h=w(5, fan_in=16),
        W1=w(2 * U, He), b1=w(He, fan_in=16), W2=w(He, He), b2=w(He, fan_in=16),
        w_out=w(He), b_out=np.float32(0.1) * np.ones((), np.float32),
    )
    # keep predicted steps ETH/UCY-sized: mean displacement ~0.1, sigma ~ exp(-2)
    p["W_h"] *= np.float32(0.1)
    p["b_h"] = np.array([0.0, 0.0, -2.0, -2.0, 0.0], np.float32)
    return p
