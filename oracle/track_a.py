"""CPU oracle, Track A: the reference's as-written tensor algebra (TEST INFRASTRUCTURE ONLY).

Nothing in the product path (``multimodaltraj_2_b200``) may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do.

Every function restates one piece of the TensorFlow-1.14 graph of serenetech90/multimodaltraj_2
in NumPy (fp64 by default, as the reference is fp64 throughout) and cites the reference lines
it follows (paths relative to ``/root/reference``).

Pinning status
--------------
TensorFlow 1.14 cannot run here, and the reference has no tests.  What pins this restatement:
  * ``tests/golden/track_a_ckpt.npz`` -- tensors decoded from the reference's own TF checkpoints
    (``save/g2k_mcrAttn_model_kfold_train_4_0.ckpt-79``): five instantiations where the saved
    TF-evaluated ``Variable[24,8] == weight_c @ cost`` holds to 1e-17, i.e. a TF-computed output
    of ``models/g2k_lstm_mcr.py:122`` (first matmul of the band) with its inputs.
  * ``save/g2k_mcr_model_val_0.ckpt-0``: the saved ``temp_path`` tensors of the ``g2k_lstm_mc``
    forward are exactly zero (``models/g2k_lstm_mc.py:59-69`` -> band == 0).
  * seed-0 weights of every shape the model uses (``weight_v[8,12]``, ``bias_v[10]``, ...).
The remaining lines (attention, cost, per-frame state step) are pinned by code reading only:
"parity unpinned" at the TF-library level (no TF-evaluated outputs with known inputs exist).
"""
from __future__ import annotations

import numpy as np


def softmax_last(x):
    """tf.nn.softmax default axis=-1 (train.py:240,243,248)."""
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=-1, keepdims=True)


def input_build(pos):
    """A.1 batched: X[..., t, i] = ||pos[..., i, t, :]||_2  (train.py:76-87).

    ``pos`` is ``[..., n, T, 2]`` (``node_pos_list`` stacked); returns ``[..., T, n]``.
    The reference slices the *node* axis by frame index (defect F-4); the batched form keeps
    all agents.
    """
    x = np.sqrt(pos[..., 0] * pos[..., 0] + pos[..., 1] * pos[..., 1])
    return np.swapaxes(x, -1, -2)


def embeddings(X, V, W_i, W_ii, vemb_prev=None):
    """A.2 (train.py:178-183,194-195,231).

    X[T,n], V[2,n], W_i[n,D], W_ii[D,T] -> outputs[D+2,D], vrel[2,D], vemb[2,D].
    """
    inputs = W_ii @ (X @ W_i)                       # train.py:179-180  [D,D]
    vemb = V @ W_i                                  # train.py:183      [2,D]
    prev = vemb if vemb_prev is None else vemb_prev
    vrel = prev * vemb                              # train.py:194-195
    outputs = np.concatenate([inputs, vemb], axis=-2)   # train.py:231
    return outputs, vrel, vemb


def stat_mask(D, T, dtype=np.float64):
    """A.3 (train.py:154-155): every row = [0, 1/T, ..., (T-1)/T]."""
    dt = np.dtype(dtype)
    return np.zeros((D, T), dt) + (np.arange(T, dtype=dt) / dt.type(T))[None, :]


def static_context(C, lam, T):
    """A.3 (train.py:110,158): ngh = (lam * C) @ stat_mask, C[D,D] the conv output (an input)."""
    return (lam * C) @ stat_mask(C.shape[-1], T, C.dtype)


def mcr_forward(outputs, rel, ngh, W_v, b_v, W_r, W_c, W_o, lam, pred_len=12):
    """A.4  g2k_lstm_mcr.forward (models/g2k_lstm_mcr.py:99-124).

    outputs[D+2,D] rel[2,D] ngh[D,T] W_v[T,D+2] b_v[D] W_r[T,2] W_c[2P,T] W_o[T,n].
    Returns dict(ngh, Eo, attn, cost, band[2,P,n]).
    """
    ngh_s = lam * ngh                               # :102
    Eo = W_v @ outputs + b_v                        # :105 / :112  bias broadcast over rows
    attn = ngh_s @ (Eo * (W_r @ rel))               # :105-106  [D,D]
    cost = Eo @ ngh_s                               # :112-113  [T,T]
    tmp = (W_c @ cost) @ W_o                        # :122      [2P,n]
    n = W_o.shape[-1]
    band = tmp.reshape(tmp.shape[:-2] + (2, pred_len, n))   # :124 row-major reshape
    return dict(ngh=ngh_s, Eo=Eo, attn=attn, cost=cost, band=band)


def mc_forward(outputs, ngh, W_v, b_v, W_c, W_o, pred_len=12):
    """g2k_lstm_mc.forward (models/g2k_lstm_mc.py:54-69).

    ``tf.gradients(ys=<placeholder>, xs=[E, ngh_var], unconnected_gradients='zero')`` is a
    constant zero (defect F-13) so cost == 0 and band == 0; Eo and ngh_var are still computed.
    """
    Eo = W_v @ outputs + b_v                        # :56
    ngh_var = Eo @ ngh                              # :58
    T = W_c.shape[-1]
    cost = np.zeros(Eo.shape[:-2] + (T, T), Eo.dtype)       # :59-63
    tmp = (W_c @ cost) @ W_o                        # :66-67
    band = tmp.reshape(tmp.shape[:-2] + (2, pred_len, W_o.shape[-1]))  # :69
    return dict(Eo=Eo, ngh_var=ngh_var, cost=cost, band=band)


def gsk_forward(outputs, ngh, W_v, b_v, W_c, W_o):
    """gsk_lstm_cell (models/gsk_lstm_cell.py:52-65), the only shape-consistent reading.

    outputs[D,D] ngh[12,D] W_v[12,D] b_v[D] W_c[16,12] W_o[D,n]; cost = relu(d ngh/d ngh) = 1
    (defect F-13); band = reshape((W_c @ 1[12,D]) @ W_o, (2, 8, n)) -- the pred_len-8 era shape
    (the committed ``reshape(..., (2, 12, n))`` of a [16,n] tensor cannot run).
    """
    Eo = W_v @ outputs + b_v                        # :52
    ngh_h = Eo * ngh                                # :54
    cost = np.maximum(np.ones_like(ngh_h), 0.0)     # :55-60
    tmp = (W_c @ cost) @ W_o                        # :62-63  [16,n]
    band = tmp.reshape(tmp.shape[:-2] + (2, 8, W_o.shape[-1]))
    return dict(Eo=Eo, ngh=ngh_h, cost=cost, band=band)


def frame_state_step(A, Hs):
    """A.5 per-frame state step (train.py:240-254).

    A = attn[D,D], Hs[D,H].  Returns (a[D,D], Hs'[D,H], adj[D,1]).
    ``tf.cumsum`` default axis 0, ``tf.nn.softmax`` default axis -1.  The reference softmaxes the
    placeholder's *default random* tensor at :243-244 (defect F-9); this takes the evident
    intent: softmax of the carried Hs.
    """
    eA = np.exp(A)
    a = softmax_last(eA / np.cumsum(eA, axis=-2))   # :240
    Hs = softmax_last(Hs)                           # :243-244
    Hs = a @ Hs                                     # :247  aggregation
    adj = np.sum(softmax_last(Hs), axis=-1, keepdims=True)  # :248-249 softmax(Hs) @ ones[H,1]
    Hs = adj * Hs                                   # :252
    return a, Hs, adj


def band_to_pred(band):
    """train.py:254: pred = band.transpose(2,1,0) -> [..., n, P, 2] (leading batch dims kept)."""
    lead = tuple(range(band.ndim - 3))
    return np.transpose(band, lead + (band.ndim - 1, band.ndim - 2, band.ndim - 3))


def mcr_scene_step(X, V, C, Hs, w, lam=0.0005, pred_len=12, vemb_prev=None):
    """One full Track-A scene-frame: A.2 + A.3 + A.4 + A.5 in the order of train.py:178-254.

    ``w`` is a dict with W_i, W_ii, W_v, b_v, W_r, W_c, W_o.  Returns dict with attn, cost, band,
    pred[n,P,2], a, Hs, adj, vemb.
    """
    T = X.shape[-2]
    outputs, vrel, vemb = embeddings(X, V, w["W_i"], w["W_ii"], vemb_prev)
    ngh = static_context(C, lam, T)
    f = mcr_forward(outputs, vrel, ngh, w["W_v"], w["b_v"], w["W_r"], w["W_c"], w["W_o"], lam, pred_len)
    a, Hs2, adj = frame_state_step(f["attn"], Hs)
    band = f["band"]
    pred = band_to_pred(band)
    return dict(attn=f["attn"], cost=f["cost"], band=band, pred=pred, a=a, Hs=Hs2, adj=adj, vemb=vemb,
                Eo=f["Eo"], ngh=f["ngh"])


def mcr_scene_loop(X, V, C, Hs, w, lam=0.0005, pred_len=12, frames=1):
    """Reference execution shape: one scene at a time, one frame at a time, fp64
    (train.py:71-254).  X[S,T,n] V[S,2,n] C[S,D,D] Hs[S,D,H]; used as the timed CPU baseline for
    the Track-A step."""
    S = X.shape[0]
    out = []
    for s in range(S):
        h = Hs[s]
        r = None
        for _ in range(frames):
            r = mcr_scene_step(X[s], V[s], C[s], h, w, lam, pred_len)
            h = r["Hs"]
        out.append(r)
    return {k: np.stack([o[k] for o in out]) for k in out[0]}
