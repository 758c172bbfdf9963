"""CPU oracle: the reference's three ADE/FDE reductions + the standard best-of-K definition
(TEST INFRASTRUCTURE ONLY -- never imported by the product path).  SURVEY App. A.6."""
from __future__ import annotations

import numpy as np


def train_val_scores(pred, targets, pred_len=12, n_batch_frames=None, n_targets=None):
    """train.py:639-674 for one batch.  pred[n,P,2]; targets: list of [L_i,2] arrays (dict order).

    euc_i = sigma_max(pred[i,:L] - tgt_i[:L]) / 12   (``np.linalg.norm(M, ord=2)`` on a matrix is
    the spectral norm, defect F-7; when L < P the reference also divides by len(target_traj)).
    Returns (ADE_b = mean(euc), FDE_b = ||stack(err)||_F / len(batch), euc[n], err[n,2]).
    """
    euc, err = [], []
    nt = len(targets) if n_targets is None else n_targets
    for i, tgt in enumerate(targets):
        L = len(tgt)
        if L < pred_len:                                     # :641-645
            e = np.linalg.norm(pred[i][:L] - tgt, ord=2) / nt / 12
            d = pred[i][L - 1] - tgt[L - 1]
        else:                                                # :646-650
            e = np.linalg.norm(pred[i][:pred_len] - tgt[:pred_len], ord=2) / 12
            d = pred[i][pred_len - 1] - tgt[pred_len - 1]
        euc.append(e)
        err.append(d)
    euc = np.asarray(euc)
    err = np.asarray(err)
    nb = len(targets) if n_batch_frames is None else n_batch_frames
    return float(np.mean(euc)), float(np.linalg.norm(err) / nb), euc, err   # :668-674


def get_mean_error(predicted_traj, true_traj, observed_length, maxNumPeds):
    """sample.get_mean_error (sample.py:21-82), value-for-value.

    Inputs [n, L, 2] (agent-major); transposed to time-major (:33-34).  The signed error
    2-vectors are summed over agents before the norm (defect F-8); only steps obs..L-1 scored;
    ``counter`` accumulates over all scored steps.  Returns (ade, fde, counter).
    """
    true_traj = np.transpose(true_traj, (1, 0, 2))
    predicted_traj = np.transpose(predicted_traj, (1, 0, 2))
    L = len(true_traj)
    error = np.zeros((L - observed_length, 2))
    fde_error = []
    counter = 0
    fde_counter = 0
    for i in range(observed_length, L):
        ts = np.zeros(2)
        fde_counter = 0
        for j in range(maxNumPeds):
            ts = ts + (true_traj[i, j] - predicted_traj[i, j])          # :64
            if i == L - 1:
                fde_error.append(true_traj[i, j] - predicted_traj[i, j])  # :66-69
                fde_counter += 1
            counter += 1
        if counter != 0:
            error[i - observed_length] = ts                              # :73-74
    ade = np.mean(np.linalg.norm(error, ord=2, axis=1) / counter)        # :82
    fde = np.mean(np.linalg.norm(np.stack(fde_error), ord=2, axis=1) / fde_counter)
    return ade, fde, counter


def standard_best_of_k(samples, gt, valid=None):
    """Standard definition (SURVEY A.6 last bullet).  samples[..., K, P, 2], gt[..., P, 2].

    ADE_k = mean_t ||y_kt - y_t||, FDE_k = ||y_kP - y_P||, k* = argmin_k ADE_k (ties -> lowest k).
    fp32, fixed sequential order over t so the GPU kernel can be bit-identical.
    """
    f32 = np.float32
    s = samples.astype(f32)
    g = gt.astype(f32)[..., None, :, :]
    ex = s[..., 0] - g[..., 0]
    ey = s[..., 1] - g[..., 1]
    d = np.sqrt(ex * ex + ey * ey, dtype=f32)            # [..., K, P]
    P = d.shape[-1]
    acc = np.zeros(d.shape[:-1], f32)
    for t in range(P):
        acc = acc + d[..., t]
    ade = acc / f32(P)
    fde = d[..., P - 1]
    best = np.argmin(ade, axis=-1).astype(np.int32)     # first minimum = lowest k
    return ade, fde, best


def ade_fde_world(pred, gt, H, valid=None, scale=(480.0, 640.0)):
    """ADE / FDE in metres: data/eth/univ/getPixelCoordinates.m:8-30 run backwards (the files hold
    pinv(H) * world, rows divided by 480 and 640): world = H [p0*480, p1*640, 1]^T / third component.
    pred/gt [n,P,2] -> (ade[n], fde[n]) in fp64; invalid agents score 0."""
    pred, gt, H = np.asarray(pred, np.float64), np.asarray(gt, np.float64), np.asarray(H, np.float64)

    def world(q):
        p = np.stack([q[..., 0] * scale[0], q[..., 1] * scale[1], np.ones(q.shape[:-1])], -1)
        w = p @ H.T
        return w[..., :2] / w[..., 2:3]
    d = np.linalg.norm(world(pred) - world(gt), axis=-1)
    ade, fde = d.mean(-1), d[:, -1]
    if valid is not None:
        v = np.asarray(valid).astype(bool)
        ade, fde = np.where(v, ade, 0.0), np.where(v, fde, 0.0)
    return ade, fde
