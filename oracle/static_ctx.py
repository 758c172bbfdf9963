"""CPU oracle of the static-context branch (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Restates train.py:93-110 (scene image, padded, correlated with ONE filter of shape [width-dim+1, height-dim+1, 3, 1]
with padding='VALID' -> _2dconv[dim,dim], scaled by lambda_param) and train.py:154-158 (stat_mask rows =
tf.range(0, 1, 1/obs_len); _2dconv_in = _2dconv @ stat_mask) in fp64.  The reference draws the filter with an unseeded
tf.random_normal and its ctxt.png is not in the repository, so neither can be a golden vector: PARITY UNPINNED beyond
the algebra; the tests pin it with impulse images (the response to a single 1 is the flipped filter) and linearity."""
from __future__ import annotations

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view


def static_context(img, filt, D, T, lam):
    """img[H,W,C], filt[H+3-D, W+2-D, C] -> (_2dconv[D,D], _2dconv_in[D,T])."""
    img, filt = np.asarray(img, np.float64), np.asarray(filt, np.float64)
    imgp = np.pad(img, ((1, 1), (0, 1), (0, 0)))                 # train.py:97-99
    FH, FW = imgp.shape[0] - D + 1, imgp.shape[1] - D + 1        # train.py:100-105 (names width / height there)
    assert filt.shape == (FH, FW, img.shape[2]), (filt.shape, (FH, FW, img.shape[2]))
    win = sliding_window_view(imgp, (FH, FW, img.shape[2]))[:, :, 0]      # [D, D, FH, FW, C]
    conv = lam * np.einsum("ijabc,abc->ij", win, filt)           # train.py:103-109
    mask = np.zeros((D, T)) + np.arange(T)[None, :] * (1.0 / T)  # train.py:154-155 (tf.range(0, 1, 1/obs_len))
    return conv, conv @ mask                                     # train.py:157


def seeded_filter(H, W, C, D, seed=0):
    """The filter the host mirror draws: N(0,1) from numpy's Philox generator (the reference's draw is unseeded)."""
    g = np.random.Generator(np.random.Philox(seed))
    return g.standard_normal((H + 3 - D, W + 2 - D, C)).astype(np.float32)
