"""CPU oracle: ``tf.contrib.rnn.GridLSTMCell`` exactly as the reference instantiates it
(TEST INFRASTRUCTURE ONLY -- never imported by the product path).

TensorFlow 1.14 (un-vendored third-party dependency; version string recovered from
``save/*.meta``) is not installable here, so the cell's dataflow is restated from the op graph
under scope ``grid_lstm_cell/`` of ``save/g2k_mcr_model_val_0.ckpt-0.meta`` and
its published algorithm (Kalchbrenner et al., Grid LSTM; ``contrib/rnn/python/ops/rnn_cell.py``
``GridLSTMCell._compute``).  Call sites in the reference: ``helper.py:31-39,68`` (peepholes on,
``num_frequency_blocks=[hidden_len/grid_size]``) and ``helper.py:131-141`` (peepholes off,
``[grid_size/2]`` blocks).  Both use ``num_units=U``, ``feature_size=frequency_skip=4``,
``share_time_frequency_weights=True``, ``couple_input_forget_gates=True``,
``state_is_tuple=False``.

Parity status: "parity unpinned" -- the reference holds no TF-evaluated input/output pair for
the cell (its outputs are *fed*, ``train.py:204-207``); the parameters ``W_f_0_0[8,6]``,
``B_f_0[6]`` and the four peephole diagonals are taken from the shipped checkpoints as fixtures.
"""
from __future__ import annotations

import numpy as np


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gridlstm_step(inputs, state, W_f, B_f, w_If, w_It, w_Of, w_Ot, U, F, feat=4, peepholes=True):
    """One GridLSTMCell call.  inputs[B, >=feat*F], state[B, >=2*U*F] (flat c_time/m_time per block).

    Returns (m_out[B, 2*U*F], state_out[B, 2*U*F]).  Per block f:
        z = concat(x_f, m_t, m_f) @ W_f + B_f ; i, j, o = split(z, 3)
        g = sigmoid(i + w_If*c_f + w_It*c_t)
        c_freq = (1-g)*c_f + g*tanh(j) ; c_time = (1-g)*c_t + g*tanh(j)
        q = sigmoid(o + w_Of*c_freq + w_Ot*c_time)
        m_freq = q*tanh(c_freq) ; m_time = q*tanh(c_time)
    with (m_f, c_f) = 0 for f == 0, else block f-1's (m_freq, c_freq).
    state_out = concat_f [c_time, m_time]; m_out = concat_f [m_time, m_freq].
    """
    B = inputs.shape[0]
    dt = inputs.dtype
    m_f = np.zeros((B, U), dt)
    c_f = np.zeros((B, U), dt)
    st_out, m_out = [], []
    for f in range(F):
        x_f = inputs[:, feat * f: feat * f + feat]
        c_t = state[:, 2 * f * U: 2 * f * U + U]
        m_t = state[:, (2 * f + 1) * U: (2 * f + 2) * U]
        z = np.concatenate([x_f, m_t, m_f], axis=1) @ W_f + B_f
        i, j, o = z[:, :U], z[:, U:2 * U], z[:, 2 * U:3 * U]
        if peepholes:
            g = sigmoid(i + w_If * c_f + w_It * c_t)
        else:
            g = sigmoid(i)
        tj = np.tanh(j)
        c_freq = (1 - g) * c_f + g * tj
        c_time = (1 - g) * c_t + g * tj
        if peepholes:
            q = sigmoid(o + w_Of * c_freq + w_Ot * c_time)
        else:
            q = sigmoid(o)
        m_freq = q * np.tanh(c_freq)
        m_time = q * np.tanh(c_time)
        st_out += [c_time, m_time]
        m_out += [m_time, m_freq]
        m_f, c_f = m_freq, c_freq
    return np.concatenate(m_out, axis=1), np.concatenate(st_out, axis=1)
