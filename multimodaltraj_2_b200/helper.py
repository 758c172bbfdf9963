"""Mirror of ``helper.py`` (reference :10-141): the two neighbourhood encoders wrapping
``tf.contrib.rnn.GridLSTMCell``.  Same class names, constructor keywords and attribute names;
``forward()`` evaluates the cell with ``mmt_gridlstm_step_f32`` (SURVEY App. B dataflow)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _glstm_params(U, seed, device, sess_g=None):
    """W_f[4+2U,3U], B_f[3U] (zero-init as TF), peephole diagonals [U]; taken from ``sess_g`` (a decoded
    checkpoint dict with ``grid_lstm_cell/*`` keys) when shapes match."""
    rng = np.random.Generator(np.random.Philox(seed))
    fan = 4 + 2 * U

    def get(name, shape, init):
        if isinstance(sess_g, dict):
            for k, v in sess_g.items():
                if k.split("/")[-1].split(":")[0] == name and tuple(np.shape(v)) == tuple(shape):
                    return torch.as_tensor(np.asarray(v), dtype=torch.float32).contiguous().to(device)
        return torch.from_numpy(init.astype(np.float32)).to(device)

    lim = np.sqrt(6.0 / (fan + 3 * U))      # glorot-uniform, TF's default variable initializer
    return dict(W_f=get("W_f_0_0", (fan, 3 * U), rng.uniform(-lim, lim, (fan, 3 * U))),
                B_f=get("B_f_0", (3 * U,), np.zeros(3 * U)),
                w_If=get("W_I_diag_freqf_0", (U,), rng.uniform(-1, 1, U)),
                w_It=get("W_I_diag_freqt_0", (U,), rng.uniform(-1, 1, U)),
                w_Of=get("W_O_diag_freqf_0", (U,), rng.uniform(-1, 1, U)),
                w_Ot=get("W_O_diag_freqt_0", (U,), rng.uniform(-1, 1, U)))


class neighborhood_vis_loc_encoder():
    """helper.py:10-75.  GridLSTMCell(num_units=num_layers, feature_size=frequency_skip=grid_size,
    use_peepholes=True, num_frequency_blocks=[hidden_len/grid_size], shared weights, coupled gates)."""

    def __init__(self, hidden_size, hidden_len, num_layers, grid_size, embedding_size, dropout=0, sess_g=None,
                 device="cuda"):
        assert grid_size == 4, "the cell kernel is built for feature_size = 4 (the reference's grid_size)"
        self.hidden_size, self.embedding_size = hidden_size, embedding_size
        self.hidden_len, self.U, self.F = hidden_len, num_layers, int(hidden_len / grid_size)
        self.device = torch.device(device)
        self.input = torch.zeros((hidden_len, hidden_len), dtype=torch.float32, device=self.device)
        self.state_f00_b00_c = self.init_hidden(hidden_len)
        self.c_hidden_state = self.init_hidden(hidden_len)
        self.output = torch.zeros((hidden_len, hidden_len), dtype=torch.float32, device=self.device)
        self.w = _glstm_params(self.U, 0, self.device, sess_g)
        self.peepholes = True
        self.forward()

    def update_input_size(self, new_size):
        self.input = torch.zeros((new_size, new_size), dtype=torch.float32, device=self.device)
        self.hidden_state = self.init_hidden(new_size)

    def forward(self):
        """self.output, self.c_hidden_state = rnn(inputs=self.input, state=self.state_f00_b00_c)  (helper.py:68)"""
        w = self.w
        self.output, self.c_hidden_state = ops.gridlstm_step(
            self.input.contiguous(), self.state_f00_b00_c.contiguous(), w["W_f"], w["B_f"], w["w_If"], w["w_It"],
            w["w_Of"], w["w_Ot"], self.U, self.F, self.peepholes)
        return self.output, self.c_hidden_state

    def init_hidden(self, size):
        return torch.zeros((size, self.hidden_size), dtype=torch.float32, device=self.device)


class neighborhood_stat_enc():
    """helper.py:77-141.  Second GridLSTMCell, peepholes off, [grid_size/2] frequency blocks, reuse=True
    (shares W_f / B_f with the first encoder: pass its ``w`` as ``shared``)."""

    def __init__(self, ctxt_path, hidden_size, num_layers, grid_size, dim, shared=None, device="cuda"):
        self.hidden_size, self.U, self.F = hidden_size, num_layers, int(grid_size / 2)
        self.device = torch.device(device)
        self.ctxt_path = ctxt_path
        self.input = torch.zeros((dim, 8), dtype=torch.float32, device=self.device)
        self.hidden_state = torch.zeros((dim, hidden_size), dtype=torch.float32, device=self.device)
        self.w = shared if shared is not None else _glstm_params(self.U, 0, self.device)
        self.forward()

    def forward(self):
        """self.output, self.c_hidden_states = rnn(self.input, self.hidden_state)  (helper.py:141)"""
        w = self.w
        self.output, self.c_hidden_states = ops.gridlstm_step(
            self.input.contiguous(), self.hidden_state.contiguous(), w["W_f"], w["B_f"], w["w_If"], w["w_It"],
            w["w_Of"], w["w_Ot"], self.U, self.F, False)
        return self.output, self.c_hidden_states


def static_context(ctxt_img, dim, obs_len, lambda_param, filt=None, seed=0, device="cuda"):
    """Static-context branch of train.py:93-110,154-158 on the device (``mmt_static_context_f32``): the scene image
    (``imread('ctxt.png')``, [H,W,3]) correlated with one [H+3-dim, W+2-dim, 3] filter -> ``_2dconv[dim,dim]`` scaled by
    ``lambda_param``, and ``_2dconv_in = _2dconv x stat_mask`` -> the ``ngh[dim, obs_len]`` fed to the model.
    The reference draws the filter with an unseeded ``tf.random_normal``; here it is an argument, or drawn from numpy's
    Philox generator with ``seed`` so that runs are reproducible.  Returns (_2dconv, _2dconv_in) as CUDA tensors."""
    import numpy as np
    img = torch.as_tensor(np.asarray(ctxt_img, np.float32)).to(device).contiguous()
    H, W, C = img.shape
    if filt is None:
        g = np.random.Generator(np.random.Philox(seed))
        filt = g.standard_normal((H + 3 - dim, W + 2 - dim, C)).astype(np.float32)
    filt = torch.as_tensor(np.asarray(filt, np.float32)).to(device).contiguous()
    return ops.static_context(img, filt, int(dim), int(obs_len), float(lambda_param))
