"""Mirror of ``models/gsk_lstm_cell.py`` (reference :4-65).

The reference class is a dead Hadamard stub that cannot build at pred_len 12.  The
constructor signature is kept; the object is the *fused cell* the north_star names: the
GridLSTMCell gate equations of helper.py:31-39 at U = 128 over the graph neighbourhood, one
CUDA kernel per step (``mmt_gsk_cell``; fp32 CUDA-core parity mode or bf16 tcgen05 mode)."""
from __future__ import annotations

import torch

from .. import ops, synth
from ._weights import size0


class gsk_lstm_cell():
    def __init__(self, in_features, out_size, obs_len, num_nodes, lambda_reg, params: ops.CellParams = None,
                 precision="fp16", device="cuda"):
        self.out_size = int(num_nodes)
        self.hidden_size = int(out_size)
        self.obs_len = int(obs_len)
        self.lambda_reg = float(lambda_reg)
        self.in_size = size0(in_features)
        self.device = torch.device(device)
        self.prec = ops.prec_from_name(precision)
        self.params = params if params is not None else ops.CellParams.from_numpy(
            synth.init_params(seed=0, U=self.hidden_size), self.device)
        self.pred_path_band = None

    def init_state(self, rows):
        z = torch.zeros((rows, self.hidden_size), dtype=torch.float32, device=self.device)
        return z, z.clone()

    def __call__(self, x, h, c, mh, mc, valid, cur_pos=None):
        """One step on R = S*N rows: x[R,4] (dx,dy,vx,vy); returns (h', c', m_f[, params[R,5], next_pos])."""
        return ops.gsk_cell(x, h, c, mh, mc, valid, self.params, self.prec, cur_pos=cur_pos,
                            want_head=cur_pos is not None)

    forward = __call__
