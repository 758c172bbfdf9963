"""Mirror of ``models/g2k_lstm_mc.py`` (reference :4-69): the multi-cue model.  As written, its
``forward`` routes through ``tf.gradients(ys=<placeholder>, ..., unconnected_gradients='zero')``
(:59-61) so ``cost`` and ``pred_path_band`` are identically zero; the kernel variant 1
reproduces exactly that, while ``Eo`` and the per-frame state step are still computed."""
from __future__ import annotations

from ._weights import init_normal, size0
from .g2k_lstm_mcr import g2k_lstm_mcr


class g2k_lstm_mc(g2k_lstm_mcr):
    variant = 1
    relational = False         # no edge MLP: ``forecast_batched`` reaches the fused persistent rollout kernel in bf16 mode

    def __init__(self, in_features, out_size, obs_len, num_nodes, lambda_reg, pred_len=12, device="cuda"):
        # reference signature: (in_features, out_size, obs_len, num_nodes, lambda_reg); out_size is the
        # hidden width there too (train.py passes rnn_size)
        super().__init__(in_features=in_features, hidden_size=out_size, obs_len=obs_len, num_nodes=num_nodes,
                         lambda_reg=lambda_reg, sess_g=None, pred_len=pred_len, device=device)
        self.visual_path = init_normal((1, size0(in_features)), 0, self.device)     # [1,D] in the mc model (:18-20)
