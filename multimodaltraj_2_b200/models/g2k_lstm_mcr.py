"""Mirror of ``models/g2k_lstm_mcr.py`` (reference :3-124): same class name, constructor keywords and
attribute names; the attributes are CUDA tensors and ``forward()`` runs the hand-written kernels
(``mmt_mcr_forward_f32`` / ``mmt_mcr_step_f32``) instead of building a TF-1 graph.

Reference semantics kept: ``outputs`` [D+2,D], ``rel_features`` [2,D], ``ngh`` [D,T],
``hidden_states`` [D,H] are inputs ("placeholders": assign, then call ``forward()``);
``attn`` [D,D], ``cost`` [T,T], ``pred_path_band`` [2,P,n] are results.  ``forward_batched`` is the
added entry point that runs S scenes at once (one CTA per scene) including the per-frame state
step of train.py:240-254.
"""
from __future__ import annotations

import torch

from .. import ops
from ._weights import init_normal, lookup, size0


class g2k_lstm_mcr():
    variant = 0

    def __init__(self, in_features, hidden_size, obs_len, num_nodes, lambda_reg, sess_g=None, pred_len=12,
                 device="cuda"):
        D, T, n = size0(in_features), int(obs_len), int(num_nodes)
        self.device = torch.device(device)
        self.out_size = n
        self.lambda_reg = float(lambda_reg)
        self.pred_len = int(pred_len)
        self.hidden_size = int(hidden_size)
        # "placeholders" (defaults are seed-0 normals, as tf.placeholder_with_default in the reference)
        self.outputs = init_normal((D + 2, D), 0, self.device)
        self.rel_features = init_normal((2, D), 0, self.device)
        self.visual_path = init_normal((2, D), 0, self.device)
        self.ngh = init_normal((D, T), 0, self.device)
        self.hidden_states = init_normal((D, hidden_size), 0, self.device)
        # variables (reference :38-76): by-name lookup in sess_g, else seed-0 init
        self.cost = lookup(sess_g, "cost", (T, T), 0, self.device)
        self.attn = lookup(sess_g, "attn", (D, D), 0, self.device)
        self.weight_v = lookup(sess_g, "weight_v", (T, D + 2), 0, self.device)
        self.bias_v = lookup(sess_g, "bias_v", (D,), 0, self.device)
        self.weight_o = init_normal((T, n), 0, self.device)              # always fresh (reference :61-64)
        self.weight_c = lookup(sess_g, "weight_c", (2 * self.pred_len, T), 0, self.device)
        self.weight_r = lookup(sess_g, "weight_r", (T, 2), 0, self.device)
        self.pred_path_band = None
        self.forward()

    # ------------------------------------------------------------------------------------------
    def _w(self, extra=None):
        w = dict(W_v=self.weight_v, b_v=self.bias_v, W_r=self.weight_r, W_c=self.weight_c, W_o=self.weight_o)
        if extra:
            w.update(extra)
        return w

    def forward(self):
        """models/g2k_lstm_mcr.py:99-124 on the current placeholder values (one scene)."""
        o = ops.mcr_forward(self.outputs[None].contiguous(), self.rel_features[None].contiguous(),
                            self.ngh[None].contiguous(), self._w(), self.lambda_reg, self.out_size, self.pred_len,
                            self.variant)
        self.ngh_scaled = self.lambda_reg * self.ngh
        self.attn, self.cost = o["attn"][0], o["cost"][0]
        self.temp_path = o["band"][0].reshape(2 * self.pred_len, self.out_size)
        self.pred_path_band = o["band"][0]                                # [2, P, n]
        return self.pred_path_band

    def forward_batched(self, X, V, C, Hs, weight_i, weight_ii, vemb_prev=None):
        """S scenes at once: embeddings (train.py:178-183,194-195) + forward + per-frame state step
        (train.py:240-254).  X[S,T,n] V[S,2,n] C[S,D,D] Hs[S,D,H] -> dict(attn, cost, band, pred, Hs, adj, vemb)."""
        return ops.mcr_step(X, V, C, Hs, self._w(dict(W_i=weight_i, W_ii=weight_ii)), self.lambda_reg,
                            self.pred_len, self.variant, vemb_prev)

    # ------------------------------------------------------------------------------------------
    relational = True          # g2k_lstm_mcr: edge-MLP scores join the attention logits (nri_learned.py:5-28)

    def forecast_batched(self, pos, vis, valid, params, K=20, r2=4.0, inv_2sigma2=0.5, prec=ops.PREC_F16, seed=0,
                         agent_offset=0, eps=None, use_graph=False):
        """The north_star path behind the drop-in class: all scenes of a batch through pairwise kernel -> [edge MLP]
        -> aggregation -> gate update for obs_len + pred_len - 1 frames, then K-sample decode + ADE/FDE + best-of-K
        (``mmt_forecast_f32``).  pos[S,N,T+P,2], vis[S,N,T,2], valid[S,N]; ``params``: ops.CellParams.
        g2k_lstm_mc in the f16 (default) / bf16 modes with 128 % N == 0 runs the fused persistent rollout kernel; g2k_lstm_mcr the
        relational per-step tensor-core kernels.  The forecaster (workspace, CUDA graphs) is cached per shape."""
        S, N = valid.shape
        T = pos.shape[2] - self.pred_len
        key = (S, N, T, K, prec, seed, agent_offset, use_graph, id(params))
        cache = self.__dict__.setdefault("_forecasters", {})
        if key not in cache:
            cache[key] = ops.Forecaster(params, S, N, T, self.pred_len, K, r2, inv_2sigma2, relational=self.relational,
                                        prec=prec, seed=seed, agent_offset=agent_offset, device=pos.device,
                                        use_graph=use_graph)
        return cache[key](pos, vis, valid, eps=eps)

