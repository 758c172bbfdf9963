"""Mirror of ``relational_inf_models/nri_learned.py`` (reference :5-28).

``infer_rlns`` (sigmoid of the kernel matrix, :16-21) and ``eval_rln_ngh`` (softmax, :23-28) keep
their signatures.  ``graph_to_kernel`` is a non-runnable stub in the reference (three lines pasted
from fNRI referencing undefined names, :5-14); here it is the relational edge MLP those lines
gesture at (fNRI node2edge -> 2-layer ELU MLP -> per-edge score, one edge type), evaluated on the
edges of the adjacency mask by ``mmt_edge_mlp_f32``."""
from __future__ import annotations

from .. import ops


def graph_to_kernel(h, adj, params: ops.CellParams):
    """h[S,N,U] hidden states, adj[S,N,N] u8 -> relational scores[S,N,N] (0 off-graph)."""
    p = params
    return ops.edge_mlp(h, adj, p.W1, p.b1, p.W2, p.b2, p.w_out, p.b_out)


def infer_rlns(adj_mat):
    """prob_mat = sigmoid(adj_mat)  (nri_learned.py:16-21)"""
    return ops.sigmoid(adj_mat.contiguous())


def eval_rln_ngh(adj_mat, combined_ngh=None):
    """prob_mat = softmax(adj_mat) over the last axis  (nri_learned.py:23-28); combined_ngh is unused there too."""
    return ops.rowsoftmax(adj_mat.contiguous())
