"""ORACLE (test infrastructure, not product code): teacher-forced training loss of the Track-B path and its
gradients by PyTorch autograd on the CPU in fp64.  PARITY UNPINNED: the reference has no loss / optimiser
(SURVEY F2); the definition below is the one SURVEY App. C.5 proposes (mean bivariate-Gaussian NLL over valid
agent-steps + (lambda/2) ||W||^2; RMSProp lr 0.005, decay 0.95, clip 10: the unused flags of argParser.py:40-47,72).

Only tests/ and bench.py's cpu_baseline may import this file.

Teacher forcing: every cell step t = 0 .. T+P-2 reads the ground-truth frame pos[:, :, t] (the reference feeds
ground-truth windows to its per-frame loop, train.py:71-87), so adjacency and attention are functions of the inputs
only; steps t >= T-1 emit (mu, sigma, rho) of the displacement pos[t+1] - pos[t].
"""
import math

import numpy as np
import torch

PARAM_KEYS = ("W_e", "b_e", "W", "b", "w_If", "w_It", "w_Of", "w_Ot", "W_h", "b_h")
EDGE_KEYS = ("W1", "b1", "W2", "b2", "w_out", "b_out")   # g2k_lstm_mcr: relational edge MLP (track_b.edge_mlp)


def edge_scores(h, p):
    """track_b.edge_mlp in torch: sigmoid(w_out . elu(W2^T elu(W1a^T h_i + W1b^T h_j + b1) + b2) + b_out), all pairs."""
    U = h.shape[-1]
    a, b = h @ p["W1"][:U], h @ p["W1"][U:]
    e1 = torch.nn.functional.elu(a[:, :, None, :] + b[:, None, :, :] + p["b1"])
    e2 = torch.nn.functional.elu(e1 @ p["W2"] + p["b2"])
    return torch.sigmoid(e2 @ p["w_out"] + p["b_out"])


def attention(cur, valid, r2, inv_2sigma2, score=None):
    """cur[S,N,2], valid[S,N] bool -> att[S,N,N] (masked softmax of the kernel [+ edge score] over the adjacency; rows
    without neighbours are zero).  Same arithmetic as track_b.pairwise_adj + masked_softmax."""
    d = cur[:, :, None, :] - cur[:, None, :, :]
    d2 = (d * d).sum(-1)
    N = cur.shape[1]
    adj = (d2 < r2) & ~torch.eye(N, dtype=torch.bool)[None] & valid[:, :, None] & valid[:, None, :]
    kern = torch.exp(-d2 * inv_2sigma2)
    if score is not None:
        kern = kern + score
    lg = torch.where(adj, kern, torch.full_like(kern, -math.inf))
    mx = lg.max(-1, keepdim=True).values
    mx = torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))
    e = torch.where(adj, torch.exp(lg - mx), torch.zeros_like(kern))
    den = e.sum(-1, keepdim=True)
    return torch.where(den > 0, e / torch.where(den > 0, den, torch.ones_like(den)), torch.zeros_like(e))


def cell(x, h, c, mh, mc, valid, p):
    U = h.shape[-1]
    e = torch.relu(x @ p["W_e"] + p["b_e"])
    z = torch.cat([e, h, mh], -1) @ p["W"] + p["b"]
    i, j, o = z[..., :U], z[..., U:2 * U], z[..., 2 * U:]
    g = torch.sigmoid(i + p["w_If"] * mc + p["w_It"] * c)
    tj = torch.tanh(j)
    c_f = (1 - g) * mc + g * tj
    c_t = (1 - g) * c + g * tj
    q = torch.sigmoid(o + p["w_Of"] * c_f + p["w_Ot"] * c_t)
    v = valid[..., None].to(h.dtype)
    return q * torch.tanh(c_t) * v, c_t * v, q * torch.tanh(c_f) * v


def nll(y, d):
    """y[...,5] raw head outputs, d[...,2] target displacement -> negative log likelihood of the bivariate Gaussian
    with mu = y[0:2], sigma = exp(y[2:4]), rho = tanh(y[4])."""
    sx, sy, rho = torch.exp(y[..., 2]), torch.exp(y[..., 3]), torch.tanh(y[..., 4])
    zx, zy = (d[..., 0] - y[..., 0]) / sx, (d[..., 1] - y[..., 1]) / sy
    om = 1 - rho * rho
    return math.log(2 * math.pi) + y[..., 2] + y[..., 3] + 0.5 * torch.log(om) + (zx * zx - 2 * rho * zx * zy + zy * zy) / (2 * om)


def loss_fn(pos, vis, valid, p, T=8, P=12, r2=4.0, inv_2sigma2=0.5, lam=0.0005, relational=False):
    """Scalar training loss for torch tensors (any dtype).  relational: g2k_lstm_mcr (logits = kern + edge score)."""
    S, N = valid.shape
    U = p["w_If"].shape[0]
    h = torch.zeros((S, N, U), dtype=pos.dtype)
    c = torch.zeros((S, N, U), dtype=pos.dtype)
    vb = valid.bool()
    total = torch.zeros((), dtype=pos.dtype)
    for t in range(T + P - 1):
        cur = pos[:, :, t]
        disp = cur - pos[:, :, t - 1] if t > 0 else torch.zeros_like(cur)
        x = torch.cat([disp, vis[:, :, min(t, T - 1)]], -1)
        att = attention(cur, vb, r2, inv_2sigma2, edge_scores(h, p) if relational else None)
        mh, mc = att @ h, att @ c
        h, c, m_f = cell(x, h, c, mh, mc, vb, p)
        if t >= T - 1:
            y = torch.cat([h, m_f], -1) @ p["W_h"] + p["b_h"]
            total = total + (nll(y, pos[:, :, t + 1] - cur) * vb.to(pos.dtype)).sum()
    n = vb.sum().to(pos.dtype) * P
    return total / n + 0.5 * lam * (p["W"] * p["W"]).sum()


def loss_and_grads(pos, vis, valid, p_np, T=8, P=12, r2=4.0, inv_2sigma2=0.5, lam=0.0005, relational=False):
    """numpy in, numpy out: (loss, {name: gradient}) in fp64."""
    keys = PARAM_KEYS + (EDGE_KEYS if relational else ())
    p = {k: torch.tensor(np.asarray(p_np[k], np.float64), requires_grad=True) for k in keys}
    loss = loss_fn(torch.tensor(pos, dtype=torch.float64), torch.tensor(vis, dtype=torch.float64),
                   torch.tensor(valid), p, T, P, r2, inv_2sigma2, lam, relational)
    loss.backward()
    return float(loss.detach()), {k: p[k].grad.numpy() for k in keys}


def rmsprop_step(p, g, ms, lr=0.005, decay=0.95, eps=1e-10, clip=10.0):
    """One RMSProp update with global-norm clipping (the reference's unused flags: argParser.py:40-47).
    p, g, ms: dicts of numpy arrays; returns (new p, new ms)."""
    gn = math.sqrt(sum(float((g[k].astype(np.float64) ** 2).sum()) for k in g))
    s = min(1.0, clip / max(gn, 1e-30))
    out_p, out_ms = {}, {}
    for k in g:
        gk = g[k] * s
        out_ms[k] = decay * ms[k] + (1 - decay) * gk * gk
        out_p[k] = p[k] - lr * gk / (np.sqrt(out_ms[k]) + eps)
    return out_p, out_ms
