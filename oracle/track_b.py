"""CPU oracle, Track B: the batched north_star kernels (TEST INFRASTRUCTURE ONLY).

Nothing in the product path may import this module.  PARITY UNPINNED: the reference only
gestures at these operations (no N x N kernel, no bivariate-Gaussian decode, no
K-sampling, ``gsk_lstm_cell`` and ``nri_learned`` are dead stubs), so their semantics are
*defined here* (SURVEY App. C), each anchored on the reference lines it generalises:

  pairwise / adjacency   networkx_graph.py:71,83-85 (dist_mat, L2 edge norm); argParser.py:56-60
  aggregation            train.py:240-247 (row-softmax attention times hidden states)
  edge MLP               relational_inf_models/nri_learned.py:5-28 (sigmoid gate, softmax);
                         fNRI-master.zip modules.py:17-49,94-152 (node2edge -> MLP(ELU) -> score)
  cell                   helper.py:31-39 + SURVEY App. B (GridLSTMCell gates), generalised to
                         U=128 with the frequency axis := graph neighbourhood
  decode + score         train.py:639-674, sample.py:21-82 (what is scored); standard best-of-K

All arithmetic is fp32.  The integer-producing parts (adjacency, degree, neighbour lists,
best-of-K) use a fixed operation order with no fused multiply-add so the CUDA kernels can be
bit-identical.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------------------------
# C.1 pairwise kernel + adjacency
# ----------------------------------------------------------------------------------------------
def pairwise_adj(pos, valid, r2, inv_2sigma2):
    """pos[S,N,2] f32, valid[S,N] u8 -> kern[S,N,N] f32, adj[S,N,N] u8, deg[S,N] i32.

    d2 = (dx*dx) + (dy*dy) (fp32, each op rounded, no FMA); adj = d2 < r2 and i != j and both
    valid; kern = adj * exp(-d2 * inv_2sigma2); deg = row-sum of adj.
    """
    pos = pos.astype(f32)
    x, y = pos[..., 0], pos[..., 1]
    dx = x[:, :, None] - x[:, None, :]
    dy = y[:, :, None] - y[:, None, :]
    d2 = (dx * dx).astype(f32) + (dy * dy).astype(f32)
    N = pos.shape[1]
    v = valid.astype(bool)
    adj = (d2 < f32(r2)) & ~np.eye(N, dtype=bool)[None] & v[:, :, None] & v[:, None, :]
    kern = np.where(adj, np.exp(-(d2 * f32(inv_2sigma2)), dtype=f32), f32(0)).astype(f32)
    return kern, adj.astype(np.uint8), adj.sum(-1).astype(np.int32)


def neighbor_index(adj, max_nbr):
    """adj[S,N,N] u8 -> nbr[S,N,max_nbr] i32 (ascending j, padded with -1), cnt[S,N] i32
    (number actually written = min(deg, max_nbr))."""
    S, N, _ = adj.shape
    nbr = np.full((S, N, max_nbr), -1, np.int32)
    cnt = np.zeros((S, N), np.int32)
    for s in range(S):
        for i in range(N):
            js = np.nonzero(adj[s, i])[0][:max_nbr]
            nbr[s, i, :len(js)] = js
            cnt[s, i] = len(js)
    return nbr, cnt


# ----------------------------------------------------------------------------------------------
# C.2 aggregation
# ----------------------------------------------------------------------------------------------
def masked_softmax(logits, adj):
    """Row softmax over j in adj[i,:]; rows with no neighbour -> all zeros."""
    m = adj.astype(bool)
    lg = np.where(m, logits, f32(-np.inf)).astype(f32)
    mx = np.max(lg, axis=-1, keepdims=True)
    mx = np.where(np.isfinite(mx), mx, f32(0))
    e = np.where(m, np.exp(lg - mx, dtype=f32), f32(0)).astype(f32)
    den = e.sum(-1, keepdims=True, dtype=f32)
    return np.where(den > 0, e / np.where(den > 0, den, f32(1)), f32(0)).astype(f32)


def aggregate(logits, adj, feat):
    """a = masked_softmax(logits, adj); out = a @ feat.  feat[S,N,C] -> (a[S,N,N], out[S,N,C])."""
    a = masked_softmax(logits, adj)
    return a, np.matmul(a, feat.astype(f32)).astype(f32)


# ----------------------------------------------------------------------------------------------
# C.3 relational edge MLP (mcr only)
# ----------------------------------------------------------------------------------------------
def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0), dtype=f32)).astype(f32)


def edge_mlp(h, adj, p):
    """score_ij = sigmoid(w_out . elu(W2^T elu(W1a^T h_i + W1b^T h_j + b1) + b2) + b_out) on
    edges of adj, 0 elsewhere.  h[S,N,U]; p: W1[2U,He] b1[He] W2[He,He] b2[He] w_out[He] b_out[]
    (nri_learned.infer_rlns sigmoid gate, nri_learned.py:16-21; fNRI node2edge + MLP)."""
    U = h.shape[-1]
    a = np.matmul(h, p["W1"][:U]).astype(f32)          # sender-independent half  [S,N,He]
    b = np.matmul(h, p["W1"][U:]).astype(f32)
    e1 = elu(a[:, :, None, :] + b[:, None, :, :] + p["b1"])            # [S,N,N,He]
    e2 = elu(np.matmul(e1, p["W2"]).astype(f32) + p["b2"])
    s = (np.matmul(e2, p["w_out"]) + p["b_out"]).astype(f32)
    sc = (f32(1) / (f32(1) + np.exp(-s, dtype=f32))).astype(f32)
    return np.where(adj.astype(bool), sc, f32(0)).astype(f32)


# ----------------------------------------------------------------------------------------------
# C.4 gsk cell (GridLSTM gates at U=128 over the graph neighbourhood)
# ----------------------------------------------------------------------------------------------
def sigmoid(x):
    return (f32(1) / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)


def gsk_cell(x, h, c, mh, mc, valid, p):
    """One fused cell step.  x[S,N,4] (dx,dy,vx,vy), h/c own state [S,N,U], mh/mc aggregated
    neighbour state [S,N,U].  p: W_e[4,E] b_e[E] W[E+2U,3U] b[3U] w_If w_It w_Of w_Ot [U].

        e = relu(x @ W_e + b_e);  z = [e | h | mh] @ W + b;  i, j, o = split(z)
        g = sigmoid(i + w_If*mc + w_It*c)
        c_f = (1-g)*mc + g*tanh(j);  c_t = (1-g)*c + g*tanh(j)
        q = sigmoid(o + w_Of*c_f + w_Ot*c_t)
        m_f = q*tanh(c_f);  m_t = q*tanh(c_t)
    Returns (h'=m_t, c'=c_t, m_f); rows of invalid agents are zero.
    """
    U = h.shape[-1]
    e = np.maximum(np.matmul(x.astype(f32), p["W_e"]) + p["b_e"], f32(0)).astype(f32)
    u = np.concatenate([e, h, mh], axis=-1).astype(f32)
    z = (np.matmul(u, p["W"]) + p["b"]).astype(f32)
    i, j, o = z[..., :U], z[..., U:2 * U], z[..., 2 * U:]
    g = sigmoid(i + p["w_If"] * mc + p["w_It"] * c)
    tj = np.tanh(j, dtype=f32)
    c_f = ((f32(1) - g) * mc + g * tj).astype(f32)
    c_t = ((f32(1) - g) * c + g * tj).astype(f32)
    q = sigmoid(o + p["w_Of"] * c_f + p["w_Ot"] * c_t)
    m_f = (q * np.tanh(c_f, dtype=f32)).astype(f32)
    m_t = (q * np.tanh(c_t, dtype=f32)).astype(f32)
    v = valid.astype(bool)[..., None]
    z0 = f32(0)
    return np.where(v, m_t, z0), np.where(v, c_t, z0), np.where(v, m_f, z0)


def head(m_t, m_f, p):
    """y = [m_t | m_f] @ W_h + b_h -> activated (mu_x, mu_y, sigma_x, sigma_y, rho):
    sigma = exp(.), rho = tanh(.)."""
    y = (np.matmul(np.concatenate([m_t, m_f], -1), p["W_h"]) + p["b_h"]).astype(f32)
    out = y.copy()
    out[..., 2] = np.exp(y[..., 2], dtype=f32)
    out[..., 3] = np.exp(y[..., 3], dtype=f32)
    out[..., 4] = np.tanh(y[..., 4], dtype=f32)
    return out.astype(f32)


# ----------------------------------------------------------------------------------------------
# C.5 decode + score
# ----------------------------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint32 arrays (Salmon et al. 2011); exact integer arithmetic."""
    c0, c1, c2, c3 = (np.asarray(a, np.uint32) for a in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c0.astype(np.uint64)
            p1 = _PHILOX_M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _PHILOX_W0)
            k1 = np.uint32(k1 + _PHILOX_W1)
    return c0, c1, c2, c3


def philox_eps(seed, S, N, K, P, agent_offset=0):
    """eps[S,N,K,P,2]: one Philox call serves two steps: counter = (global agent index, k, t//2, 0),
    key = (seed_lo, seed_hi); step t uses words (x0,x1) if t is even else (x2,x3);
    u0 = (xa + 1) * 2^-32 in (0,1], u1 = xb * 2^-32; Box-Muller
    eps1 = sqrt(-2 ln u0) cos(2 pi u1), eps2 = sqrt(-2 ln u0) sin(2 pi u1)."""
    a = (np.arange(S * N, dtype=np.uint64) + np.uint64(agent_offset)).astype(np.uint32).reshape(S, N, 1, 1)
    k = np.arange(K, dtype=np.uint32).reshape(1, 1, K, 1)
    t = np.arange(P, dtype=np.uint32).reshape(1, 1, 1, P)
    a, k, t = np.broadcast_arrays(a, k, t)
    x0, x1, x2, x3 = philox4x32_10(a, k, t >> np.uint32(1), np.zeros_like(a),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    odd = (t & np.uint32(1)).astype(bool)
    xa = np.where(odd, x2, x0)
    xb = np.where(odd, x3, x1)
    u0 = ((xa.astype(np.float64) + 1.0) * 2.0 ** -32).astype(f32)
    u1 = (xb.astype(np.float64) * 2.0 ** -32).astype(f32)
    r = np.sqrt(f32(-2) * np.log(u0, dtype=f32), dtype=f32)
    th = f32(2 * np.pi) * u1
    return np.stack([r * np.cos(th, dtype=f32), r * np.sin(th, dtype=f32)], -1).astype(f32)


def decode_score(params, eps, last_obs, gt, valid):
    """params[S,N,P,5] activated (mu_x,mu_y,sig_x,sig_y,rho); eps[S,N,K,P,2]; last_obs[S,N,2];
    gt[S,N,P,2]; valid[S,N].  Sampled displacement per step:
        dx = mu_x + sig_x*e1 ;  dy = mu_y + sig_y*(rho*e1 + sqrt(1-rho^2)*e2)
    positions accumulate from the last observed point.  Every op is a separately rounded fp32
    op in the order written (bit-exact contract with the CUDA kernel).
    Returns ade[S,N,K], fde[S,N,K], best_k[S,N] i32 (argmin ADE, ties -> lowest k; -1 for
    invalid agents), best_traj[S,N,P,2], samples[S,N,K,P,2].
    """
    p = params.astype(f32)
    S, N, P, _ = p.shape
    K = eps.shape[2]
    mux, muy, sx, sy, rho = (p[..., None, :, i] for i in range(5))      # [S,N,1,P]
    e1, e2 = eps[..., 0].astype(f32), eps[..., 1].astype(f32)          # [S,N,K,P]
    om = np.sqrt(f32(1) - rho * rho, dtype=f32)
    dx = mux + sx * e1
    dy = muy + sy * ((rho * e1).astype(f32) + (om * e2).astype(f32))
    px = np.broadcast_to(last_obs[..., 0].astype(f32)[..., None], (S, N, K)).copy()
    py = np.broadcast_to(last_obs[..., 1].astype(f32)[..., None], (S, N, K)).copy()
    samples = np.zeros((S, N, K, P, 2), f32)
    acc = np.zeros((S, N, K), f32)
    d = None
    for t in range(P):
        px = px + dx[..., t]
        py = py + dy[..., t]
        samples[..., t, 0] = px
        samples[..., t, 1] = py
        ex = px - gt[..., None, t, 0].astype(f32)
        ey = py - gt[..., None, t, 1].astype(f32)
        d = np.sqrt((ex * ex).astype(f32) + (ey * ey).astype(f32), dtype=f32)
        acc = acc + d
    ade = (acc / f32(P)).astype(f32)
    fde = d.astype(f32)
    best = np.argmin(ade, axis=-1).astype(np.int32)
    bt = np.take_along_axis(samples, best[:, :, None, None, None].astype(np.int64), axis=2)[:, :, 0]
    v = valid.astype(bool)
    ade = np.where(v[..., None], ade, f32(0))
    fde = np.where(v[..., None], fde, f32(0))
    best = np.where(v, best, -1).astype(np.int32)
    bt = np.where(v[..., None, None], bt, f32(0)).astype(f32)
    return ade, fde, best, bt, samples


# ----------------------------------------------------------------------------------------------
# whole path: obs 8 -> pred 12 rollout
# ----------------------------------------------------------------------------------------------
def rollout(pos, vis, valid, p, T=8, P=12, r2=4.0, inv_2sigma2=0.5, relational=False, trace=False):
    """pos[S,N,T+P,2] (observed + ground-truth future; only the first T frames are read here),
    vis[S,N,T,2], valid[S,N].  Runs T+P-1 cell steps; steps t >= T-1 emit the parameters of
    frame t+1 and the mean displacement feeds the next step's position.

    Returns dict(params[S,N,P,5] activated, pred_mean[S,N,P,2], h, c) and, with ``trace``,
    per-step adjacency / degree / positions.
    """
    S, N = valid.shape
    U = p["w_If"].shape[0]
    h = np.zeros((S, N, U), f32)
    c = np.zeros((S, N, U), f32)
    params = np.zeros((S, N, P, 5), f32)
    pred_mean = np.zeros((S, N, P, 2), f32)
    cur = pos[:, :, 0].astype(f32)
    prev = cur
    tr = dict(adj=[], deg=[], pos=[])
    for t in range(T + P - 1):
        if t < T:
            cur = pos[:, :, t].astype(f32)
            v_t = vis[:, :, t].astype(f32)
        else:
            v_t = vis[:, :, T - 1].astype(f32)
        disp = (cur - prev).astype(f32) if t > 0 else np.zeros_like(cur)
        x = np.concatenate([disp, v_t], -1).astype(f32)
        kern, adj, deg = pairwise_adj(cur, valid, r2, inv_2sigma2)
        logits = kern
        if relational:
            logits = (kern + edge_mlp(h, adj, p)).astype(f32)
        a = masked_softmax(logits, adj)
        mh = np.matmul(a, h).astype(f32)
        mc = np.matmul(a, c).astype(f32)
        h, c, m_f = gsk_cell(x, h, c, mh, mc, valid, p)
        if trace:
            tr["adj"].append(adj)
            tr["deg"].append(deg)
            tr["pos"].append(cur.copy())
        prev = cur
        if t >= T - 1:
            y = head(h, m_f, p)
            y = np.where(valid.astype(bool)[..., None], y, f32(0)).astype(f32)
            params[:, :, t - (T - 1)] = y
            cur = (cur + y[..., :2]).astype(f32)
            pred_mean[:, :, t - (T - 1)] = cur
    out = dict(params=params, pred_mean=pred_mean, h=h, c=c)
    if trace:
        out["trace"] = tr
    return out


def forecast(pos, vis, valid, p, eps, T=8, P=12, r2=4.0, inv_2sigma2=0.5, relational=False):
    """Whole unit of work: rollout + K-sample decode + ADE/FDE + best-of-K."""
    ro = rollout(pos, vis, valid, p, T, P, r2, inv_2sigma2, relational)
    ade, fde, best, bt, _ = decode_score(ro["params"], eps, pos[:, :, T - 1], pos[:, :, T:T + P], valid)
    return dict(ade=ade, fde=fde, best_k=best, best_traj=bt, **ro)


def forecast_scene_loop(pos, vis, valid, p, eps, **kw):
    """The reference's execution shape (train.py:71-254): one scene at a time."""
    outs = [forecast(pos[s:s + 1], vis[s:s + 1], valid[s:s + 1], p, eps[s:s + 1], **kw) for s in range(pos.shape[0])]
    return {k: np.concatenate([o[k] for o in outs]) for k in ("ade", "fde", "best_k", "best_traj", "params")}
