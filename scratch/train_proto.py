"""CPU prototype of the manual BPTT that the CUDA training step implements, checked against autograd (oracle/train_b.py)."""
import sys, math, numpy as np, torch
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import train_b as tb
from multimodaltraj_2_b200 import synth
torch.set_default_dtype(torch.float64)
S, N, T, P, U, E = 3, 16, 8, 12, 128, 64
pos, vis, valid = synth.make_crowd(S, N, seed=5, half_extent=3.0, ragged=True)
p_np = synth.init_params(seed=1)
loss_ref, g_ref = tb.loss_and_grads(pos, vis, valid, p_np)
p = {k: torch.tensor(np.asarray(p_np[k], np.float64)) for k in tb.PARAM_KEYS}
pos_t, vis_t, vb = torch.tensor(pos, dtype=torch.float64), torch.tensor(vis, dtype=torch.float64), torch.tensor(valid).bool()
v = vb[..., None].double()
nsteps = T + P - 1
# ---- forward, saving
h = torch.zeros((S, N, U)); c = torch.zeros((S, N, U)); sv = []
loss = 0.0
nvalid = float(vb.sum()) * P
for t in range(nsteps):
    cur = pos_t[:, :, t]
    disp = cur - pos_t[:, :, t - 1] if t > 0 else torch.zeros_like(cur)
    x = torch.cat([disp, vis_t[:, :, min(t, T - 1)]], -1)
    att = tb.attention(cur, vb, 4.0, 0.5)
    mh, mc = att @ h, att @ c
    hn, cn, mf = tb.cell(x, h, c, mh, mc, vb, p)
    rec = dict(x=x, h=h, c=c, att=att, mh=mh, mc=mc, hn=hn, mf=mf)
    if t >= T - 1:
        y = torch.cat([hn, mf], -1) @ p["W_h"] + p["b_h"]
        rec["y"] = y; rec["d"] = pos_t[:, :, t + 1] - cur
        loss += float((tb.nll(y, rec["d"]) * vb.double()).sum())
    sv.append(rec); h, c = hn, cn
loss = loss / nvalid + 0.5 * 0.0005 * float((p["W"] ** 2).sum())
# ---- backward
g = {k: torch.zeros_like(p[k]) for k in p}
Gh = torch.zeros((S, N, U)); Gc = torch.zeros((S, N, U))
sig = torch.sigmoid
for t in reversed(range(nsteps)):
    r = sv[t]
    dmt, dmf = Gh.clone(), torch.zeros((S, N, U))
    if "y" in r:
        y, d = r["y"], r["d"]
        sx, sy, rho = torch.exp(y[..., 2]), torch.exp(y[..., 3]), torch.tanh(y[..., 4])
        zx, zy = (d[..., 0] - y[..., 0]) / sx, (d[..., 1] - y[..., 1]) / sy
        om = 1 - rho * rho; Q = zx * zx - 2 * rho * zx * zy + zy * zy
        dy = torch.stack([-(zx - rho * zy) / (om * sx), -(zy - rho * zx) / (om * sy), 1 - (zx * zx - rho * zx * zy) / om,
                          1 - (zy * zy - rho * zx * zy) / om, -rho + (-zx * zy * om + rho * Q) / om], -1) * (vb.double() / nvalid)[..., None]
        hm = torch.cat([r["hn"], r["mf"]], -1)
        g["W_h"] += hm.reshape(-1, 2 * U).T @ dy.reshape(-1, 5); g["b_h"] += dy.reshape(-1, 5).sum(0)
        dhm = dy @ p["W_h"].T
        dmt = dmt + dhm[..., :U]; dmf = dhm[..., U:]
    x, hp, cp, mh, mc = r["x"], r["h"], r["c"], r["mh"], r["mc"]
    e = torch.relu(x @ p["W_e"] + p["b_e"])
    A = torch.cat([e, hp, mh], -1)
    z = A @ p["W"] + p["b"]
    i, j, o = z[..., :U], z[..., U:2 * U], z[..., 2 * U:]
    gg = sig(i + p["w_If"] * mc + p["w_It"] * cp); tj = torch.tanh(j)
    cf = (1 - gg) * mc + gg * tj; ct = (1 - gg) * cp + gg * tj
    q = sig(o + p["w_Of"] * cf + p["w_Ot"] * ct); tcf, tct = torch.tanh(cf), torch.tanh(ct)
    dmt = dmt * v; dmf = dmf * v; Gcv = Gc * v
    dq = dmt * tct + dmf * tcf
    dct = dmt * q * (1 - tct * tct) + Gcv
    dcf = dmf * q * (1 - tcf * tcf)
    dpo = dq * q * (1 - q)
    dcf = dcf + dpo * p["w_Of"]; dct = dct + dpo * p["w_Ot"]
    g["w_Of"] += (dpo * cf).sum((0, 1)); g["w_Ot"] += (dpo * ct).sum((0, 1))
    dg = dcf * (tj - mc) + dct * (tj - cp)
    dmc = dcf * (1 - gg); dcp = dct * (1 - gg); dtj = (dcf + dct) * gg
    dj = dtj * (1 - tj * tj)
    dpi = dg * gg * (1 - gg)
    dmc = dmc + dpi * p["w_If"]; dcp = dcp + dpi * p["w_It"]
    g["w_If"] += (dpi * mc).sum((0, 1)); g["w_It"] += (dpi * cp).sum((0, 1))
    dz = torch.cat([dpi, dj, dpo], -1)
    g["W"] += A.reshape(-1, E + 2 * U).T @ dz.reshape(-1, 3 * U); g["b"] += dz.reshape(-1, 3 * U).sum(0)
    dA = dz @ p["W"].T
    de, dhp, dmh = dA[..., :E], dA[..., E:E + U], dA[..., E + U:]
    dpre = de * (e > 0)
    g["W_e"] += x.reshape(-1, 4).T @ dpre.reshape(-1, E); g["b_e"] += dpre.reshape(-1, E).sum(0)
    attT = r["att"].transpose(1, 2)
    Gh = dhp + attT @ dmh
    Gc = dcp + attT @ dmc
g["W"] += 0.0005 * p["W"]
print("loss", loss, loss_ref)
for k in tb.PARAM_KEYS:
    a, b = g[k].numpy(), g_ref[k]
    print(f"{k:6s} rel err {np.abs(a - b).max() / max(np.abs(b).max(), 1e-30):.2e}  |g| {np.abs(b).max():.3e}")
