"""Torch-tensor wrappers over the C-ABI (``include/mmt.h``).  PyTorch only supplies device memory
and the current stream; every operation is a hand-written CUDA kernel in ``csrc/``.

All functions require CUDA tensors and raise if the extension cannot be loaded -- there is no
eager/CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch

from . import _lib

PREC_F32, PREC_BF16, PREC_BF16_STEPWISE, PREC_BF16X3, PREC_F16 = 0, 1, 2, 3, 4


def prec_from_name(name):
    """--precision of the drivers (argParser.py) -> MMT_PREC_*: 'fp16' (default: fp16 tensor-core operands, inside the 1e-3
    ADE / FDE bar), 'bf16', 'bf16x3' (split bf16, fp32-grade), 'fp32' (CUDA-core parity mode)."""
    table = {"fp16": PREC_F16, "f16": PREC_F16, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "fp32": PREC_F32, "f32": PREC_F32}
    if name not in table:
        raise ValueError(f"unknown precision {name!r}: one of {sorted(table)}")
    return table[name]


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{name}: expected a CUDA tensor (libmmt has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


_graph_launches = 0   # kernel launches replayed through CUDA graphs (not seen by the library's own counter)


def launch_count() -> int:
    """Kernels launched by this process through libmmt (direct launches + CUDA-graph replays)."""
    return int(_lib.load().mmt_launch_count()) + _graph_launches


# --------------------------------------------------------------------------------------------
def pairwise_adj(pos, valid, r2, inv_2sigma2, want_kern=True, want_adj=True, want_deg=True):
    """pos[S,N,2] f32, valid[S,N] u8 -> (kern[S,N,N] f32, adj[S,N,N] u8, deg[S,N] i32)."""
    lib = _lib.load()
    _chk(pos, torch.float32, "pos"); _chk(valid, torch.uint8, "valid")
    S, N, _ = pos.shape
    kern = torch.empty((S, N, N), dtype=torch.float32, device=pos.device) if want_kern else None
    adj = torch.empty((S, N, N), dtype=torch.uint8, device=pos.device) if want_adj else None
    deg = torch.empty((S, N), dtype=torch.int32, device=pos.device) if want_deg else None
    _lib.check(lib.mmt_pairwise_adj_f32(_p(pos), _p(valid), S, N, r2, inv_2sigma2, _p(kern), _p(adj), _p(deg),
                                        _stream()), "mmt_pairwise_adj_f32")
    return kern, adj, deg


def neighbor_index(adj, max_nbr):
    lib = _lib.load()
    _chk(adj, torch.uint8, "adj")
    S, N, _ = adj.shape
    nbr = torch.empty((S, N, max_nbr), dtype=torch.int32, device=adj.device)
    cnt = torch.empty((S, N), dtype=torch.int32, device=adj.device)
    _lib.check(lib.mmt_neighbor_index_i32(_p(adj), S, N, max_nbr, _p(nbr), _p(cnt), _stream()),
               "mmt_neighbor_index_i32")
    return nbr, cnt


def aggregate(logits, adj, feat, want_attn=True):
    lib = _lib.load()
    _chk(logits, torch.float32, "logits"); _chk(adj, torch.uint8, "adj"); _chk(feat, torch.float32, "feat")
    S, N, Cc = feat.shape
    attn = torch.empty((S, N, N), dtype=torch.float32, device=feat.device) if want_attn else None
    out = torch.empty_like(feat)
    _lib.check(lib.mmt_aggregate_f32(_p(logits), _p(adj), _p(feat), S, N, Cc, _p(attn), _p(out), _stream()),
               "mmt_aggregate_f32")
    return attn, out


def aggregate_transpose(attn, d):
    """out[S,N,C] = attn^T @ d per scene (mmt_aggregate_transpose_f32): the adjoint of ``aggregate``."""
    lib = _lib.load()
    _chk(attn, torch.float32, "attn"); _chk(d, torch.float32, "d")
    S, N, Cc = d.shape
    out = torch.empty_like(d)
    _lib.check(lib.mmt_aggregate_transpose_f32(_p(attn), _p(d), S, N, Cc, _p(out), _stream()), "mmt_aggregate_transpose_f32")
    return out


def pack_edge_weights(W1, W2):
    """bf16 tensor-core operand images of the edge-MLP weights (mmt_pack_edge_weights_bf16; U = He = 128)."""
    lib = _lib.load()
    U, He = W1.shape[0] // 2, W2.shape[0]
    nb = lib.mmt_edge_weights_packed_bytes(U, He)
    if nb == 0:
        raise ValueError("tensor-core edge MLP needs U = He = 128")
    packed = torch.empty((nb,), dtype=torch.uint8, device=W1.device)
    _lib.check(lib.mmt_pack_edge_weights_bf16(_p(W1), _p(W2), U, He, _p(packed), _stream()), "mmt_pack_edge_weights_bf16")
    return packed


def edge_mlp(h, adj, W1, b1, W2, b2, w_out, b_out, prec=PREC_F32, packed=None):
    """score[S,N,N] of the relational edge MLP on the edges of adj (0 elsewhere).  prec = PREC_BF16 runs the
    tcgen05 version (U = He = 128; ``packed``: the images of pack_edge_weights, built here when None)."""
    lib = _lib.load()
    for n, t in dict(h=h, W1=W1, b1=b1, W2=W2, b2=b2, w_out=w_out, b_out=b_out).items():
        _chk(t, torch.float32, n)
    _chk(adj, torch.uint8, "adj")
    S, N, U = h.shape
    He = W2.shape[0]
    score = torch.empty((S, N, N), dtype=torch.float32, device=h.device)
    work = torch.empty((2 * S * N * He,), dtype=torch.float32, device=h.device)
    if prec != PREC_F32:
        if packed is None:
            packed = pack_edge_weights(W1, W2)
        _lib.check(lib.mmt_edge_mlp_bf16(_p(h), _p(adj), _p(packed), _p(b1), _p(b2), _p(w_out), _p(b_out), S, N, U, He,
                                         _p(score), _p(work), work.numel() * 4, _stream()), "mmt_edge_mlp_bf16")
        return score
    _lib.check(lib.mmt_edge_mlp_f32(_p(h), _p(adj), _p(W1), _p(b1), _p(W2), _p(b2), _p(w_out), _p(b_out), S, N, U, He,
                                    _p(score), _p(work), work.numel() * 4, _stream()), "mmt_edge_mlp_f32")
    return score


def attention_score_grad(attn, adj, dm, v):
    """d loss / d attention logits [S,N,N] through the aggregated state (mmt_attention_score_grad_f32):
    attn_ij (dm_i . v_j - sum_k attn_ik dm_i . v_k) on the edges.  dm, v: [S,N,C]."""
    lib = _lib.load()
    _chk(attn, torch.float32, "attn"); _chk(adj, torch.uint8, "adj"); _chk(dm, torch.float32, "dm"); _chk(v, torch.float32, "v")
    S, N, Cc = v.shape
    out = torch.empty((S, N, N), dtype=torch.float32, device=v.device)
    _lib.check(lib.mmt_attention_score_grad_f32(_p(attn), _p(adj), _p(dm), _p(v), S, N, Cc, _p(out), _stream()),
               "mmt_attention_score_grad_f32")
    return out


def edge_mlp_backward(h, adj, dlogit, p, g, prec=PREC_F32, packed=None):
    """Backward of ``edge_mlp`` on the edges of adj (mmt_edge_mlp_backward_f32 / _bf16), no host synchronisation.
    p: CellParams (edge weights); g: dict of gradient tensors, W2 / b1 / b2 / w_out / b_out accumulated in place.
    Returns dab[S*N, 2 He] = [d loss / d a | d loss / d b] of the node projections a = h W1[:U], b = h W1[U:]."""
    lib = _lib.load()
    _chk(h, torch.float32, "h"); _chk(adj, torch.uint8, "adj"); _chk(dlogit, torch.float32, "dlogit")
    S, N, U = h.shape
    He = p.W2.shape[0]
    for k in ("W2", "b1", "b2", "w_out", "b_out"):
        _chk(g[k], torch.float32, "g." + k)
    dab = torch.empty((S * N, 2 * He), dtype=torch.float32, device=h.device)
    work = torch.empty((2 * S * N * He,), dtype=torch.float32, device=h.device)
    if prec != PREC_F32:
        if packed is None:
            packed = pack_edge_weights(p.W1, p.W2)
        _lib.check(lib.mmt_edge_mlp_backward_bf16(_p(h), _p(adj), _p(dlogit), _p(packed), _p(p.b1), _p(p.b2), _p(p.w_out),
                                                  _p(p.b_out), S, N, U, He, _p(dab), _p(g["W2"]), _p(g["b2"]), _p(g["w_out"]),
                                                  _p(g["b_out"]), _p(work), work.numel() * 4, _stream()),
                   "mmt_edge_mlp_backward_bf16")
        g["b1"] += dab[:, :He].sum(0)
        return dab
    _lib.check(lib.mmt_edge_mlp_backward_f32(_p(h), _p(adj), _p(dlogit), _p(p.W1), _p(p.b1), _p(p.W2), _p(p.b2), _p(p.w_out),
                                             _p(p.b_out), S, N, U, He, _p(dab), _p(g["W2"]), _p(g["b1"]), _p(g["b2"]),
                                             _p(g["w_out"]), _p(g["b_out"]), _p(work), work.numel() * 4, _stream()),
               "mmt_edge_mlp_backward_f32")
    return dab


# --------------------------------------------------------------------------------------------
@dataclass
class CellParams:
    """Device-resident weights of the fused gsk cell (+ head, + relational edge MLP)."""
    W_e: torch.Tensor; b_e: torch.Tensor; W: torch.Tensor; b: torch.Tensor
    w_If: torch.Tensor; w_It: torch.Tensor; w_Of: torch.Tensor; w_Ot: torch.Tensor
    W_h: torch.Tensor = None; b_h: torch.Tensor = None
    W1: torch.Tensor = None; b1: torch.Tensor = None; W2: torch.Tensor = None; b2: torch.Tensor = None
    w_out: torch.Tensor = None; b_out: torch.Tensor = None
    W_packed: torch.Tensor = field(default=None, repr=False)
    W_packed_x3: torch.Tensor = field(default=None, repr=False)
    W_packed_f16: torch.Tensor = field(default=None, repr=False)

    @property
    def E(self):
        return self.W_e.shape[1]

    @property
    def U(self):
        return self.w_If.shape[0]

    def pack(self):
        """Build the bf16 tcgen05 operand image of W once (mmt_pack_gate_weights_bf16)."""
        lib = _lib.load()
        nbytes = lib.mmt_gate_weights_packed_bytes(self.E, self.U)
        if nbytes == 0:
            raise RuntimeError("bf16 packing is built for E=64, U=128")
        if self.W_packed is None:     # re-packed IN PLACE afterwards: forecasters (and their CUDA graphs) hold its address
            self.W_packed = torch.empty((nbytes,), dtype=torch.uint8, device=self.W.device)
        _lib.check(lib.mmt_pack_gate_weights_bf16(_p(self.W), self.E, self.U, _p(self.W_packed), _stream()),
                   "mmt_pack_gate_weights_bf16")
        return self

    def pack_x3(self):
        """Split-bf16 operand image [W_hi ; W_hi ; W_lo] of PREC_BF16X3 (mmt_pack_gate_weights_bf16x3)."""
        lib = _lib.load()
        nbytes = lib.mmt_gate_weights_packed_x3_bytes(self.E, self.U)
        if nbytes == 0:
            raise RuntimeError("bf16x3 packing is built for E=64, U=128")
        if self.W_packed_x3 is None:
            self.W_packed_x3 = torch.empty((nbytes,), dtype=torch.uint8, device=self.W.device)
        _lib.check(lib.mmt_pack_gate_weights_bf16x3(_p(self.W), self.E, self.U, _p(self.W_packed_x3), _stream()),
                   "mmt_pack_gate_weights_bf16x3")
        return self

    def pack_f16(self):
        """fp16 operand image of W for PREC_F16 (mmt_pack_gate_weights_f16; same layout and size as the bf16 one)."""
        lib = _lib.load()
        nbytes = lib.mmt_gate_weights_packed_bytes(self.E, self.U)
        if nbytes == 0:
            raise RuntimeError("fp16 packing is built for E=64, U=128")
        if self.W_packed_f16 is None:
            self.W_packed_f16 = torch.empty((nbytes,), dtype=torch.uint8, device=self.W.device)
        _lib.check(lib.mmt_pack_gate_weights_f16(_p(self.W), self.E, self.U, _p(self.W_packed_f16), _stream()),
                   "mmt_pack_gate_weights_f16")
        return self

    def repack(self):
        """After the weights changed (a training step): refresh the operand images that exist, in place."""
        if self.W_packed is not None:
            self.pack()
        if self.W_packed_x3 is not None:
            self.pack_x3()
        if self.W_packed_f16 is not None:
            self.pack_f16()
        return self

    def c_cell(self):
        w = _lib.CellWeights()
        for n in ("W_e", "b_e", "W", "b", "w_If", "w_It", "w_Of", "w_Ot", "W_h", "b_h"):
            t = getattr(self, n)
            setattr(w, n, None if t is None else t.data_ptr())
        w.W_packed_bf16 = None if self.W_packed is None else self.W_packed.data_ptr()
        w.W_packed_bf16x3 = None if self.W_packed_x3 is None else self.W_packed_x3.data_ptr()
        w.W_packed_f16 = None if self.W_packed_f16 is None else self.W_packed_f16.data_ptr()
        w.E, w.U = self.E, self.U
        return w

    def c_edge(self):
        if self.W1 is None:
            return None
        w = _lib.EdgeWeights()
        for n in ("W1", "b1", "W2", "b2", "w_out", "b_out"):
            setattr(w, n, getattr(self, n).data_ptr())
        w.He = self.W2.shape[0]
        return w

    @staticmethod
    def from_numpy(p: dict, device="cuda"):
        kw = {k: torch.as_tensor(v, dtype=torch.float32).contiguous().to(device) for k, v in p.items()
              if k in CellParams.__dataclass_fields__}
        if "b_out" in kw:
            kw["b_out"] = kw["b_out"].reshape(1)
        return CellParams(**kw)


def gsk_cell(x, h, c, mh, mc, valid, params: CellParams, prec=PREC_F32, cur_pos=None, want_head=False):
    """One fused cell step on R rows.  Returns (h', c', m_f[, params[R,5], next_pos[R,2]])."""
    lib = _lib.load()
    for n, t in dict(x=x, h=h, c=c, mh=mh, mc=mc).items():
        _chk(t, torch.float32, n)
    _chk(valid, torch.uint8, "valid")
    R = x.shape[0]
    h_out, c_out, mf = torch.empty_like(h), torch.empty_like(c), torch.empty_like(h)
    par = nxt = None
    if want_head:
        par = torch.empty((R, 5), dtype=torch.float32, device=x.device)
        nxt = torch.empty((R, 2), dtype=torch.float32, device=x.device)
        _chk(cur_pos, torch.float32, "cur_pos")
    if prec == PREC_BF16 and params.W_packed is None:
        params.pack()
    if prec == PREC_BF16X3 and params.W_packed_x3 is None:
        params.pack_x3()
    if prec == PREC_F16 and params.W_packed_f16 is None:
        params.pack_f16()
    w = params.c_cell()
    _lib.check(lib.mmt_gsk_cell(_p(x), _p(h), _p(c), _p(mh), _p(mc), _p(valid), C.byref(w), R, prec, _p(h_out),
                                _p(c_out), _p(mf), _p(cur_pos), _p(par), 5, _p(nxt), _stream()), "mmt_gsk_cell")
    return (h_out, c_out, mf, par, nxt) if want_head else (h_out, c_out, mf)


def head_nll(m_t, m_f, valid, W_h, b_h, target, scale, loss_sum):
    """Raw head + bivariate-Gaussian NLL of target[R,2]; accumulates scale * sum(nll) into loss_sum[1] and returns
    dy[R,5] = scale * d nll / d y (mmt_head_nll_f32)."""
    lib = _lib.load()
    for n, t in dict(m_t=m_t, m_f=m_f, W_h=W_h, b_h=b_h, target=target, loss_sum=loss_sum).items():
        _chk(t, torch.float32, n)
    _chk(valid, torch.uint8, "valid")
    R, U = m_t.shape
    dy = torch.empty((R, 5), dtype=torch.float32, device=m_t.device)
    _lib.check(lib.mmt_head_nll_f32(_p(m_t), _p(m_f), _p(valid), _p(W_h), _p(b_h), _p(target), R, U, float(scale),
                                    _p(loss_sum), _p(dy), _stream()), "mmt_head_nll_f32")
    return dy


def gsk_gates(z, c, mc, valid, params: "CellParams"):
    """Forward gate update from pre-activations z[R,3U] (mmt_gsk_gates_f32) -> (h', c', m_f)."""
    lib = _lib.load()
    for n, t in dict(z=z, c=c, mc=mc).items():
        _chk(t, torch.float32, n)
    _chk(valid, torch.uint8, "valid")
    R, U = c.shape
    h_out, c_out, mf = torch.empty_like(c), torch.empty_like(c), torch.empty_like(c)
    _lib.check(lib.mmt_gsk_gates_f32(_p(z), _p(c), _p(mc), _p(valid), _p(params.w_If), _p(params.w_It), _p(params.w_Of),
                                     _p(params.w_Ot), R, U, _p(h_out), _p(c_out), _p(mf), _stream()), "mmt_gsk_gates_f32")
    return h_out, c_out, mf


# ---- packed training path (Trainer(gemm="tc")): hc = [h | c], mhc = [mh | mc], A = [e | h | mh], z = A W without bias
def train_frame_inputs(pos, vis, t, want_target):
    """Teacher-forced inputs of frame t (mmt_train_frame_inputs_f32): cur[S,N,2], x[R,4], target[R,2] or None."""
    lib = _lib.load()
    _chk(pos, torch.float32, "pos"); _chk(vis, torch.float32, "vis")
    S, N, F, _ = pos.shape
    T = vis.shape[2]
    R = S * N
    cur = torch.empty((S, N, 2), dtype=torch.float32, device=pos.device)
    x = torch.empty((R, 4), dtype=torch.float32, device=pos.device)
    target = torch.empty((R, 2), dtype=torch.float32, device=pos.device) if want_target else None
    _lib.check(lib.mmt_train_frame_inputs_f32(_p(pos), _p(vis), R, F, T, int(t), _p(cur), _p(x), _p(target), _stream()),
               "mmt_train_frame_inputs_f32")
    return cur, x, target


def train_gate_input(x, hc, mhc, params: "CellParams"):
    """A[R,E+2U] = [relu(x W_e + b_e) | h | mh] from the packed rows (mmt_train_gate_input_f32)."""
    lib = _lib.load()
    for n, t in dict(x=x, hc=hc, mhc=mhc).items():
        _chk(t, torch.float32, n)
    R, U, E = x.shape[0], params.U, params.E
    A = torch.empty((R, E + 2 * U), dtype=torch.float32, device=x.device)
    _lib.check(lib.mmt_train_gate_input_f32(_p(x), _p(hc), _p(mhc), _p(params.W_e), _p(params.b_e), R, E, U, _p(A), _stream()),
               "mmt_train_gate_input_f32")
    return A


def gsk_gates_packed(z, hc, mhc, valid, params: "CellParams"):
    """Gate update from z (no bias: params.b is added in the kernel) on packed rows -> (hc' [R,2U], h' [R,U], m_f [R,U])."""
    lib = _lib.load()
    for n, t in dict(z=z, hc=hc, mhc=mhc).items():
        _chk(t, torch.float32, n)
    _chk(valid, torch.uint8, "valid")
    R, U = z.shape[0], params.U
    hc_out = torch.empty((R, 2 * U), dtype=torch.float32, device=z.device)
    h_out = torch.empty((R, U), dtype=torch.float32, device=z.device)
    mf = torch.empty((R, U), dtype=torch.float32, device=z.device)
    _lib.check(lib.mmt_gsk_gates_packed_f32(_p(z), _p(params.b), _p(hc), _p(mhc), _p(valid), _p(params.w_If), _p(params.w_It),
                                            _p(params.w_Of), _p(params.w_Ot), R, U, _p(hc_out), _p(h_out), _p(mf), _stream()),
               "mmt_gsk_gates_packed_f32")
    return hc_out, h_out, mf


def gsk_cell_backward_packed(z, hc, mhc, valid, params: "CellParams", d_mt, d_head, d_ct, dpeep, db):
    """mmt_gsk_cell_backward_packed_f32 -> (dz[R,3U], dc[R,U], dmhc[R,2U] with its d mc half written)."""
    lib = _lib.load()
    for n, t in dict(z=z, hc=hc, mhc=mhc, d_mt=d_mt, dpeep=dpeep, db=db).items():
        _chk(t, torch.float32, n)
    R, U = z.shape[0], params.U
    dz = torch.empty_like(z)
    dc = torch.empty((R, U), dtype=torch.float32, device=z.device)
    dmhc = torch.empty((R, 2 * U), dtype=torch.float32, device=z.device)
    _lib.check(lib.mmt_gsk_cell_backward_packed_f32(_p(z), _p(params.b), _p(hc), _p(mhc), _p(valid), _p(params.w_If),
                                                    _p(params.w_It), _p(params.w_Of), _p(params.w_Ot), _p(d_mt), _p(d_head),
                                                    _p(d_ct), R, U, _p(dz), _p(dc), _p(dmhc), _p(dpeep), _p(db), _stream()),
               "mmt_gsk_cell_backward_packed_f32")
    return dz, dc, dmhc


def train_backward_split(dA, A, dmhc, gbe, params: "CellParams"):
    """dpre[R,E] = dA[:, :E] * (e > 0) (column sums accumulated into gbe); dmhc[:, :U] = dA[:, E+U:] (in place)."""
    lib = _lib.load()
    for n, t in dict(dA=dA, A=A, dmhc=dmhc, gbe=gbe).items():
        _chk(t, torch.float32, n)
    R, U, E = dA.shape[0], params.U, params.E
    dpre = torch.empty((R, E), dtype=torch.float32, device=dA.device)
    _lib.check(lib.mmt_train_backward_split_f32(_p(dA), _p(A), R, E, U, _p(dpre), _p(dmhc), _p(gbe), _stream()),
               "mmt_train_backward_split_f32")
    return dpre


def train_backward_merge(dA, back, dc, params: "CellParams"):
    """(Gh, Gc) = (dA[:, E:E+U] + back[:, :U], dc + back[:, U:]) (mmt_train_backward_merge_f32)."""
    lib = _lib.load()
    for n, t in dict(dA=dA, back=back, dc=dc).items():
        _chk(t, torch.float32, n)
    R, U, E = dA.shape[0], params.U, params.E
    Gh, Gc = torch.empty_like(dc), torch.empty_like(dc)
    _lib.check(lib.mmt_train_backward_merge_f32(_p(dA), _p(back), _p(dc), R, E, U, _p(Gh), _p(Gc), _stream()),
               "mmt_train_backward_merge_f32")
    return Gh, Gc


def gsk_cell_backward(z, c, mc, valid, params: "CellParams", d_mt, d_mf, d_ct, dpeep, db=None):
    """Backward of the gate update from the saved pre-activations (mmt_gsk_cell_backward_f32):
    returns (dz[R,3U], dc[R,U], dmc[R,U]); dpeep[4,U] and (if given) db[3U], the gate-bias gradient, are accumulated in place."""
    lib = _lib.load()
    for n, t in dict(z=z, c=c, mc=mc, d_mt=d_mt, dpeep=dpeep).items():
        _chk(t, torch.float32, n)
    _chk(valid, torch.uint8, "valid")
    R, U = c.shape
    dz, dc, dmc = torch.empty_like(z), torch.empty_like(c), torch.empty_like(c)
    _lib.check(lib.mmt_gsk_cell_backward_f32(_p(z), _p(c), _p(mc), _p(valid), _p(params.w_If), _p(params.w_It),
                                             _p(params.w_Of), _p(params.w_Ot), _p(d_mt), _p(d_mf), _p(d_ct), R, U,
                                             _p(dz), _p(dc), _p(dmc), _p(dpeep), _p(db), _stream()), "mmt_gsk_cell_backward_f32")
    return dz, dc, dmc


def gridlstm_step(inputs, state, W_f, B_f, w_If, w_It, w_Of, w_Ot, U, F, peepholes=True):
    """GridLSTMCell as helper.py instantiates it.  inputs[B,>=4F], state[B,>=2UF] -> (m_out, state_out)."""
    lib = _lib.load()
    _chk(inputs, torch.float32, "inputs"); _chk(state, torch.float32, "state")
    B = inputs.shape[0]
    m_out = torch.empty((B, 2 * U * F), dtype=torch.float32, device=inputs.device)
    st_out = torch.empty((B, 2 * U * F), dtype=torch.float32, device=inputs.device)
    _lib.check(lib.mmt_gridlstm_step_f32(_p(inputs), inputs.shape[1], _p(state), state.shape[1], _p(W_f), _p(B_f),
                                         _p(w_If), _p(w_It), _p(w_Of), _p(w_Ot), B, U, F, int(peepholes), _p(m_out),
                                         _p(st_out), _stream()), "mmt_gridlstm_step_f32")
    return m_out, st_out


def mcr_step(X, V, Cc, Hs, w: dict, lam, P=12, variant=0, vemb_prev=None):
    """Track-A batched scene-frame step.  X[S,T,n] V[S,2,n] C[S,D,D] Hs[S,D,H]; w: dict of device tensors
    W_i[n,D] W_ii[D,T] W_v[T,D+2] b_v[D] W_r[T,2] W_c[2P,T] W_o[T,n].  Returns dict."""
    lib = _lib.load()
    for n, t in dict(X=X, V=V, C=Cc, Hs=Hs).items():
        _chk(t, torch.float32, n)
    S, T, n = X.shape
    D, H = Hs.shape[1], Hs.shape[2]
    dev = X.device
    out = dict(attn=torch.empty((S, D, D), dtype=torch.float32, device=dev),
               cost=torch.empty((S, T, T), dtype=torch.float32, device=dev),
               band=torch.empty((S, 2, P, n), dtype=torch.float32, device=dev),
               Hs=torch.empty((S, D, H), dtype=torch.float32, device=dev),
               adj=torch.empty((S, D), dtype=torch.float32, device=dev),
               vemb=torch.empty((S, 2, D), dtype=torch.float32, device=dev))
    cw = _lib.McrWeights()
    for k in ("W_i", "W_ii", "W_v", "b_v", "W_r", "W_c", "W_o"):
        setattr(cw, k, _chk(w[k], torch.float32, k).data_ptr())
    _lib.check(lib.mmt_mcr_step_f32(_p(X), _p(V), _p(Cc), _p(Hs), _p(vemb_prev), C.byref(cw), S, n, D, T, P, H,
                                    float(lam), variant, _p(out["attn"]), _p(out["cost"]), _p(out["band"]),
                                    _p(out["Hs"]), _p(out["adj"]), _p(out["vemb"]), _stream()), "mmt_mcr_step_f32")
    out["pred"] = out["band"].permute(0, 3, 2, 1)          # train.py:254  [S,n,P,2] (view)
    return out


def decode_score(params, last_obs, gt, valid, K, eps=None, seed=0, agent_offset=0, want_all=True, want_traj=True,
                 dump_eps=False):
    """params[S,N,P,5] activated; returns dict(ade, fde [S,N,K], best_k, best_ade, best_fde, best_traj)."""
    lib = _lib.load()
    _chk(params, torch.float32, "params"); _chk(last_obs, torch.float32, "last_obs")
    _chk(gt, torch.float32, "gt"); _chk(valid, torch.uint8, "valid")
    S, N, P, _ = params.shape
    dev = params.device
    o = dict(best_k=torch.empty((S, N), dtype=torch.int32, device=dev),
             best_ade=torch.empty((S, N), dtype=torch.float32, device=dev))
    if dump_eps:
        o["eps"] = torch.empty((S, N, K, P, 2), dtype=torch.float32, device=dev)
        _lib.check(lib.mmt_decode_score_dump_eps_f32(_p(params), seed, agent_offset, _p(last_obs), _p(gt), _p(valid),
                                                     S, N, P, K, _p(o["best_k"]), _p(o["best_ade"]), _p(o["eps"]),
                                                     _stream()), "mmt_decode_score_dump_eps_f32")
        return o
    o["best_fde"] = torch.empty((S, N), dtype=torch.float32, device=dev)
    o["ade"] = torch.empty((S, N, K), dtype=torch.float32, device=dev) if want_all else None
    o["fde"] = torch.empty((S, N, K), dtype=torch.float32, device=dev) if want_all else None
    o["best_traj"] = torch.empty((S, N, P, 2), dtype=torch.float32, device=dev) if want_traj else None
    if eps is not None:
        _chk(eps, torch.float32, "eps")
    _lib.check(lib.mmt_decode_score_f32(_p(params), _p(eps), seed, agent_offset, _p(last_obs), _p(gt), _p(valid), S,
                                        N, P, K, _p(o["ade"]), _p(o["fde"]), _p(o["best_k"]), _p(o["best_ade"]),
                                        _p(o["best_fde"]), _p(o["best_traj"]), _stream()), "mmt_decode_score_f32")
    return o


def scene_batch(frame_ids, frame_row_start, ped_id, xy, vis, win_start, N, F, fstride):
    """Device-side padded scene batching -> (pos[S,N,F,2], vis[S,N,F,2]|None, valid[S,N], ped_of_slot[S,N])."""
    lib = _lib.load()
    for n, t in dict(frame_ids=frame_ids, frame_row_start=frame_row_start, ped_id=ped_id, win_start=win_start).items():
        _chk(t, torch.int32, n)
    _chk(xy, torch.float32, "xy")
    S = win_start.shape[0]
    dev = xy.device
    pos = torch.empty((S, N, F, 2), dtype=torch.float32, device=dev)
    vo = torch.empty((S, N, F, 2), dtype=torch.float32, device=dev) if vis is not None else None
    valid = torch.empty((S, N), dtype=torch.uint8, device=dev)
    slot = torch.empty((S, N), dtype=torch.int32, device=dev)
    _lib.check(lib.mmt_scene_batch_f32(_p(frame_ids), _p(frame_row_start), frame_ids.shape[0], _p(ped_id), _p(xy),
                                       _p(vis), _p(win_start), S, N, F, fstride, _p(pos), _p(vo), _p(valid), _p(slot),
                                       _stream()), "mmt_scene_batch_f32")
    return pos, vo, valid, slot


def rollout_bf16(pos, vis, valid, params: CellParams, T=8, P=12, r2=4.0, inv_2sigma2=0.5, out=None, timeline=None,
                 f16=False):
    """The T+P-1 step recurrence as one persistent tcgen05 kernel with the state on chip (mmt_rollout_bf16; f16=True:
    mmt_rollout_f16, the same kernel with fp16 operands): pos[S,N,T+P,2], vis[S,N,T,2], valid[S,N] -> params[S,N,P,5].
    N in {8,16,32,64,128}."""
    lib = _lib.load()
    _chk(pos, torch.float32, "pos"); _chk(vis, torch.float32, "vis"); _chk(valid, torch.uint8, "valid")
    S, N = valid.shape
    if f16 and params.W_packed_f16 is None:
        params.pack_f16()
    if not f16 and params.W_packed is None:
        params.pack()
    if out is None:
        out = torch.empty((S, N, P, 5), dtype=torch.float32, device=pos.device)
    cw = params.c_cell()
    fn, name = (lib.mmt_rollout_f16, "mmt_rollout_f16") if f16 else (lib.mmt_rollout_bf16, "mmt_rollout_bf16")
    _lib.check(fn(_p(pos), _p(vis), _p(valid), C.byref(cw), S, N, T, P, r2, inv_2sigma2, _p(out), _p(timeline), _stream()), name)
    return out


def rollout_f16(pos, vis, valid, params: CellParams, **kw):
    return rollout_bf16(pos, vis, valid, params, f16=True, **kw)


# --------------------------------------------------------------------------------------------
class Forecaster:
    """The whole hot path behind one call (mmt_forecast_f32): workspace and outputs are allocated
    once and reused, so a step is pure kernel launches on the current stream."""

    def __init__(self, params: CellParams, S, N, T=8, P=12, K=20, r2=4.0, inv_2sigma2=0.5, relational=False,
                 prec=PREC_F32, seed=0, agent_offset=0, device="cuda", want_all=False, use_graph=False):
        self.lib = _lib.load()
        self.use_graph = use_graph
        self._graphs = {}      # input-pointer tuple -> (CUDAGraph, launches per replay); at most 8 entries
        self._graph_n = 0
        self.p = params
        if prec in (PREC_BF16, PREC_BF16_STEPWISE) and params.W_packed is None:
            params.pack()
        if prec == PREC_BF16X3 and params.W_packed_x3 is None:
            params.pack_x3()
        if prec == PREC_F16 and params.W_packed_f16 is None:
            params.pack_f16()
        self.cfg = _lib.ForecastCfg(S, N, T, P, K, r2, inv_2sigma2, int(relational), prec, seed, agent_offset)
        He = params.W2.shape[0] if (relational and params.W2 is not None) else 0
        nbytes = self.lib.mmt_forecast_workspace_bytes(C.byref(self.cfg), params.U, He)
        self.work = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        self.out = dict(params=torch.empty((S, N, P, 5), dtype=torch.float32, device=device),
                        best_k=torch.empty((S, N), dtype=torch.int32, device=device),
                        best_ade=torch.empty((S, N), dtype=torch.float32, device=device),
                        best_fde=torch.empty((S, N), dtype=torch.float32, device=device),
                        best_traj=torch.empty((S, N, P, 2), dtype=torch.float32, device=device))
        self.out["ade"] = torch.empty((S, N, K), dtype=torch.float32, device=device) if want_all else None
        self.out["fde"] = torch.empty((S, N, K), dtype=torch.float32, device=device) if want_all else None
        self._cw = params.c_cell()
        self._ew = params.c_edge()

    def __call__(self, pos, vis, valid, eps=None):
        """Run the path.  With ``use_graph`` the ~60 launches of a rollout are captured once into a CUDA
        graph (keyed on the input pointers) and replayed, removing the per-launch host overhead."""
        if not self.use_graph:
            return self._launch(pos, vis, valid, eps)
        global _graph_launches
        key = (pos.data_ptr(), vis.data_ptr(), valid.data_ptr(), None if eps is None else eps.data_ptr())
        if key not in self._graphs:
            self._launch(pos, vis, valid, eps)          # eager warm-up: sets kernel attributes, validates arguments
            torch.cuda.synchronize()
            n0 = int(self.lib.mmt_launch_count())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch(pos, vis, valid, eps)
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = (g, int(self.lib.mmt_launch_count()) - n0)
        g, n = self._graphs[key]
        g.replay()
        _graph_launches += n
        return self.out

    def _launch(self, pos, vis, valid, eps=None):
        _chk(pos, torch.float32, "pos"); _chk(vis, torch.float32, "vis"); _chk(valid, torch.uint8, "valid")
        c = self.cfg
        want = {"pos": (pos, (c.S, c.N, c.T + c.P, 2)), "vis": (vis, (c.S, c.N, c.T, 2)), "valid": (valid, (c.S, c.N))}
        if eps is not None:
            _chk(eps, torch.float32, "eps")
            want["eps"] = (eps, (c.S, c.N, c.K, c.P, 2))
        for name, (t, shape) in want.items():        # the kernels index by these strides: a [S,N,F,2] vislet tensor
            if tuple(t.shape) != shape:               # (scene_batch's) passed unsliced would be read with the wrong one
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
        o = self.out
        ew = C.byref(self._ew) if self._ew is not None else None
        _lib.check(self.lib.mmt_forecast_f32(_p(pos), _p(vis), _p(valid), C.byref(self._cw), ew, C.byref(self.cfg),
                                             _p(eps), _p(o["params"]), _p(o["ade"]), _p(o["fde"]), _p(o["best_k"]),
                                             _p(o["best_ade"]), _p(o["best_fde"]), _p(o["best_traj"]), _p(self.work),
                                             self.work.numel(), _stream()), "mmt_forecast_f32")
        return o


# --------------------------------------------------------------------------------------------
# reference-compatible pieces (Track A forward alone, scores, relational helpers)
def mcr_forward(outputs, rel, ngh, w: dict, lam, n, P=12, variant=0):
    """g2k_lstm_mcr.forward alone on caller-built placeholders.  outputs[S,D+2,D] rel[S,2,D] ngh[S,D,T]."""
    lib = _lib.load()
    for nm, t in dict(outputs=outputs, rel=rel, ngh=ngh).items():
        _chk(t, torch.float32, nm)
    S, D = outputs.shape[0], outputs.shape[2]
    T = ngh.shape[2]
    dev = outputs.device
    out = dict(attn=torch.empty((S, D, D), dtype=torch.float32, device=dev),
               cost=torch.empty((S, T, T), dtype=torch.float32, device=dev),
               band=torch.empty((S, 2, P, n), dtype=torch.float32, device=dev))
    cw = _lib.McrWeights()
    for k in ("W_v", "b_v", "W_r", "W_c", "W_o"):
        setattr(cw, k, _chk(w[k], torch.float32, k).data_ptr())
    _lib.check(lib.mmt_mcr_forward_f32(_p(outputs), _p(rel), _p(ngh), C.byref(cw), S, n, D, T, P, float(lam), variant,
                                       _p(out["attn"]), _p(out["cost"]), _p(out["band"]), _stream()),
               "mmt_mcr_forward_f32")
    return out


def mean_error(predicted, truth, observed_length, maxNumPeds):
    """sample.get_mean_error on device: predicted/truth [n,L,2] -> tensor (ade, fde, counter)."""
    lib = _lib.load()
    _chk(predicted, torch.float32, "predicted"); _chk(truth, torch.float32, "truth")
    n, L, _ = predicted.shape
    out = torch.empty((3,), dtype=torch.float32, device=predicted.device)
    _lib.check(lib.mmt_mean_error_f32(_p(predicted), _p(truth), n, L, observed_length, maxNumPeds, _p(out), _stream()),
               "mmt_mean_error_f32")
    return out


def train_val_scores(pred, tgt, lens, n_targets):
    """train.py:639-674 per-agent scores: pred/tgt [n,P,2], lens[n] i32 -> (euc[n], err[n,2])."""
    lib = _lib.load()
    _chk(pred, torch.float32, "pred"); _chk(tgt, torch.float32, "tgt"); _chk(lens, torch.int32, "lens")
    n, P, _ = pred.shape
    euc = torch.empty((n,), dtype=torch.float32, device=pred.device)
    err = torch.empty((n, 2), dtype=torch.float32, device=pred.device)
    _lib.check(lib.mmt_train_val_scores_f32(_p(pred), _p(tgt), _p(lens), n, P, n_targets, _p(euc), _p(err), _stream()),
               "mmt_train_val_scores_f32")
    return euc, err


def ade_fde_world(pred, gt, H, valid=None, scale=(480.0, 640.0)):
    """ADE / FDE in metres (mmt_ade_fde_world_f32; data/eth/univ/getPixelCoordinates.m:8-30): pred/gt [n,P,2] in the
    data files' normalised pixel units, H[3,3] pixel -> world.  Returns (ade[n], fde[n], sums[3] = (sum ade, sum fde,
    number of valid agents))."""
    lib = _lib.load()
    _chk(pred, torch.float32, "pred"); _chk(gt, torch.float32, "gt"); _chk(H, torch.float32, "H")
    if valid is not None:
        _chk(valid, torch.uint8, "valid")
    n, P, _ = pred.shape
    ade = torch.empty((n,), dtype=torch.float32, device=pred.device)
    fde = torch.empty((n,), dtype=torch.float32, device=pred.device)
    sums = torch.empty((3,), dtype=torch.float32, device=pred.device)
    _lib.check(lib.mmt_ade_fde_world_f32(_p(pred), _p(gt), _p(valid), n, P, _p(H), float(scale[0]), float(scale[1]),
                                         _p(ade), _p(fde), _p(sums), _stream()), "mmt_ade_fde_world_f32")
    return ade, fde, sums


def static_context(img, filt, D, T, lam):
    """Static-context branch (mmt_static_context_f32; train.py:93-110,154-158): img[H,W,C], filt[H+3-D, W+2-D, C]
    -> (_2dconv[D,D], _2dconv_in = ngh[D,T])."""
    lib = _lib.load()
    _chk(img, torch.float32, "img"); _chk(filt, torch.float32, "filt")
    H, W, C = img.shape
    if tuple(filt.shape) != (H + 3 - D, W + 2 - D, C):
        raise ValueError(f"filt must be [{H + 3 - D}, {W + 2 - D}, {C}], got {tuple(filt.shape)}")
    conv = torch.empty((D, D), dtype=torch.float32, device=img.device)
    ngh = torch.empty((D, T), dtype=torch.float32, device=img.device)
    nbytes = lib.mmt_static_context_workspace_bytes(H, D)
    ws = torch.empty((max(nbytes, 16) // 4,), dtype=torch.float32, device=img.device)
    _lib.check(lib.mmt_static_context_f32(_p(img), H, W, C, _p(filt), D, T, float(lam), _p(conv), _p(ngh), _p(ws),
                                          ws.numel() * 4, _stream()), "mmt_static_context_f32")
    return conv, ngh


def sigmoid(x):
    lib = _lib.load()
    _chk(x, torch.float32, "x")
    y = torch.empty_like(x)
    _lib.check(lib.mmt_sigmoid_f32(_p(x), _p(y), x.numel(), _stream()), "mmt_sigmoid_f32")
    return y


def rowsoftmax(x):
    lib = _lib.load()
    _chk(x, torch.float32, "x")
    y = torch.empty_like(x)
    _lib.check(lib.mmt_rowsoftmax_f32(_p(x), _p(y), x.numel() // x.shape[-1], x.shape[-1], _stream()),
               "mmt_rowsoftmax_f32")
    return y


# --------------------------------------------------------------------------------------------
def gemm_tf32(A, B, transA=False, transB=False, out=None, alpha=1.0, accumulate=False):
    """C = alpha * op(A) @ op(B) (+ C) on the tensor cores (mmt_gemm_tf32: TMA tensor maps -> tcgen05 kind::tf32, fp32
    accumulation in TMEM).  A, B: 2-D fp32 CUDA tensors whose last dimension is contiguous (row stride a multiple of 4)."""
    lib = _lib.load()
    for n, t in (("A", A), ("B", B)):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1):
            raise ValueError(f"gemm_tf32: {n} must be a 2-D fp32 CUDA tensor with a contiguous last dimension")
    M, K = (A.shape[1], A.shape[0]) if transA else (A.shape[0], A.shape[1])
    K2, N = (B.shape[1], B.shape[0]) if transB else (B.shape[0], B.shape[1])
    if K != K2:
        raise ValueError(f"gemm_tf32: inner dimensions differ ({K} vs {K2})")
    if out is None:
        if accumulate:
            raise ValueError("gemm_tf32: accumulate needs out")
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    if tuple(out.shape) != (M, N) or out.stride(1) != 1:
        raise ValueError(f"gemm_tf32: out must be [{M}, {N}] with a contiguous last dimension")
    _lib.check(lib.mmt_gemm_tf32(_p(A), A.stride(0), int(transA), _p(B), B.stride(0), int(transB), _p(out), out.stride(0),
                                 M, N, K, float(alpha), int(accumulate), _stream()), "mmt_gemm_tf32")
    return out


# --------------------------------------------------------------------------------------------
_comm_cache = {}


def nccl_comm_ptr(group=None):
    """ncclComm_t of the process group's NCCL backend on the current device, as an integer (None if the group is not
    NCCL-backed): what mmt_allreduce_f32 takes.  The communicator stays owned by torch.distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    pg = group if group is not None else dist.distributed_c10d._get_default_group()
    key = (id(pg), torch.cuda.current_device())
    if key not in _comm_cache:
        try:
            be = pg._get_backend(torch.device("cuda", torch.cuda.current_device()))
            ptr = be._comm_ptr() if hasattr(be, "_comm_ptr") else None
        except Exception:            # noqa: BLE001 -- gloo group, or a torch without _comm_ptr
            ptr = None
        _comm_cache[key] = ptr if ptr else None
    return _comm_cache[key]


def allreduce_(t, op="sum", group=None):
    """In-place all-reduce of a contiguous fp32 CUDA tensor over the ranks of ``group`` through the C-ABI
    (mmt_allreduce_f32 / _max_f32: NCCL over NVLink on the current stream).  Raises if the group has no NCCL communicator."""
    lib = _lib.load()
    _chk(t, torch.float32, "t")
    comm = nccl_comm_ptr(group)
    if comm is None:
        raise RuntimeError("allreduce_: the process group has no NCCL communicator on this device")
    fn = lib.mmt_allreduce_f32 if op == "sum" else lib.mmt_allreduce_max_f32
    _lib.check(fn(C.c_void_p(comm), _p(t), t.numel(), _stream()), "mmt_allreduce_f32")
    return t
