"""GPU tests of the host-side mirror of the reference interface: same class / function names,
constructor keywords and attributes as the reference's models, encoders, scoring and loader,
running on the CUDA kernels and checked against the oracle."""
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import gridlstm as o_gl  # noqa: E402
import scene_batch as o_sb  # noqa: E402
import scores as o_sc  # noqa: E402
import track_a as o_a  # noqa: E402
import track_b as o_b  # noqa: E402

from multimodaltraj_2_b200 import helper, ops, sample, synth, train  # noqa: E402
from multimodaltraj_2_b200.load_traj import DataLoader  # noqa: E402
from multimodaltraj_2_b200.models import g2k_lstm_mc as mc  # noqa: E402
from multimodaltraj_2_b200.models import g2k_lstm_mcr as mcr  # noqa: E402
from multimodaltraj_2_b200.models import gsk_lstm_cell as gsk  # noqa: E402
from multimodaltraj_2_b200.relational_inf_models import nri_learned  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(ROOT / "tests" / "golden" / "track_a_ckpt.npz")
ZARA = np.load(ROOT / "tests" / "golden" / "zara01_slice.npz")


def npy(t):
    return t.detach().cpu().numpy()


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_g2k_lstm_mcr_reference_call_pattern(cuda):
    """train.py:161-165,226-254: construct with the reference's keywords, feed the placeholders, read
    pred_path_band [2,P,n] and transpose to [n,P,2]."""
    D, T, n, H, lam = 10, 8, 9, 128, 0.0005
    sess_g = {"krnl_weights_21/weight_v:0": GOLD["seed0_weight_v"], "krnl_weights_21/bias_v:0": GOLD["seed0_bias_v"],
              "krnl_weights_21/weight_c:0": GOLD["seed0_weight_c"], "krnl_embed_21/weight_r:0": GOLD["seed0_weight_r"],
              "krnl_weights_21/cost:0": GOLD["seed0_cost"], "krnl_weights_21/attn:0": GOLD["seed0_attn"]}
    in_features = torch.zeros((D, D))
    m = mcr.g2k_lstm_mcr(in_features=in_features, num_nodes=n, obs_len=T, hidden_size=H, lambda_reg=lam, sess_g=sess_g)
    for attr in ("weight_r", "weight_v", "bias_v", "cost", "attn", "weight_c", "weight_o", "outputs", "ngh",
                 "rel_features", "hidden_states", "out_size", "pred_path_band"):
        assert hasattr(m, attr)
    assert np.allclose(npy(m.weight_v), GOLD["seed0_weight_v"]) and tuple(m.pred_path_band.shape) == (2, 12, n)
    rng = np.random.default_rng(0)
    outputs, rel, ngh = rng.standard_normal((D + 2, D)), rng.standard_normal((2, D)), rng.standard_normal((D, T)) * 100
    m.outputs, m.rel_features, m.ngh = (torch.as_tensor(a, dtype=torch.float32).cuda() for a in (outputs, rel, ngh))
    band = m.forward()
    want = o_a.mcr_forward(outputs, rel, ngh, GOLD["seed0_weight_v"], GOLD["seed0_bias_v"], GOLD["seed0_weight_r"],
                           GOLD["seed0_weight_c"], npy(m.weight_o).astype(np.float64), lam)
    assert rel_err(npy(band), want["band"]) < 1e-4 and rel_err(npy(m.attn), want["attn"]) < 1e-4
    assert rel_err(npy(m.cost), want["cost"]) < 1e-4
    pred = np.transpose(npy(m.pred_path_band), (2, 1, 0))                 # train.py:254
    assert pred.shape == (n, 12, 2)
    # batched entry point == per-scene oracle loop
    S = 5
    X, V = np.abs(rng.standard_normal((S, T, n))), rng.standard_normal((S, 2, n))
    C, Hs = rng.standard_normal((S, D, D)) * 100, rng.standard_normal((S, D, H))
    W_i, W_ii = rng.standard_normal((n, D)), GOLD["seed0_weight_ii"]
    f32 = lambda a: torch.as_tensor(np.asarray(a, np.float32)).cuda()  # noqa: E731
    got = m.forward_batched(f32(X), f32(V), f32(C), f32(Hs), f32(W_i), f32(W_ii))
    w = dict(W_i=W_i, W_ii=W_ii, W_v=GOLD["seed0_weight_v"], b_v=GOLD["seed0_bias_v"], W_r=GOLD["seed0_weight_r"],
             W_c=GOLD["seed0_weight_c"], W_o=npy(m.weight_o).astype(np.float64))
    ref = o_a.mcr_scene_loop(X, V, C, Hs, w, lam, 12)
    for k in ("attn", "band", "Hs"):
        assert rel_err(npy(got[k]), ref[k]) < 1e-4


def test_g2k_lstm_mc_band_is_zero(cuda):
    m = mc.g2k_lstm_mc(in_features=torch.zeros((16, 16)), out_size=128, obs_len=8, num_nodes=7, lambda_reg=0.0005)
    assert tuple(m.pred_path_band.shape) == (2, 12, 7) and float(m.pred_path_band.abs().max()) == 0.0
    assert float(m.cost.abs().max()) == 0.0 and tuple(m.visual_path.shape) == (1, 16)


def test_gsk_lstm_cell_object(cuda):
    cell = gsk.gsk_lstm_cell(in_features=torch.zeros((16, 16)), out_size=128, obs_len=8, num_nodes=64,
                             lambda_reg=0.0005, precision="fp32")
    R = 96
    rng = np.random.default_rng(1)
    x = torch.as_tensor((rng.standard_normal((R, 4)) * 0.3).astype(np.float32)).cuda()
    h, c = cell.init_state(R)
    valid = torch.ones(R, dtype=torch.uint8, device="cuda")
    h1, c1, mf = cell(x, h, c, h, c, valid)
    p = synth.init_params(seed=0)
    oh, oc, of = o_b.gsk_cell(npy(x)[None], npy(h)[None], npy(c)[None], npy(h)[None], npy(c)[None], npy(valid)[None], p)
    assert rel_err(npy(h1), oh[0]) < 1e-4 and rel_err(npy(c1), oc[0]) < 1e-4


def test_neighbourhood_encoders(cuda):
    ck = {"grid_lstm_cell/" + k[6:]: GOLD[k] for k in GOLD.files if k.startswith("glstm_")}
    enc = helper.neighborhood_vis_loc_encoder(hidden_size=128, hidden_len=16, num_layers=2, grid_size=4,
                                              embedding_size=64, dropout=0.8, sess_g=ck)
    for attr in ("input", "state_f00_b00_c", "c_hidden_state", "output"):
        assert hasattr(enc, attr)
    rng = np.random.default_rng(2)
    x, st = rng.standard_normal((16, 16)), rng.standard_normal((16, 128)) * 0.5
    enc.input = torch.as_tensor(x, dtype=torch.float32).cuda()
    enc.state_f00_b00_c = torch.as_tensor(st, dtype=torch.float32).cuda()
    out, cst = enc.forward()
    pe = [GOLD[k] for k in ("glstm_W_I_diag_freqf_0", "glstm_W_I_diag_freqt_0", "glstm_W_O_diag_freqf_0",
                            "glstm_W_O_diag_freqt_0")]
    om, os_ = o_gl.gridlstm_step(x, st, GOLD["glstm_W_f_0_0"], GOLD["glstm_B_f_0"], *pe, U=2, F=4)
    assert rel_err(npy(out), om) < 1e-4 and rel_err(npy(cst), os_) < 1e-4
    stat = helper.neighborhood_stat_enc(ctxt_path=[], hidden_size=128, num_layers=2, grid_size=4, dim=16, shared=enc.w)
    stat.input = torch.as_tensor(x[:, :8], dtype=torch.float32).cuda().contiguous()
    stat.hidden_state = enc.state_f00_b00_c
    o2, _ = stat.forward()
    om2, _ = o_gl.gridlstm_step(x[:, :8], st, GOLD["glstm_W_f_0_0"], GOLD["glstm_B_f_0"], *pe, U=2, F=2, peepholes=False)
    assert rel_err(npy(o2), om2) < 1e-4
    assert float(enc.init_hidden(5).abs().sum()) == 0 and tuple(enc.init_hidden(5).shape) == (5, 128)


def test_nri_learned_functions(cuda):
    rng = np.random.default_rng(3)
    a = torch.as_tensor(rng.standard_normal((7, 7)).astype(np.float32)).cuda()
    np.testing.assert_allclose(npy(nri_learned.infer_rlns(a)), 1 / (1 + np.exp(-npy(a))), rtol=1e-5)
    np.testing.assert_allclose(npy(nri_learned.eval_rln_ngh(a, None)), o_a.softmax_last(npy(a)), rtol=1e-5)
    p = synth.init_params(seed=4)
    P = ops.CellParams.from_numpy(p, "cuda")
    pos, _, valid = synth.make_crowd(2, 16, seed=9, half_extent=3.0)
    _, adj, _ = o_b.pairwise_adj(pos[:, :, 0], valid, 4.0, 0.5)
    h = (rng.standard_normal((2, 16, 128)) * 0.5).astype(np.float32)
    sc = nri_learned.graph_to_kernel(torch.as_tensor(h).cuda(), torch.as_tensor(adj).cuda(), P)
    np.testing.assert_allclose(npy(sc), o_b.edge_mlp(h, adj, p), rtol=1e-4, atol=1e-5)


def test_reference_scores_on_device(cuda):
    rng = np.random.default_rng(5)
    n, L, obs = 20, 20, 8
    pred, true = rng.standard_normal((n, L, 2)), rng.standard_normal((n, L, 2))
    got = sample.get_mean_error(pred, true, obs, n)
    want = o_sc.get_mean_error(pred, true, obs, n)
    assert got[2] == want[2] and got[0] == pytest.approx(want[0], rel=1e-4) and got[1] == pytest.approx(want[1], rel=1e-4)
    P = 12
    lens = np.array([12] * 15 + [7, 3, 12, 1, 9], np.int32)
    p12, t12 = rng.standard_normal((n, P, 2)).astype(np.float32), rng.standard_normal((n, P, 2)).astype(np.float32)
    euc, err = ops.train_val_scores(torch.as_tensor(p12).cuda(), torch.as_tensor(t12).cuda(),
                                    torch.as_tensor(lens).cuda(), n)
    _, _, oeuc, oerr = o_sc.train_val_scores(p12.astype(np.float64), [t12[i, :lens[i]].astype(np.float64) for i in range(n)])
    np.testing.assert_allclose(npy(euc), oeuc, rtol=2e-4)
    np.testing.assert_allclose(npy(err), oerr, rtol=1e-5, atol=1e-6)


def test_loader_device_batching_and_train_driver(cuda, tmp_path):
    args = types.SimpleNamespace(batch_size=16, seq_length=12, pred_len=12, obs_len=8, embedding_size=64, rnn_size=128,
                                 max_agents=16, K=20, precision="fp32", leaveDataset=3, world_size=1,
                                 data_root=str(tmp_path))
    d = tmp_path / "ucy/zara/zara01"
    d.mkdir(parents=True)
    np.savetxt(d / "vis_body.csv", ZARA["csv"], delimiter=",", fmt="%.10g")
    dl = DataLoader(args, datasets=[0, 1, 2, 3, 4, 5, 6], sel=0, start=2)
    table = dl.device_table("cuda")
    pos, vis, valid, slot = dl.scene_batch(table, 16, 20)
    fid, rs, ped, xy, v = o_sb.table_from_csv(ZARA["csv"][:, :dl.max])
    want = o_sb.scene_batch(fid, rs, ped, xy, v, fid[:len(fid) - 19], 16, 20, 8)
    assert np.array_equal(npy(valid), want[2]) and np.array_equal(npy(pos), want[0]) and np.array_equal(npy(slot), want[3])
    assert int(valid.sum()) > 0
    res = train.train(args, datasets=(2,), rank=0, world=1)
    assert res[2]["n_agents"] == int(valid.sum()) and np.isfinite(res[2]["ade"]) and res[2]["ade"] > 0
    # sharded over two "ranks" (sequentially on one GPU): partial sums add up to the single-rank result
    a = train.train(args, datasets=(2,), rank=0, world=2)
    b = train.train(args, datasets=(2,), rank=1, world=2)
    tot = a[2]["ade"] * a[2]["n_agents"] + b[2]["ade"] * b[2]["n_agents"]
    assert a[2]["n_agents"] + b[2]["n_agents"] == res[2]["n_agents"]
    assert tot / res[2]["n_agents"] == pytest.approx(res[2]["ade"], rel=1e-4)
