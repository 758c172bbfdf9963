"""GPU tests of the real-data path (BASELINE configs C1 / C5): the ETH / UCY tables under
data/ -> device-side scene windows -> forecaster -> best-of-K ADE/FDE, against the CPU oracle on the same windows and
the same fed noise; the validation-branch and sample.main() mirrors."""
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import scene_batch as o_sb  # noqa: E402
import scores as o_sc  # noqa: E402
import track_a as o_a  # noqa: E402
import track_b as o_b  # noqa: E402

from multimodaltraj_2_b200 import networkx_graph as nx_g  # noqa: E402
from multimodaltraj_2_b200 import ops, realdata, sample, synth, train  # noqa: E402
from multimodaltraj_2_b200.load_traj import DataLoader  # noqa: E402

pytestmark = pytest.mark.gpu


def _args(**kw):
    d = dict(batch_size=16, seq_length=12, pred_len=12, obs_len=8, embedding_size=64, rnn_size=128, K=20,
             precision="fp32", leaveDataset=2, world_size=1, data_root=None, neighborhood_size=64, grid_size=4,
             lambda_param=0.0005, maxNumPeds=20)
    d.update(kw)
    return types.SimpleNamespace(**d)


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("d", [1, 2])
def test_scene_windows_match_oracle_batching(cuda, d):
    """eth/univ (4 rows, no vislet) and ucy/zara01 (6 rows): every obs+pred window of the validation columns."""
    a = _args()
    sc = realdata.scene_windows(a, d, "val", cuda)
    dl = sc["loader"]
    fid, rs, ped, xy, vis = o_sb.table_from_csv(dl.val_data)
    stride = int(np.min(np.diff(fid)))
    wins = fid[:len(fid) - 19]
    want = o_sb.scene_batch(fid, rs, ped, xy, vis, wins, sc["N"], 20, stride)
    keep = want[2].sum(1) > 0
    assert keep.sum() == sc["pos"].shape[0] > 20
    assert np.array_equal(npy(sc["valid"]), want[2][keep]) and np.array_equal(npy(sc["pos"]), want[0][keep])
    assert np.array_equal(npy(sc["slot_ped"]), want[3][keep])
    if vis is None:
        assert float(sc["vis"].abs().max()) == 0.0
    else:
        assert np.array_equal(npy(sc["vis"]), want[1][keep][:, :, :8])
    assert int(sc["valid"].sum(1).max()) <= sc["N"]


@pytest.mark.parametrize("d,prec,tol", [(1, ops.PREC_F32, 1e-3), (2, ops.PREC_F32, 1e-3), (1, ops.PREC_BF16, 1e-3), (1, ops.PREC_F16, 2e-4),
                                        (2, ops.PREC_F16, 2e-4)])
def test_best_of_k_on_real_split_matches_oracle(cuda, d, prec, tol):
    """C1: forward + best-of-20 ADE/FDE on a real split against the oracle with the same fed noise (fp32 mode: the
    north_star's 1e-3; bf16 stated separately), in the table's units and -- ETH -- in metres."""
    a = _args()
    sc = realdata.scene_windows(a, d, "val", cuda)
    n = min(96, sc["pos"].shape[0])
    sub = {k: sc[k][:n].contiguous() for k in ("pos", "vis", "valid")}
    sub["N"] = sc["N"]
    p = synth.init_params(seed=0)
    eps = o_b.philox_eps(7, n, sc["N"], 20, 12)
    res = realdata.evaluate_split(a, d, ops.CellParams.from_numpy(p, cuda), prec=prec, device=cuda, scenes=sub,
                                  eps=torch.from_numpy(eps).to(cuda))
    pos, vis, valid = (npy(sub[k]) for k in ("pos", "vis", "valid"))
    want = o_b.forecast(pos, vis, valid, p, eps, 8, 12)
    v = valid.astype(bool)
    pick = want["best_k"][..., None] == np.arange(20)[None, None, :]
    w_ade, w_fde = (want["ade"] * pick).sum(-1), (want["fde"] * pick).sum(-1)
    g = res["_out"]
    assert np.abs(npy(g["best_ade"])[v] - w_ade[v]).max() < tol and np.abs(npy(g["best_fde"])[v] - w_fde[v]).max() < tol
    assert res["n_agents"] == int(v.sum()) and res["ade"] == pytest.approx(float(w_ade[v].mean()), abs=tol)
    if prec == ops.PREC_F32:
        assert (npy(g["best_k"])[v] == want["best_k"][v]).mean() > 0.995          # near-ties may flip at 1e-6
    if d in realdata.HOMOGRAPHY:
        wa, wf = o_sc.ade_fde_world(want["best_traj"].reshape(-1, 12, 2), pos[:, :, 8:].reshape(-1, 12, 2),
                                    np.array(realdata.HOMOGRAPHY[d]), valid.reshape(-1))
        assert res["ade_m"] == pytest.approx(float(wa.sum() / v.sum()), rel=5e-3 if prec == ops.PREC_F32 else 5e-2)
        assert 0.05 < res["ade_m"] < 50.0                                        # metres, not pixels


def test_split_evaluation_is_independent_of_the_sharding(cuda):
    a = _args(K=20)
    p = ops.CellParams.from_numpy(synth.init_params(seed=0), cuda)
    sc = realdata.scene_windows(a, 0, "val", cuda)
    whole = realdata.evaluate_split(a, 0, p, prec=ops.PREC_BF16, device=cuda, scenes=sc)
    parts = [realdata.evaluate_split(a, 0, p, prec=ops.PREC_BF16, rank=r, world=3, device=cuda, scenes=sc) for r in range(3)]
    assert sum(x["n_agents"] for x in parts) == whole["n_agents"]
    tot = sum(x["ade"] * x["n_agents"] for x in parts) / whole["n_agents"]
    assert tot == pytest.approx(whole["ade"], rel=1e-6)


def test_validation_branch_mirror(cuda):
    """train.py:371-695: reference loader + ConstructGraph + the as-written model per batch, scores (i); the first
    batch is checked against the oracle's restatement of the same lines; then the north_star metric on the split."""
    a = _args(leaveDataset=2)
    dl = DataLoader(a, datasets=[0, 1, 2, 3, 4, 5], start=2, sel=0)
    dl.reset_data_pointer(valid=True, frame_pointer=dl.seed)
    graph = nx_g.online_graph(a)
    batch, targets, fp = dl.next_step()
    assert len(batch) > 0
    r = train.track_a_batch(a, batch, targets, graph, fp, device=cuda)
    n, T, dim = r["num_nodes"], 8, 16
    assert 0 < n <= 8 and tuple(r["pred"].shape) == (n, 12, 2)
    # oracle: the same lines (train.py:438-444 input build, :514-636 model + state step, :639-662 scores)
    g2 = nx_g.online_graph(a).ConstructGraph(current_batch=batch, framenum=fp, future_traj=targets)
    bv = np.array(list(g2.get_node_attr('node_pos_list').values()))[1:1 + T]
    X = np.linalg.norm(bv, axis=2).T
    rng = np.random.default_rng(0)
    W_i, W_ii = rng.standard_normal((n, dim)), rng.standard_normal((dim, T))
    from multimodaltraj_2_b200.models import g2k_lstm_mcr as mcr
    m = mcr.g2k_lstm_mcr(in_features=torch.zeros((dim, dim)), num_nodes=n, obs_len=T, hidden_size=128, lambda_reg=0.0005)
    w = dict(W_i=W_i, W_ii=W_ii, W_v=npy(m.weight_v).astype(np.float64), b_v=npy(m.bias_v).astype(np.float64),
             W_r=npy(m.weight_r).astype(np.float64), W_c=npy(m.weight_c).astype(np.float64), W_o=npy(m.weight_o).astype(np.float64))
    ref = o_a.mcr_scene_loop(X[None], np.zeros((1, 2, n)), np.zeros((1, dim, dim)), np.zeros((1, dim, 128)), w, 0.0005, 12,
                             frames=len(batch))
    np.testing.assert_allclose(npy(r["pred"]), ref["pred"][0], rtol=1e-4, atol=1e-6)
    ids = list(targets)[:n]
    _, _, oeuc, oerr = o_sc.train_val_scores(ref["pred"][0], [np.asarray(targets[k], np.float64)[:12] for k in ids],
                                            n_targets=len(targets))
    np.testing.assert_allclose(npy(r["euc"]), oeuc, rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(npy(r["err"]), oerr, rtol=1e-4, atol=1e-6)
    res = train.validate(a, l=2, device=cuda, max_batches=4)
    assert res["batches"] >= 1 and np.isfinite(res["cv_ade"]) and np.isfinite(res["cv_fde"])
    assert res["best_of_k"]["n_agents"] > 100 and np.isfinite(res["best_of_k"]["ade"])


def test_sample_main_mirror(cuda, tmp_path, capsys):
    out = sample.main(["--leaveDataset", "1", "--precision", "fp32"], device=cuda, results_path=str(tmp_path / "results.pkl"))
    txt = capsys.readouterr().out
    assert "wall-clock time taken by" in txt and "Total mean error of the model is" in txt
    assert out["scenes_scored"] > 0 and np.isfinite(out["mean_error_ade"]) and out["n_agents"] > 0
    import pickle
    res = pickle.load(open(tmp_path / "results.pkl", "rb"))
    assert len(res) == out["scenes_scored"] and res[0][1].shape[1] == 20


def test_drop_in_model_classes_reach_the_batched_path(cuda):
    """models.g2k_lstm_mc / g2k_lstm_mcr.forecast_batched == ops.Forecaster (mc in bf16 mode: the fused rollout kernel)."""
    from multimodaltraj_2_b200.models import g2k_lstm_mc as mc
    from multimodaltraj_2_b200.models import g2k_lstm_mcr as mcr
    S, N = 6, 64
    pos, vis, valid = (torch.as_tensor(a).to(cuda) for a in synth.make_crowd(S, N, seed=4, half_extent=4.0, ragged=True))
    p = ops.CellParams.from_numpy(synth.init_params(seed=0), cuda)
    m = mc.g2k_lstm_mc(in_features=torch.zeros((16, 16)), out_size=128, obs_len=8, num_nodes=N, lambda_reg=0.0005)
    got = m.forecast_batched(pos, vis, valid, p, seed=3)
    want = ops.rollout_f16(pos, vis, valid, p)
    assert torch.equal(got["params"], want)                              # the fused kernel (default mode: f16), bit for bit
    assert torch.equal(m.forecast_batched(pos, vis, valid, p, seed=3, prec=ops.PREC_BF16)["params"], ops.rollout_bf16(pos, vis, valid, p))
    r = mcr.g2k_lstm_mcr(in_features=torch.zeros((16, 16)), hidden_size=128, obs_len=8, num_nodes=N, lambda_reg=0.0005)
    got_r = r.forecast_batched(pos, vis, valid, p, seed=3)
    ref = ops.Forecaster(p, S, N, 8, 12, 20, relational=True, prec=ops.PREC_F16, seed=3, device=cuda)(pos, vis, valid)
    assert torch.equal(got_r["best_k"], ref["best_k"]) and not torch.equal(got_r["params"], got["params"])


def test_train_driver_on_the_shipped_tables(cuda):
    """train.train (train.py:23-366 outer loop) over the real UCY tables under data/: leave-one-out, both variants."""
    a = _args(leaveDataset=2, max_agents=64, precision="bf16", save_dir=None)
    res = train.train(a, datasets=(2, 3), rank=0, world=1, device=cuda)
    assert set(res) == {3} and res[3]["n_agents"] > 1000 and np.isfinite(res[3]["ade"])
    a.variant = "mc"
    res_mc = train.train(a, datasets=(2, 3), rank=0, world=1, device=cuda)
    assert res_mc[3]["n_agents"] == res[3]["n_agents"] and res_mc[3]["ade"] != res[3]["ade"]
