"""CPU tests of the host-side logic: DataLoader mirror vs what the reference's own DataLoader
produced (golden fixture), the scene-batching oracle, sharding, graph construction, and a
world_size-2 gloo run of the scene-sharded scoring."""
import os
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import scene_batch as o_sb  # noqa: E402

G = np.load(ROOT / "tests" / "golden" / "zara01_slice.npz")
ARGS = types.SimpleNamespace(batch_size=16, seq_length=12, pred_len=12, obs_len=8)


def _loader():
    from multimodaltraj_2_b200.load_traj import DataLoader
    return DataLoader(ARGS, datasets=[0, 1, 2, 3, 4], sel=0, start=2, csv=G["csv"])


def test_dataloader_matches_reference_dataloader():
    dl = _loader()
    assert dl.num_batches == int(G["num_batches"]) and dl.max == int(G["max"]) and dl.val_max == int(G["val_max"])
    assert dl.seed == float(G["seed"]) and dl.len == G["csv"].shape[1]
    rows = np.array([(fr, p, pos[0], pos[1]) for fr, lst in dl.trajectories.items() for it in lst
                     for p, pos in it.items()])
    ref = G["traj_rows"]
    assert np.array_equal(rows[np.lexsort((rows[:, 1], rows[:, 0]))], ref[np.lexsort((ref[:, 1], ref[:, 0]))])
    for c in range(3):
        batch, targets, fp = dl.next_step()
        assert np.array_equal(np.array(sorted(batch), float), G[f"batch{c}_frames"])
        assert fp == float(G[f"batch{c}_fp"]) and len(targets) == int(G[f"batch{c}_n_targets"])
        assert np.array_equal(np.array([len(v) for v in targets.values()]), G[f"batch{c}_target_lens"])
    dl.reset_data_pointer()
    assert dl.frame_pointer == dl.seed


def test_construct_graph_node_arrays():
    from multimodaltraj_2_b200.networkx_graph import online_graph
    dl = _loader()
    batch, targets, _ = dl.next_step()
    g = online_graph(ARGS).ConstructGraph(current_batch=batch, framenum=1, future_traj=targets)
    pl = g.get_node_attr('node_pos_list')
    arr = np.array(list(pl.values()))
    assert arr.ndim == 3 and arr.shape[1:] == (8, 2) and len(pl) == len(set(pl))
    # first sighting leaves a zero row (reference defect F-5): frame 0's row is zero for peds first seen in frame 0
    first = next(iter(batch.values()))
    (ped0, _), = first[0].items()
    assert np.all(pl[int(ped0)][0] == 0)


def test_scene_batch_oracle_rules():
    csv = G["csv"][:, :int(G["max"])]
    fid, rs, ped, xy, vis = o_sb.table_from_csv(csv)
    assert np.all(np.diff(fid) > 0) and rs[-1] == csv.shape[1]
    pos, vo, valid, slot = o_sb.scene_batch(fid, rs, ped, xy, vis, fid[:5], 8, 4, 8)
    for s in range(5):
        ids = slot[s][valid[s] == 1]
        assert np.all(np.diff(ids) > 0)                         # ascending ped id
        for k, p in enumerate(ids):                              # every slot really is that ped in every frame
            for f in range(4):
                col = np.where((csv[0] == fid[s] + 8 * f) & (csv[1] == p))[0]
                assert len(col) == 1 and np.allclose(pos[s, k, f], csv[2:4, col[0]])
    # a window that runs off the end has no valid agents
    _, _, v, _ = o_sb.scene_batch(fid, rs, ped, xy, None, fid[-1:], 8, 4, 8)
    assert v.sum() == 0


def test_shard_range_partitions_scenes():
    from multimodaltraj_2_b200.train import shard_range
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def _gloo_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "oracle"))
    import track_b as o_b
    from multimodaltraj_2_b200 import synth
    from multimodaltraj_2_b200.train import shard_range
    S, N, T, P, K = 6, 8, 8, 12, 4
    pos, vis, valid = synth.make_crowd(S, N, seed=3, half_extent=3.0, ragged=True)
    p = synth.init_params(seed=2)
    lo, hi = shard_range(S, rank, world)
    eps = o_b.philox_eps(9, S, N, K, P)[lo:hi]                  # global agent indexing == agent_offset = lo*N
    o = o_b.forecast(pos[lo:hi], vis[lo:hi], valid[lo:hi], p, eps, T, P)
    sel = np.take_along_axis(o["ade"], np.maximum(o["best_k"], 0)[..., None], -1)[..., 0]
    sums = torch.tensor([float(sel.sum()), float(valid[lo:hi].sum())], dtype=torch.float64)
    torch.distributed.all_reduce(sums)
    if rank == 0:
        full = o_b.forecast(pos, vis, valid, p, o_b.philox_eps(9, S, N, K, P), T, P)
        fsel = np.take_along_axis(full["ade"], np.maximum(full["best_k"], 0)[..., None], -1)[..., 0]
        ret["ok"] = bool(abs(float(sums[0]) - float(fsel.sum())) < 1e-4 and int(sums[1]) == int(valid.sum()))
    torch.distributed.destroy_process_group()


def test_scene_sharding_world_size_2_gloo():
    import torch.multiprocessing as mp
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_gloo_worker, args=(2, 29533, ret), nprocs=2, join=True)
        assert ret.get("ok") is True


def _gloo_train_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "oracle"))
    import train_b as o_t
    from multimodaltraj_2_b200 import synth
    from multimodaltraj_2_b200.train import TRAIN_KEYS, allreduce_mean_, flatten_bucket, rmsprop_update_, unflatten_bucket
    p = {k: torch.tensor(v) for k, v in synth.init_params(seed=4).items() if k in TRAIN_KEYS}
    # every rank holds the SUM-gradient of its own scene shard and its own count of valid agent-steps
    gen = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    shard = [{k: torch.randn(p[k].shape, generator=gen[r]) for k in TRAIN_KEYS} for r in range(world)]
    cnt = [torch.tensor([3.0 + r, 40.0 + 8 * r]) for r in range(world)]          # (loss sum, valid agent-steps)
    flat, counts = flatten_bucket(shard[rank]), cnt[rank].clone()
    allreduce_mean_(flat, counts)                                                # the ONE gradient all-reduce
    g = unflatten_bucket(flat / counts[1], p)
    ms = {k: torch.zeros_like(p[k]) for k in TRAIN_KEYS}
    rmsprop_update_(p, g, ms)
    if rank == 0:
        n = sum(float(c[1]) for c in cnt)
        want_g = {k: sum(s[k] for s in shard).numpy().astype(np.float64) / n for k in TRAIN_KEYS}
        p0 = {k: v.astype(np.float64) for k, v in synth.init_params(seed=4).items() if k in TRAIN_KEYS}
        want_p, _ = o_t.rmsprop_step(p0, want_g, {k: np.zeros_like(p0[k]) for k in TRAIN_KEYS})
        ok = all(np.abs(g[k].numpy() - want_g[k]).max() < 1e-6 for k in TRAIN_KEYS)
        ok = ok and all(np.abs(p[k].numpy() - want_p[k]).max() < 1e-5 for k in TRAIN_KEYS)
        ret["ok"] = bool(ok and abs(float(counts[1]) - n) < 1e-6)
    torch.distributed.destroy_process_group()


def test_gradient_bucket_allreduce_and_rmsprop_world_size_2_gloo():
    """Host logic of the data-parallel training step: flat bucket, one SUM all-reduce, division by the global count
    of valid agent-steps, clipped RMSProp -- every rank ends with the update the oracle computes from the pooled data."""
    import torch.multiprocessing as mp
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_gloo_train_worker, args=(2, 29541, ret), nprocs=2, join=True)
        assert ret.get("ok") is True


def test_train_oracle_gradient_matches_finite_differences():
    """Pins the training oracle itself: autograd gradient of the teacher-forced NLL vs central differences."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import train_b as o_t
    from multimodaltraj_2_b200 import synth
    pos, vis, valid = synth.make_crowd(2, 8, seed=1, half_extent=2.0, ragged=True)
    p = {k: np.asarray(v, np.float64) for k, v in synth.init_params(seed=3).items()}
    _, g = o_t.loss_and_grads(pos, vis, valid, p)
    rng = np.random.default_rng(0)
    for k in ("W", "W_h", "w_It", "b_e"):
        idx = tuple(rng.integers(0, s) for s in p[k].shape)
        h = 1e-5
        pp, pm = {**p, k: p[k].copy()}, {**p, k: p[k].copy()}
        pp[k][idx] += h
        pm[k][idx] -= h
        fd = (o_t.loss_and_grads(pos, vis, valid, pp)[0] - o_t.loss_and_grads(pos, vis, valid, pm)[0]) / (2 * h)
        assert abs(fd - g[k][idx]) < 1e-6 + 1e-4 * abs(fd), (k, fd, g[k][idx])


# ------------------------------------------------------------------------------------------------
# TensorFlow-bundle checkpoints (train.py:330-343 save, :383-402 restore) without TensorFlow
def _ref_ckpt_prefixes():
    from pathlib import Path
    gold = Path(__file__).resolve().parent / "golden" / "ref_ckpt"
    out = sorted({str(f)[:-len(".index")] for f in gold.glob("*.index")})
    ref = Path("/root/reference/save")           # in the build container: every checkpoint the reference ships
    if ref.exists():
        out += sorted({str(f)[:-len(".index")] for f in ref.glob("*.index")})
    return out


def test_tf_bundle_writer_reproduces_the_reference_checkpoints_byte_for_byte(tmp_path):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
    import tf_bundle as o_tb                      # the oracle's independent reader
    from multimodaltraj_2_b200 import tf_bundle
    prefixes = _ref_ckpt_prefixes()
    assert len(prefixes) >= 2
    for n, pre in enumerate(prefixes):
        got, want = tf_bundle.read_checkpoint(pre), o_tb.read_checkpoint(pre)
        assert set(got) == set(want)
        for k in want:
            assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), k
        out = tmp_path / f"c{n}"
        tf_bundle.write_checkpoint(out, got)
        for ext in (".index", ".data-00000-of-00001"):
            assert Path(str(out) + ext).read_bytes() == Path(pre + ext).read_bytes(), (pre, ext)


def test_tf_bundle_checksums_and_known_crc():
    from multimodaltraj_2_b200 import tf_bundle
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283          # the CRC-32C check value
    assert tf_bundle.crc32c(b"\x00" * 32) == 0x8A9136AA           # RFC 3720 B.4
    assert tf_bundle.crc32c(bytes(range(32))) == 0x46DD794E


def test_checkpoint_save_resume_round_trip(tmp_path):
    import torch
    from multimodaltraj_2_b200 import ops, synth, tf_bundle
    from multimodaltraj_2_b200.train import Trainer, TRAIN_KEYS, save_checkpoint, load_checkpoint
    p = ops.CellParams.from_numpy(synth.init_params(seed=4), "cpu")
    tr = Trainer(p)
    for k in TRAIN_KEYS:
        tr.ms[k] = torch.rand_like(tr.ms[k])
    prefix = save_checkpoint(tmp_path / "g2k_MPC_model_kfold_train_2_0_0.ckpt", p, trainer=tr, global_step=50)
    assert prefix.endswith(".ckpt-50") and tf_bundle.latest_checkpoint(tmp_path) == prefix
    tr2 = Trainer(ops.CellParams.from_numpy(synth.init_params(seed=5), "cpu"))
    p2 = load_checkpoint(tmp_path, device="cpu", trainer=tr2)      # through the `checkpoint` state file
    for k in TRAIN_KEYS:
        assert torch.equal(getattr(p2, k), getattr(p, k)) and torch.equal(tr2.ms[k], tr.ms[k]), k
    assert tr2.p is p2
    # a flipped byte in the data file is caught by the per-tensor CRC-32C
    f = tmp_path / "g2k_MPC_model_kfold_train_2_0_0.ckpt-50.data-00000-of-00001"
    raw = bytearray(f.read_bytes())
    raw[100] ^= 0x40
    f.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        tf_bundle.read_checkpoint(prefix)
    # flat torch file
    save_checkpoint(tmp_path / "w.pt", p)
    p3 = load_checkpoint(tmp_path / "w.pt", device="cpu")
    assert torch.equal(p3.W, p.W)


def test_bench_prints_one_json_line_on_stdout():
    """bench.py's contract: ONE JSON line on stdout.  Anything a library prints to file descriptor 1 (NCCL's version
    banner at N > 1) must end up on stderr: the line goes to the saved original descriptor."""
    import json
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench._claim_stdout(); "
            "print('library noise'); os.write(1, b'raw noise\\n'); bench.emit({'metric': 'm', 'value': 1.5})" % str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("\n") == 1 and json.loads(r.stdout) == {"metric": "m", "value": 1.5}
    assert "library noise" in r.stderr and "raw noise" in r.stderr


# ------------------------------------------------------------------------------------------------
def test_operand_rounding_emulation_explains_the_bf16_error_and_what_fp16_leaves():
    """DESIGN.md section 5, on the CPU: the oracle's rollout with its GEMM / aggregation operands rounded the way the fused
    tensor-core kernel rounds them.  No rounding reproduces the oracle; bf16 operands move ADE / FDE by what the GPU's bf16 mode
    measures; operands with 11 significand bits (fp16 = the MMT_PREC_F16 mode; tf32, the same significand with the fp32
    exponent, gives the same numbers to 5e-6) leave several times less."""
    import operand_rounding as o_r
    import track_b as o_b
    from multimodaltraj_2_b200 import synth
    S, N, T, P, K = 6, 64, 8, 12, 20
    p = synth.init_params(seed=0)
    pos, vis, valid = synth.make_crowd(S, N, seed=synth.SEED)
    eps = o_b.philox_eps(0xB200, S, N, K, P)
    want = o_b.forecast(pos, vis, valid, p, eps, T, P)
    v = valid.astype(bool)
    err = {}
    for name, rnd in (("none", o_r.r_none), ("bf16", o_r.r_bf16), ("tf32", o_r.r_tf32), ("f16", o_r.r_f16)):
        par = o_r.rollout(pos, vis, valid, p, rnd)
        ade, fde, *_ = o_b.decode_score(par, eps, pos[:, :, T - 1], pos[:, :, T:T + P], valid)
        err[name] = (float(np.abs(ade[v] - want["ade"][v]).max()), float(np.abs(fde[v] - want["fde"][v]).max()))
    assert err["none"][0] < 5e-6 and err["none"][1] < 5e-6          # the restructured softmax (numerators x 1/sum) is the oracle's
    # the fp16 exponent range costs nothing that matters (only weights below 6e-5 lose bits to its subnormals)
    assert abs(err["f16"][0] - err["tf32"][0]) < 5e-6 and abs(err["f16"][1] - err["tf32"][1]) < 5e-6
    assert err["f16"][0] < 1e-4 and err["f16"][1] < 2e-4
    assert err["bf16"][0] > 4 * err["f16"][0] and err["bf16"][1] > 4 * err["f16"][1]


def test_precision_names_of_the_drivers():
    from multimodaltraj_2_b200 import argParser, ops
    assert ops.prec_from_name("fp16") == ops.PREC_F16 == 4 and ops.prec_from_name("bf16") == ops.PREC_BF16
    assert ops.prec_from_name("fp32") == ops.PREC_F32 and ops.prec_from_name("bf16x3") == ops.PREC_BF16X3
    with pytest.raises(ValueError):
        ops.prec_from_name("int8")
    assert argParser.ArgsParser.parser.parse_args([]).precision == "fp16"      # the drivers' default: inside the 1e-3 bar
