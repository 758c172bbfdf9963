"""CPU tests of the round-2 host logic: scene sharding, agent padding, the shipped data tables through the mirror loader,
the build stamp, and the bench line's post-mortem helper."""
import importlib
import json
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_shard_ranges_partition_the_scenes():
    from multimodaltraj_2_b200.realdata import pad_agents, shard_range
    for n in (0, 1, 7, 253, 4096):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    assert [pad_agents(k) for k in (1, 8, 9, 27, 33, 64, 65, 200)] == [8, 8, 16, 32, 64, 64, 128, 256]


def test_shipped_tables_load_like_the_reference_expects():
    """data/: the alphabetically first table of each directory (load_traj.py:77-86), 4 rows (ETH) or 6 (UCY); the batch
    counts of SURVEY App. E (22 / 43 / 80 for zara01 / zara02 / ucy-univ)."""
    from multimodaltraj_2_b200.load_traj import DataLoader
    args = types.SimpleNamespace(batch_size=16, seq_length=12, pred_len=12, obs_len=8)
    want = {0: (4, 6544, None), 1: (4, 5492, None), 2: (6, 6279, 22), 3: (6, 11930, 43), 4: (6, 21985, 80)}
    for d, (rows, cols, nb) in want.items():
        dl = DataLoader(args, sel=0, start=d, parent_dir=str(ROOT / "data"))
        assert dl.raw_data.shape == (rows, cols) and dl.max == int(cols * 0.7)
        assert dl.sel_file.endswith("pixel_pos.csv.gz" if rows == 4 else "vis_body.csv.gz")
        if nb is not None:
            assert dl.num_batches == nb
        assert len(dl.trajectories) > 100


def test_build_stamp_follows_the_sources(tmp_path, monkeypatch):
    b = importlib.import_module("multimodaltraj_2_b200.build")
    h0 = b.source_hash()
    assert h0 == b.source_hash() and len(h0) == 64
    if b.LIB.exists() and b.STAMP.exists() and b.STAMP.read_text().strip() == h0:
        assert not b.needs_build()
    monkeypatch.setattr(b, "NVCC_FLAGS", b.NVCC_FLAGS + ["-DX"])
    assert b.source_hash() != h0 and b.needs_build()          # any change of the inputs (here: the flags) invalidates the stamp


def test_bench_post_mortem_names_stage_and_error(tmp_path, monkeypatch, capsys):
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "ROOT", tmp_path)
    monkeypatch.setenv("RANK", "3")
    bench.stage("timed steps")
    try:
        raise RuntimeError("CUDA error: unspecified launch failure")
    except RuntimeError as e:
        bench.record_failure(e)
    text = (tmp_path / "gpurun_out" / "rank3.err").read_text()
    info = json.loads(text.splitlines()[0])
    assert info["stage"] == "timed steps" and info["rank"] == "3" and "unspecified launch failure" in info["error"]
    assert "Traceback" in text
