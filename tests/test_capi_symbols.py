"""CPU tests: libmmt.so builds for sm_100a, loads, and exports every symbol include/mmt.h declares
(no compute calls here -- there is no GPU in the build container)."""
import ctypes
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "mmt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("mmt_pairwise_adj_f32", "mmt_aggregate_f32", "mmt_edge_mlp_f32", "mmt_gsk_cell",
                 "mmt_gridlstm_step_f32", "mmt_mcr_step_f32", "mmt_decode_score_f32", "mmt_scene_batch_f32",
                 "mmt_forecast_f32", "mmt_mean_error_f32"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from multimodaltraj_2_b200 import _lib
    lib = _lib.load()
    assert lib.mmt_version() == 100
    raw = ctypes.CDLL(str(_lib.lib_path()))
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in include/mmt.h but not exported"
    assert set(_declared()) == set(_lib.SIGNATURES), "ctypes SIGNATURES and include/mmt.h differ"
    assert lib.mmt_gate_weights_packed_bytes(64, 128) == 4 * 5 * 12288
    assert lib.mmt_gate_weights_packed_bytes(32, 128) == 0


def test_argument_validation_without_a_gpu():
    from multimodaltraj_2_b200 import _lib
    lib = _lib.load()
    assert lib.mmt_pairwise_adj_f32(None, None, 1, 6, 1.0, 1.0, None, None, None, None) == -1      # N % 4 != 0
    assert b"N % 4" in lib.mmt_last_error()
    assert lib.mmt_pairwise_adj_f32(None, None, 0, 8, 1.0, 1.0, None, None, None, None) == 0       # empty batch
    assert lib.mmt_decode_score_f32(None, None, 0, 0, None, None, None, 1, 4, 40, 20, None, None, None, None, None,
                                    None, None) == -1                                              # P > 32
    # fused rollout: pointer / shape checks come before any launch
    assert lib.mmt_rollout_bf16(None, None, None, None, 1, 64, 8, 12, 4.0, 0.5, None, None, None) == -1
    assert lib.mmt_rollout_bf16(None, None, None, None, 0, 64, 8, 12, 4.0, 0.5, None, None, None) == 0   # empty batch
    cw = _lib.CellWeights()
    cw.E, cw.U = 64, 128
    import ctypes as C
    one = C.c_void_p(16)          # non-NULL, 16-byte aligned placeholder: rejected by the shape check before any use
    assert lib.mmt_rollout_bf16(one, one, one, C.byref(cw), 1, 12, 8, 12, 4.0, 0.5, one, None, None) == -1   # 128 % N != 0
    assert b"N in {8,16,32,64,128}" in lib.mmt_last_error()


def test_sass_contains_blackwell_tensor_and_bulk_copy_instructions():
    from multimodaltraj_2_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.lib_path())], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    assert "UBLKCP" in sass, "cp.async.bulk (TMA bulk copy) missing from SASS"
    assert "STTM" in sass, "tcgen05.st (operands written to tensor memory) missing from SASS"


def test_product_path_does_not_import_the_oracle():
    for f in (ROOT / "multimodaltraj_2_b200").rglob("*.py"):
        src = f.read_text()
        for ln in src.splitlines():
            if ln.strip().startswith(("import", "from")):
                assert "oracle" not in ln and "track_a" not in ln and "track_b" not in ln, (f, ln)
