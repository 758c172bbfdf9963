"""Generates the committed golden fixtures from the read-only reference at /root/reference.

Run once in the build container (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs (all under tests/golden/):
  track_a_ckpt.npz   tensors decoded from the reference's TF-1.14 checkpoints (save/*):
                     - five (weight_c, cost, W_c@cost) triples where the TF-evaluated product saved in
                       the checkpoint equals weight_c @ cost to 1e-17 (models/g2k_lstm_mcr.py:122);
                     - the seed-0 N(0,1) weights of every Track-A shape;
                     - the saved g2k_lstm_mc ``temp_path`` tensors (all zero, models/g2k_lstm_mc.py:59-69);
                     - a GridLSTMCell parameter set (helper.py:31-39).
  zara01_slice.npz   first 2000 columns of data/ucy/zara/zara01/vis_body.csv plus what the reference's
                     own load_traj.DataLoader (imported from /root/reference, only its hard-coded
                     parent_dir patched) makes of them: the frame dictionary and three next_step() calls.
  ref_ckpt/          two of the reference's own TF-1.14 checkpoints, copied verbatim (binary data files, 107 KB):
                     save/g2k_mcr_model_val_0.ckpt-0 (17 tensors, one scalar) and
                     save/g2k_mp_model_kfold_train_2_9_12.ckpt-210 (214 tensors, 14 restart groups) --
                     golden files for the bundle writer (multimodaltraj_2_b200/tf_bundle.py must reproduce them
                     byte for byte from their decoded tensors).
"""
import shutil
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle"))
from tf_bundle import read_checkpoint  # noqa: E402


def ref_ckpt():
    (HERE / "ref_ckpt").mkdir(exist_ok=True)
    for stem in ("g2k_mcr_model_val_0.ckpt-0", "g2k_mp_model_kfold_train_2_9_12.ckpt-210"):
        for ext in (".index", ".data-00000-of-00001"):
            shutil.copyfile(REF / "save" / (stem + ext), HERE / "ref_ckpt" / (stem + ext))


def track_a():
    out = {}
    ck = read_checkpoint(REF / "save" / "g2k_mcrAttn_model_kfold_train_4_0.ckpt-79")
    n = 0
    for k in range(20):
        b = 1728 + 6 * k
        cost, wcc = ck[f"Variable_{b + 3}"], ck[f"Variable_{b + 4}"]
        for sc in range(288, 308):
            Wc = ck[f"krnl_weights_{sc}/weight_c"]
            if np.abs(Wc @ cost - wcc).max() < 1e-15:
                out[f"wc_{n}_W_c"], out[f"wc_{n}_cost"], out[f"wc_{n}_out"] = Wc, cost, wcc
                out[f"wc_{n}_ngh"], out[f"wc_{n}_attn"], out[f"wc_{n}_Eo"] = (ck[f"Variable_{b}"], ck[f"Variable_{b + 1}"],
                                                                            ck[f"Variable_{b + 2}"])
                n += 1
                break
    out["n_wc"] = np.int64(n)
    ck2 = read_checkpoint(REF / "save" / "g2k_mp_model_kfold_train_2_9_12.ckpt-210")
    for nm in ("weight_v", "bias_v", "weight_c", "cost", "attn"):
        out["seed0_" + nm] = ck2["krnl_weights/" + nm]
    out["seed0_weight_r"] = ck2["krnl_embed/weight_r"]
    out["seed0_weight_ii"] = ck2["weight_input/weight_ii"]
    out["fwd_ngh_scaled"] = ck2["Variable"]           # lambda * ngh, models/g2k_lstm_mcr.py:102
    ck3 = read_checkpoint(REF / "save" / "g2k_mcr_model_val_0.ckpt-0")
    out["seed0_weight_i_9x10"] = ck3["weight_input/weight_i"]
    out["seed0_weight_o_10x9"] = ck3["krnl_weights/weight_o"]
    zeros = [v for k, v in ck3.items() if k.startswith("Variable") and v.ndim == 2 and v.shape[0] == 16]
    out["mc_temp_path_0"], out["mc_temp_path_1"] = zeros[0], zeros[1]
    for nm in ("W_f_0_0", "B_f_0", "W_I_diag_freqf_0", "W_I_diag_freqt_0", "W_O_diag_freqf_0", "W_O_diag_freqt_0"):
        out["glstm_" + nm] = ck3["grid_lstm_cell/" + nm]
    np.savez_compressed(HERE / "track_a_ckpt.npz", **out)
    print("track_a_ckpt.npz:", n, "TF-evaluated W_c@cost triples;", len(out), "arrays")


def loader():
    csv = np.genfromtxt(REF / "data/ucy/zara/zara01/vis_body.csv", delimiter=",")[:, :2000]
    with tempfile.TemporaryDirectory() as td:
        d = Path(td) / "data/ucy/zara/zara01"
        d.mkdir(parents=True)
        np.savetxt(d / "vis_body.csv", csv, delimiter=",", fmt="%.10g")
        for other in ("eth/hotel", "eth/univ", "ucy/zara/zara02", "ucy/univ"):
            (Path(td) / "data" / other).mkdir(parents=True)
        src = (REF / "load_traj.py").read_text().replace(
            "'/home/siri0005/Documents/multimodaltraj_2/data'", repr(str(Path(td) / "data")))
        mod = types.ModuleType("ref_load_traj")
        exec(compile(src, "ref_load_traj", "exec"), mod.__dict__)
        args = types.SimpleNamespace(batch_size=16, seq_length=12, pred_len=12, obs_len=8)
        dl = mod.DataLoader(args=args, datasets=[0, 1, 2, 3, 4], sel=0, start=2)
        rows = []
        for fr, lst in dl.trajectories.items():
            for item in lst:
                (ped, pos), = item.items()
                rows.append((fr, ped, pos[0], pos[1]))
        rows = np.array(rows, np.float64)
        out = dict(csv=csv, traj_rows=rows, num_batches=np.int64(dl.num_batches), seed=np.float64(dl.seed),
                   max=np.int64(dl.max), val_max=np.int64(dl.val_max))
        # frame_preprocess leaves frame_pointer past the end; the reference's driver then sees an empty
        # batch and calls reset_data_pointer() (train.py:63-67) -- do the same.
        assert len(dl.next_step()[0]) == 0
        dl.reset_data_pointer()
        for call in range(3):
            batch, targets, fp = dl.next_step()
            out[f"batch{call}_frames"] = np.array(sorted(batch.keys()), np.float64)
            out[f"batch{call}_fp"] = np.float64(fp)
            out[f"batch{call}_n_targets"] = np.int64(len(targets))
            out[f"batch{call}_target_lens"] = np.array([len(v) for v in targets.values()], np.int64)
    np.savez_compressed(HERE / "zara01_slice.npz", **out)
    print("zara01_slice.npz:", rows.shape, "trajectory rows; batches:",
          [len(out[f"batch{c}_frames"]) for c in range(3)])


if __name__ == "__main__":
    track_a()
    loader()
    ref_ckpt()
