"""CPU estimate: how much of the bf16 mode's ADE/FDE error is operand rounding (and what fp16 / tf32-like operands would give).
The emulation lives in oracle/operand_rounding.py (test infrastructure); this is its command line.
    python scratch/operand_rounding_emulate.py [scenes]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import operand_rounding as o_r  # noqa: E402
import track_b as o_b  # noqa: E402
from multimodaltraj_2_b200 import synth  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    N, T, P, K = 64, 8, 12, 20
    p = synth.init_params(seed=0)
    pos, vis, valid = synth.make_crowd(S, N, seed=synth.SEED)
    eps = o_b.philox_eps(0xB200, S, N, K, P)
    want = o_b.forecast(pos, vis, valid, p, eps, T, P)
    v = valid.astype(bool)
    for name, rnd in (("none", o_r.r_none), ("bf16", o_r.r_bf16), ("tf32 (10-bit)", o_r.r_tf32), ("fp16", o_r.r_f16)):
        par = o_r.rollout(pos, vis, valid, p, rnd)
        ade, fde, *_ = o_b.decode_score(par, eps, pos[:, :, T - 1], pos[:, :, T:T + P], valid)
        print(f"{name:14s} max|dADE| {np.abs(ade[v] - want['ade'][v]).max():.3e}  max|dFDE| {np.abs(fde[v] - want['fde'][v]).max():.3e}")


if __name__ == "__main__":
    main()
