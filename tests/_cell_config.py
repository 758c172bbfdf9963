"""Worker of tests/test_gpu_parity.py::test_blocked_cell_kernel_wide_equals_narrow: the per-step tensor-core paths
(g2k_lstm_mc stepwise, g2k_lstm_mcr, N = 256) with the blocked cell kernel in the configuration MMT_CELL_WIDE selects
(read once per process by the library); writes the outputs to argv[1]."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from multimodaltraj_2_b200 import ops, synth  # noqa: E402


def main(out_path):
    dev = torch.device("cuda")
    p = ops.CellParams.from_numpy(synth.init_params(seed=3), dev)
    res = {}
    for name, S, N, rel, prec in (("stepwise", 41, 64, False, ops.PREC_BF16_STEPWISE), ("mcr", 13, 64, True, ops.PREC_BF16),
                                  ("n256", 3, 256, False, ops.PREC_BF16), ("mcr16", 23, 16, True, ops.PREC_BF16)):
        pos, vis, valid = (torch.from_numpy(a).to(dev) for a in synth.make_crowd(S, N, seed=31 + N, half_extent=4.0 if N < 100 else 8.0,
                                                                               ragged=True))
        fc = ops.Forecaster(p, S, N, 8, 12, 20, relational=rel, prec=prec, seed=5, device=dev)
        o = fc(pos, vis, valid)
        torch.cuda.synchronize()
        res[name] = {k: o[k].cpu() for k in ("params", "best_k", "best_ade", "best_traj")}
    torch.save(res, out_path)
    print("CELL_CONFIG_OK")


if __name__ == "__main__":
    main(sys.argv[1])
