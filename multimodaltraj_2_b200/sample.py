"""Mirror of ``sample.py`` (reference :21-355): ``get_mean_error`` keeps its signature and returns the
same three values (computed by ``mmt_mean_error_f32`` on the device); ``evaluate`` is the batched
replacement of ``sample.main()``'s per-batch loop: scenes -> rollout -> K-sample decode -> ADE/FDE."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def get_mean_error(predicted_traj, true_traj, observed_length, maxNumPeds):
    """sample.py:21-82.  predicted_traj / true_traj: [n, L, 2] agent-major (arrays or CUDA tensors).
    Returns (ade, fde, counter) with the reference's reductions (signed error summed over agents per
    step before the norm; only steps observed_length..L-1)."""
    def dev(a):
        if isinstance(a, torch.Tensor):
            return a.to("cuda", torch.float32).contiguous()
        return torch.as_tensor(np.asarray(a, np.float32)).cuda().contiguous()
    out = ops.mean_error(dev(predicted_traj), dev(true_traj), int(observed_length), int(maxNumPeds)).cpu().numpy()
    print('ADE = ', float(out[0]))
    print('FDE = ', float(out[1]))
    return float(out[0]), float(out[1]), int(out[2])


def evaluate(forecaster: ops.Forecaster, pos, vis, valid, eps=None):
    """Standard best-of-K ADE/FDE over all valid agents of a batch (SURVEY A.6 'standard'):
    returns dict(ade, fde, n_agents) with device-side partial sums so ranks can be combined."""
    o = forecaster(pos, vis, valid, eps)
    n = valid.sum()
    return dict(ade_sum=o["best_ade"].sum(), fde_sum=o["best_fde"].sum(), n_agents=n, out=o)
