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


def main(argv=None, params=None, device=None, results_path=None):
    """``sample.main()`` (sample.py:85-355): evaluation of one split with per-phase wall-clock prints (:256-325),
    ``get_mean_error`` on every batch (:330) and a pickle of (observed, predicted) per batch (:347-348).
    The per-batch TF graph construction and the four ``sess.run`` phases are one batched forecast here; the phases that
    are timed are the ones that exist on this path: scene batching, forecast (rollout + decode + scoring), host copy."""
    import pickle
    import time

    from . import argParser as argsParser
    from . import realdata, synth
    np.random.seed(1)                                                        # sample.py:87
    args = argsParser.ArgsParser().parser.parse_args(argv)
    device = device or torch.device("cuda", 0)
    l = args.leaveDataset
    p = params if params is not None else ops.CellParams.from_numpy(
        synth.init_params(seed=0, E=args.embedding_size, U=args.rnn_size), device)
    prec = ops.prec_from_name(getattr(args, "precision", "fp16"))
    t0 = time.time()
    sc = realdata.scene_windows(args, l, "val", device)
    torch.cuda.synchronize()
    print('wall-clock time taken by scene batching = ', time.time() - t0)
    t1 = time.time()
    res = realdata.evaluate_split(args, l, p, part="val", prec=prec, device=device, scenes=sc)
    torch.cuda.synchronize()
    print('wall-clock time taken by forecast (rollout + decode + ADE/FDE) = ', time.time() - t1)
    t2 = time.time()
    T = args.obs_len
    pos, valid, best = sc["pos"].cpu().numpy(), sc["valid"].cpu().numpy().astype(bool), res["_out"]["best_traj"].cpu().numpy()
    print('wall-clock time taken by device -> host copy = ', time.time() - t2)
    # get_mean_error per scene on [n, obs+pred, 2] tracks (predicted = observed part + best sample), sample.py:330
    results, tot_a, tot_f, cnt = [], 0.0, 0.0, 0
    for s in range(min(pos.shape[0], getattr(args, "max_scenes", 64))):
        v = valid[s]
        if not v.any():
            continue
        true_traj = pos[s][v]
        pred_traj = np.concatenate([true_traj[:, :T], best[s][v]], 1)
        a, f, c = get_mean_error(pred_traj, true_traj, T, min(args.maxNumPeds, len(true_traj)))
        tot_a, tot_f, cnt = tot_a + a, tot_f + f, cnt + 1
        results.append((true_traj[:, :T], pred_traj))
    out = dict(realdata.public(res), mean_error_ade=tot_a / max(cnt, 1), mean_error_fde=tot_f / max(cnt, 1), scenes_scored=cnt)
    print('Total mean error of the model is ', out["mean_error_ade"])          # sample.py:343
    print('Total final error of the model is ', out["mean_error_fde"])         # sample.py:344
    if results_path:
        with open(results_path, 'wb') as f:                                    # sample.py:347-348
            pickle.dump(results, f)
    return out
