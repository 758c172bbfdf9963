#!/usr/bin/env python
"""Benchmark of the multimodaltraj forecasting hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--prec bf16|f32]
                  [--variant mc|mcr] [--scenes S] [--agents N]

One "step" = one pass of the whole hot path (obs 8 -> pred 12 rollout: pairwise kernel/adjacency ->
[edge MLP] -> aggregation -> gate update, 19 cell steps, then K=20-sample decode + ADE/FDE + best-of-K)
over one batch of synthetic ETH/UCY-shaped crowds.  Workload at N=1: BASELINE config C3,
4096 scenes x 64 agents, hidden 128.  metric = agent-trajectories/sec (whole job).
Multi-GPU: scenes shard across ranks with no data-path collective (weak scaling: 4096 scenes per rank).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

# rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on the first
# communicator of a process whenever NCCL_DEBUG=VERSION comes from the environment or from a configuration file): file
# descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved original descriptor.
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    """The one JSON line of this run, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

_STAGE = ["start"]


def stage(name):
    """Progress marker on stderr (one line per phase and rank): a failure names the phase it happened in."""
    _STAGE[0] = name
    sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')}] {name} t={time.perf_counter():.1f}\n")
    sys.stderr.flush()


def record_failure(exc):
    """Per-rank post-mortem under gpurun_out/ (and stderr): traceback, phase, libmmt's last error and trap record."""
    import traceback
    rank = os.environ.get("RANK", "0")
    info = {"rank": rank, "stage": _STAGE[0], "error": repr(exc)}
    try:
        from multimodaltraj_2_b200 import _lib
        if _lib._lib is not None:
            msg = _lib._lib.mmt_last_error()
            info["mmt_last_error"] = msg.decode() if msg else ""
            info["mmt_last_trap"] = _lib.last_trap()
    except Exception as e2:          # the post-mortem must not mask the failure
        info["postmortem_error"] = repr(e2)
    text = json.dumps(info) + "\n" + traceback.format_exc()
    sys.stderr.write(f"[bench rank {rank}] FAILED in stage '{_STAGE[0]}':\n{text}\n")
    sys.stderr.flush()
    try:
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"rank{rank}.err").write_text(text)
    except OSError:
        pass

T_OBS, P_PRED, K_SAMPLES, HIDDEN, EMBED = 8, 12, 20, 128, 64
R2, INV_2SIGMA2 = 4.0, 0.5


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons of ONE GPU sampled DURING the timed region: NVML polled every ~5 ms from a
    thread (the timed region of 20 steps lasts ~50 ms -- `nvidia-smi -lms 100` saw 0-2 samples of it); nvidia-smi
    as the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.rows, self.proc, self.h, self.stop_flag = index, uuid, [], None, None, False
        self.sm, self.reasons, self.mx, self.thread = [], set(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = (pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid) if uuid
                      else pynvml.nvmlDeviceGetHandleByIndex(index))
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:            # noqa: BLE001
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.reasons |= {n for b, n in self.BITS.items() if r & b}
            except Exception:        # noqa: BLE001
                pass
            time.sleep(0.004)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:            # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            if not self.sm:
                return None
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        if not sm:
            return None
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, str(ROOT / "oracle"))
    import track_b as o_b  # oracle: the checker of the parity sample, bench's cpu_baseline and the reference arm only
    return o_b


def cpu_forecast_sample(n_scenes, N, variant):
    """The oracle (CPU restatement of the path, numpy fp32) on a bounded sample, one process."""
    o_b = _oracle()
    from multimodaltraj_2_b200 import synth
    pos, vis, valid = synth.make_crowd(n_scenes, N, seed=synth.SEED)
    p = synth.init_params(seed=0)
    eps = o_b.philox_eps(0, n_scenes, N, K_SAMPLES, P_PRED)
    t0 = time.perf_counter()
    o_b.forecast(pos, vis, valid, p, eps, T_OBS, P_PRED, R2, INV_2SIGMA2, relational=(variant == "mcr"))
    return time.perf_counter() - t0


_W = {}


def _cpu_init(n_scenes, N, variant, seed):
    """Pool initializer: every worker process owns a fixed share of the sample, built once outside the timing."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)                      # one BLAS thread per process: the processes are the parallelism
    except Exception:
        pass
    o_b = _oracle()
    from multimodaltraj_2_b200 import synth
    pos, vis, valid = synth.make_crowd(n_scenes, N, seed=seed)
    _W.update(o_b=o_b, pos=pos, vis=vis, valid=valid, p=synth.init_params(seed=0), variant=variant,
              eps=o_b.philox_eps(0, n_scenes, N, K_SAMPLES, P_PRED))


def _cpu_step(_):
    w = _W
    w["o_b"].forecast(w["pos"], w["vis"], w["valid"], w["p"], w["eps"], T_OBS, P_PRED, R2, INV_2SIGMA2,
                      relational=(w["variant"] == "mcr"))
    return 0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow 1.14 cannot be installed
    offline, so this is the oracle port of its algebra on all host cores (kind 'port'): one process
    per core, each forecasting its own `per` scenes per step; process start-up and input generation are outside
    the timed region."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per = 8                                        # scenes per process per step (~40 ms of numpy each)
    N = args.agents
    n_scenes = per * cores
    steps, warmup = max(1, args.steps), max(0, args.warmup)   # as given: a step is ~0.1 s of wall time (8 scenes per core)
    with mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(per, N, args.variant, 7)) as pool:
        for _ in range(warmup):
            pool.map(_cpu_step, range(cores), chunksize=1)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            pool.map(_cpu_step, range(cores), chunksize=1)
            ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts))
    val = n_scenes * N / t
    line = {"impl": "reference", "metric": "agent-trajectories/sec (obs8->pred12, K=20)", "value": val,
            "unit": "agent-trajectories/s", "n_gpus": world, "steps": len(ts), "warmup": warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config(args, n_scenes, reference=True),
            "cpu_baseline": {"value": val, "unit": "agent-trajectories/s", "cores": cores, "kind": "port",
                             "sample": f"{n_scenes} scenes x {N} agents per step ({per} scenes per process), numpy fp32 "
                                       f"oracle of the whole path on {cores} processes, 1 BLAS thread each "
                                       f"(TF 1.14 reference not installable offline)"},
            "e2e": {"value": val, "unit": "agent-trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def config(args, scenes_per_rank, reference=False):
    return {"workload": f"C3 synthetic crowds: {scenes_per_rank} scenes x {args.agents} agents per GPU, hidden {HIDDEN}, "
                        f"obs {T_OBS} / pred {P_PRED}, K={K_SAMPLES}, g2k_lstm_{args.variant} batched inference",
            "variant": f"g2k_lstm_{args.variant}", "precision_mode": args.prec, "scenes_per_gpu": scenes_per_rank,
            "agents_per_scene": args.agents, "noise": "in-kernel Philox4x32-10",
            **({"precision_mode": "f32 (numpy oracle on host cores)", "noise": "Philox4x32-10 (oracle restatement)",
                "sample_of": "C3: 4096 scenes x 64 agents per GPU"} if reference else {}),
            "l2": "inputs larger than L2: 4 device copies of the batch (59 MB each) are cycled and every step writes ~150 MB of outputs/workspace; no explicit flush"}


# ------------------------------------------------------------------------------------------------
_PARITY_CACHE = {}


def parity_sample(ops, synth, params, dev, prec, variant, n_scenes, N):
    """ADE/FDE (best of K) and predicted positions of the benched mode against the CPU oracle on a sample of the
    benched workload (the first `n_scenes` scenes of rank 0's C3 batch).  The K-sample noise is FED to both sides
    (the oracle's Philox restatement), so the deltas measure the arithmetic of the rollout, not the generator."""
    import torch
    key = (n_scenes, N, variant)
    if key not in _PARITY_CACHE:          # the oracle's answer is computed once and shared by the modes
        o_b = _oracle()
        pos, vis, valid = synth.make_crowd(n_scenes, N, seed=synth.SEED)
        eps = o_b.philox_eps(0xB200, n_scenes, N, K_SAMPLES, P_PRED)
        # forecast() = rollout + decode; the rollout is run with its trace for the adjacency masks of the predicted steps
        ro = o_b.rollout(pos, vis, valid, synth.init_params(seed=0), T_OBS, P_PRED, R2, INV_2SIGMA2,
                         relational=(variant == "mcr"), trace=True)
        a_, f_, b_, bt_, _ = o_b.decode_score(ro["params"], eps, pos[:, :, T_OBS - 1], pos[:, :, T_OBS:T_OBS + P_PRED], valid)
        adj_pred = np.stack(ro.pop("trace")["adj"][T_OBS:])          # [P-1, S, N, N]: steps T .. T+P-2 see predicted positions
        _PARITY_CACHE[key] = (pos, vis, valid, eps, dict(ade=a_, fde=f_, best_k=b_, best_traj=bt_, adj_pred=adj_pred, **ro))
    pos, vis, valid, eps, want = _PARITY_CACHE[key]
    fc = ops.Forecaster(params, n_scenes, N, T_OBS, P_PRED, K_SAMPLES, R2, INV_2SIGMA2, relational=(variant == "mcr"),
                        prec=prec, device=dev, want_all=True)
    o = fc(*(torch.from_numpy(a).to(dev) for a in (pos, vis, valid)), eps=torch.from_numpy(eps).to(dev))
    torch.cuda.synchronize()
    g = {k: o[k].cpu().numpy() for k in ("ade", "fde", "best_k", "params")}
    v = valid.astype(bool)
    idx = np.arange(K_SAMPLES)[None, None, :]
    w_best = want["best_k"][..., None] == idx
    g_best = g["best_k"][..., None] == idx
    w_ade, w_fde = (want["ade"] * w_best).sum(-1)[v], (want["fde"] * w_best).sum(-1)[v]
    g_ade, g_fde = (g["ade"] * g_best).sum(-1)[v], (g["fde"] * g_best).sum(-1)[v]
    # mean predicted trajectory (cumulative mu) of the rollout, relative to the largest displacement
    w_mu, g_mu = want["params"][..., :2].cumsum(2)[v], g["params"][..., :2].cumsum(2)[v]
    # Which scenes saw a DIFFERENT neighbour set than the oracle?  The kernel's adjacency test (bit-exact arithmetic: tests)
    # re-evaluated on ITS predicted positions (last observed point + running fp32 sum of the emitted means, the kernel's
    # own order of additions) against the oracle's masks.  One flipped neighbour changes a softmax row by O(0.1): the
    # recurrence is discontinuous there, and that scene's deviation says nothing about the arithmetic before the flip.
    o_b = _oracle()
    cur = pos[:, :, T_OBS - 1].astype(np.float32)
    flip = np.zeros(n_scenes, bool)
    for k in range(P_PRED - 1):
        cur = (cur + g["params"][:, :, k, :2]).astype(np.float32)
        flip |= (o_b.pairwise_adj(cur, valid, R2, INV_2SIGMA2)[1] != want["adj_pred"][k]).any((1, 2))
    keep = v & ~flip[:, None]
    d_ade, d_fde = np.abs(g["ade"][v] - want["ade"][v]), np.abs(g["fde"][v] - want["fde"][v])      # [agents, K]
    nf_ade = float(np.abs(g["ade"][keep] - want["ade"][keep]).max()) if keep.any() else 0.0
    nf_fde = float(np.abs(g["fde"][keep] - want["fde"][keep]).max()) if keep.any() else 0.0
    over = (np.abs(g["ade"] - want["ade"]).max(-1) > 1e-3) | (np.abs(g["fde"] - want["fde"]).max(-1) > 1e-3)   # [S, N]
    over &= v
    return {"sample": f"{n_scenes} scenes x {N} agents of the benched batch, fed noise",
            "max_abs_d_ade": float(d_ade.max()),
            "max_abs_d_fde": float(d_fde.max()),
            # the distribution behind the maxima: the adjacency test d^2 < r^2 makes the recurrence discontinuous, so in a
            # large enough sample any arithmetic that is not bit-identical flips a neighbour somewhere
            "quantiles_abs_d_ade": {q: float(np.quantile(d_ade, float(q))) for q in ("0.5", "0.99", "0.9999")},
            "quantiles_abs_d_fde": {q: float(np.quantile(d_fde, float(q))) for q in ("0.5", "0.99", "0.9999")},
            "frac_samples_within_1e-3": float(((d_ade <= 1e-3) & (d_fde <= 1e-3)).mean()),
            "agents_over_1e-3": int(over.sum()), "scenes_with_an_agent_over_1e-3": int(over.any(1).sum()),
            "scenes_with_a_flipped_neighbour": int(flip.sum()),
            "scenes_over_1e-3_without_a_flip": int((over.any(1) & ~flip).sum()),
            "max_abs_d_ade_where_adjacency_agrees": nf_ade, "max_abs_d_fde_where_adjacency_agrees": nf_fde,
            "within_1e-3_where_adjacency_agrees": bool(nf_ade <= 1e-3 and nf_fde <= 1e-3),
            "d_mean_best_ade": float(abs(g_ade.mean() - w_ade.mean())), "d_mean_best_fde": float(abs(g_fde.mean() - w_fde.mean())),
            "oracle_mean_best_ade": float(w_ade.mean()), "oracle_mean_best_fde": float(w_fde.mean()),
            "best_k_equal_frac": float((g["best_k"][v] == want["best_k"][v]).mean()),
            "pos_rel_err": float(np.abs(g_mu - w_mu).max() / max(np.abs(w_mu).max(), 1e-30)),
            "within_1e-3": bool(np.abs(g["ade"][v] - want["ade"][v]).max() <= 1e-3 and np.abs(g["fde"][v] - want["fde"][v]).max() <= 1e-3)}


def run_ours(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from multimodaltraj_2_b200 import _lib, ops, synth

    S, N = args.scenes, args.agents
    PRECS = {"bf16": ops.PREC_BF16, "f32": ops.PREC_F32, "bf16-stepwise": ops.PREC_BF16_STEPWISE}
    if hasattr(ops, "PREC_BF16X3"):
        PRECS["bf16x3"] = ops.PREC_BF16X3
    PRECS["f16"] = ops.PREC_F16
    prec = PRECS[args.prec]
    relational = args.variant == "mcr"
    pos_h, vis_h, valid_h = synth.make_crowd(S, N, seed=synth.SEED + rank)
    params = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
    fc = ops.Forecaster(params, S, N, T_OBS, P_PRED, K_SAMPLES, R2, INV_2SIGMA2, relational=relational,
                        prec=prec, seed=0xB200, agent_offset=rank * S * N, device=dev, use_graph=not args.no_graph)
    pos_p, vis_p, valid_p = (torch.from_numpy(a).pin_memory() for a in (pos_h, vis_h, valid_h))
    pos, vis, valid = pos_p.to(dev), vis_p.to(dev), valid_p.to(dev)
    # NSETS device copies of the inputs are cycled so that a step never finds its inputs in the 126 MB L2
    # (one set = pos 42 MB + vis 17 MB; outputs and workspace add ~150 MB per step)
    NSETS = 4
    sets = [(pos, vis, valid)] + [(pos.clone(), vis.clone(), valid.clone()) for _ in range(NSETS - 1)]
    n_valid = int(valid_h.sum())
    units = n_valid * world          # every rank holds the same number of valid agents (all valid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_launch = [0]

    def timed_steps(f, n_warm, n_steps, sampler=None):
        """W untimed + K timed steps of forecaster f on this rank, CUDA events on the launching stream,
        barrier + synchronize on both sides, MAX over ranks.  Returns ms per step."""
        k = [0]

        def one():
            st = sets[k[0] % NSETS]
            k[0] += 1
            return f(*st)
        for st in sets:                  # first call per input set: eager validation + CUDA-graph capture (untimed)
            f(*st)
        for _ in range(n_warm):
            one()
        barrier()
        if sampler is not None:
            sampler.start()
        l0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            one()
        e1.record()
        barrier()
        n_launch[0] = ops.launch_count() - l0
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n_steps

    stage("warmup + timed steps")
    uuid = getattr(torch.cuda.get_device_properties(dev), "uuid", None)
    sampler = ClockSampler(local_rank, f"GPU-{uuid}" if uuid is not None else None) if rank == 0 else None
    ms_step = timed_steps(fc, args.warmup, args.steps, sampler)
    launches = n_launch[0]           # this rank's kernels launched between the two events (graph replays counted per node)
    clocks = sampler.stop() if sampler is not None else None
    value = units / (ms_step * 1e-3)
    stage("timed steps done")

    # ---- the other precision modes of the same workload, a few steps each, in the same line ("modes"):
    # f32 = the parity mode of the north_star tolerance (CUDA-core FMA); bf16x3 = split-bf16 tensor-core mode
    modes = {args.prec: {"value": value, "ms_per_step": ms_step}}
    if args.modes and not relational and args.prec in ("bf16", "f16"):
        for name in [m for m in ("f16", "bf16", "bf16x3", "f32") if m in PRECS and m != args.prec]:
            stage(f"mode {name}")
            fm = ops.Forecaster(params, S, N, T_OBS, P_PRED, K_SAMPLES, R2, INV_2SIGMA2, prec=PRECS[name], seed=0xB200,
                                agent_offset=rank * S * N, device=dev, use_graph=not args.no_graph)
            ms_m = timed_steps(fm, 3, 5 if name == "f32" else 10)
            modes[name] = {"value": units / (ms_m * 1e-3), "ms_per_step": ms_m}
            del fm

    # ---- e2e: host buffers in, forecast out, copies inside the timed region.  A two-deep pipeline as a serving loop
    # would run it: the H2D copy of step i+1 (copy stream, pinned host memory) and the D2H copy of step i-1's
    # forecast (best trajectory of every agent + its ADE / FDE, second copy stream, pinned host memory) overlap the
    # kernels of step i; every step uploads its own inputs and the host holds every step's forecast.
    stage("e2e warmup")
    o0 = fc.out
    res_h = [{k: torch.empty(o0[k].shape, dtype=o0[k].dtype).pin_memory() for k in ("best_traj", "best_ade", "best_fde")}
             for _ in range(2)]
    stage_d = [{k: torch.empty_like(o0[k]) for k in ("best_traj", "best_ade", "best_fde")} for _ in range(2)]
    bufs = [(torch.empty_like(pos), torch.empty_like(vis), torch.empty_like(valid)) for _ in range(2)]
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_staged = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n):
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(h2d_stream):
                if i >= 2:
                    h2d_stream.wait_event(ev_free[b])           # the kernels of step i-2 have consumed this buffer
                bufs[b][0].copy_(pos_p, non_blocking=True)
                bufs[b][1].copy_(vis_p, non_blocking=True)
                bufs[b][2].copy_(valid_p, non_blocking=True)
                ev_copied[b].record(h2d_stream)
            main_stream.wait_event(ev_copied[b])
            if i >= 2:
                main_stream.wait_event(ev_done[b])              # staging buffer b has left the device
            o = fc(*bufs[b])
            ev_free[b].record(main_stream)
            for k, t in stage_d[b].items():                     # device-side hand-over (25 MB at HBM speed), so that the
                t.copy_(o[k], non_blocking=True)                # next step may overwrite the forecaster's outputs
            ev_staged[b].record(main_stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ev_staged[b])
                for k, t in stage_d[b].items():
                    res_h[b][k].copy_(t, non_blocking=True)
                ev_done[b].record(d2h_stream)
            if i >= 1:
                ev_done[(i - 1) & 1].synchronize()              # the host consumes the previous step's forecast
        ev_done[(n - 1) & 1].synchronize()
        return res_h[(n - 1) & 1]

    e2e_run(3)
    barrier()
    stage("e2e timed")
    n_e2e = max(4, min(args.steps, 50))       # long enough that the fill of the two-deep pipeline (one exposed H2D copy) is amortised
    t0 = time.perf_counter()
    last = e2e_run(n_e2e)
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = units / float(te.item())
    h2d = pos_p.numel() * 4 + vis_p.numel() * 4 + valid_p.numel()
    d2h = sum(t.numel() * t.element_size() for t in res_h[0].values())
    ade, fde = float(last["best_ade"].sum() / n_valid), float(last["best_fde"].sum() / n_valid)

    stage("roofline kernels")
    # ---- roofline of the dominant kernel, timed alone with CUDA events on the launching stream
    R = S * N
    pk = peaks()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    nsteps = T_OBS + P_PRED - 1
    fused = prec in (ops.PREC_BF16, ops.PREC_F16) and args.variant == "mc" and 128 % N == 0 and N >= 8
    f16 = prec == ops.PREC_F16
    prof = {}
    for name in ("r02_traffic.json", "r01_traffic.json"):
        pf = ROOT / "profiles" / name
        if pf.exists():
            prof = json.loads(pf.read_text())
            break
    if fused:
        # rollout_tc_kernel: the whole T+P-1 step recurrence (pairwise + softmax -> aggregation MMA -> gate MMA ->
        # gate update -> head) in one launch.  Algorithmic FLOPs per agent-step: gate GEMM 2*320*384 + aggregation
        # of h and c over the scene 2*N*2U.
        par_out = torch.empty((S, N, P_PRED, 5), device=dev)
        for _ in range(3):
            ops.rollout_bf16(pos, vis, valid, params, T_OBS, P_PRED, R2, INV_2SIGMA2, out=par_out, f16=f16)
        torch.cuda.synchronize()
        k0.record()
        for i in range(reps):
            p_, v_, m_ = sets[i % NSETS]
            ops.rollout_bf16(p_, v_, m_, params, T_OBS, P_PRED, R2, INV_2SIGMA2, out=par_out, f16=f16)
        k1.record()
        torch.cuda.synchronize()
        ro_ms = k0.elapsed_time(k1) / reps
        flops = float(R) * nsteps * (2.0 * (EMBED + 2 * HIDDEN) * 3 * HIDDEN + 2.0 * N * 2 * HIDDEN)
        achieved = flops / (ro_ms * 1e-3) / 1e12
        peak = pk["bf16"]
        roof = {"kernel": "rollout_tc_kernel" + ("<f16 operands>" if f16 else ""), "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": prof.get("rollout_tc_kernel_dram_bytes_per_launch"),
                "peak_source": f"{pk['src']} (burst bf16 cuBLAS: kernel timed alone); sustained peak {pk['bf16_sustained']}",
                "frac_of_sustained": achieved / pk["bf16_sustained"], "ms_per_launch": ro_ms,
                "algorithmic_flops_per_launch": flops, "launches_per_step": 1,
                "algorithmic_hbm_bytes_per_launch": float(R) * (T_OBS * 16 + P_PRED * 20 + 1)}
    else:
        x = torch.randn((R, 4), device=dev) * 0.3
        h, c, mh, mc = (torch.randn((R, HIDDEN), device=dev) * 0.5 for _ in range(4))
        vflat = valid.reshape(-1).contiguous()
        cur = torch.randn((R, 2), device=dev)
        cprec = prec if prec in (ops.PREC_F32, getattr(ops, "PREC_BF16X3", -1)) else ops.PREC_BF16
        for _ in range(3):
            ops.gsk_cell(x, h, c, mh, mc, vflat, params, cprec, cur_pos=cur, want_head=True)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(reps):
            ops.gsk_cell(x, h, c, mh, mc, vflat, params, cprec, cur_pos=cur, want_head=True)
        k1.record()
        torch.cuda.synchronize()
        cell_ms = k0.elapsed_time(k1) / reps
        flops = 2.0 * R * (EMBED + 2 * HIDDEN) * 3 * HIDDEN
        achieved = flops / (cell_ms * 1e-3) / 1e12
        peak = pk["bf16_sustained"]
        x3 = cprec == getattr(ops, "PREC_BF16X3", -1)
        roof = {"kernel": ("gsk_cell_tc_kernel<x3>" if x3 else "gsk_cell_tc_kernel") if prec != ops.PREC_F32 else "gsk_cell_f32_kernel", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": f"{pk['src']} (sustained bf16 cuBLAS)", "ms_per_launch": cell_ms,
                "algorithmic_flops_per_launch": flops, "launches_per_step": nsteps,
                **({"note": "split bf16: three tensor-core MMAs per algorithmic product; the per-step kernel moves the fp32 state "
                            "through HBM (3.6 KB per agent-step) and is bound by its epilogue, not by the tensor pipe"} if x3 else {})}
    # decode + ADE/FDE epilogue kernel alone (Philox mode): 500 algorithmic bytes per agent
    o_dec = fc.out
    lo = pos[:, :, T_OBS - 1].contiguous()
    gt = pos[:, :, T_OBS:].contiguous()
    for _ in range(2):
        ops.decode_score(o_dec["params"], lo, gt, valid, K_SAMPLES, seed=1, want_all=False)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(reps):
        ops.decode_score(o_dec["params"], lo, gt, valid, K_SAMPLES, seed=1, want_all=False)
    k1.record()
    torch.cuda.synchronize()
    dec_ms = k0.elapsed_time(k1) / reps
    roof_dec = {"kernel": "decode_score_kernel", "bound": "hbm", "achieved": R * 500.0 / (dec_ms * 1e-3) / 1e9,
                "peak": pk["hbm"], "unit": "GB/s", "frac": R * 500.0 / (dec_ms * 1e-3) / 1e9 / pk["hbm"],
                "traffic": prof.get("decode_score_kernel_dram_bytes_per_launch"), "ms_per_launch": dec_ms,
                "note": "Philox + Box-Muller in-kernel: ALU-bound in practice (K*P = 240 normal pairs per agent)"}
    # pairwise kernel: HBM roofline (5N^2 + 9N bytes per scene-frame)
    # all T observed frames of the batch in one launch: S*T scene-frames (output 5 N^2 S T bytes >> L2)
    fc_pos = pos[:, :, :T_OBS].permute(0, 2, 1, 3).reshape(S * T_OBS, N, 2).contiguous()
    fc_valid = valid[:, None, :].expand(S, T_OBS, N).reshape(S * T_OBS, N).contiguous()
    pw_out = (torch.empty((S * T_OBS, N, N), device=dev), torch.empty((S * T_OBS, N, N), dtype=torch.uint8, device=dev))
    lib = _lib.load()
    import ctypes as C

    def pw():
        _lib.check(lib.mmt_pairwise_adj_f32(C.c_void_p(fc_pos.data_ptr()), C.c_void_p(fc_valid.data_ptr()), S * T_OBS, N,
                                            R2, INV_2SIGMA2, C.c_void_p(pw_out[0].data_ptr()),
                                            C.c_void_p(pw_out[1].data_ptr()), None,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for _ in range(3):
        pw()
    torch.cuda.synchronize()
    k0.record()
    for _ in range(reps):
        pw()
    k1.record()
    torch.cuda.synchronize()
    pw_ms = k0.elapsed_time(k1) / reps
    pw_bytes = S * T_OBS * (5.0 * N * N + 9.0 * N)
    roof_pw = {"kernel": "pairwise_adj_kernel", "bound": "hbm", "achieved": pw_bytes / (pw_ms * 1e-3) / 1e9,
               "peak": pk["hbm"], "unit": "GB/s", "frac": pw_bytes / (pw_ms * 1e-3) / 1e9 / pk["hbm"],
               "traffic": None, "ms_per_launch": pw_ms,
               "workload": f"{S * T_OBS} scene-frames x {N} agents (all observed frames of the batch), 690 MB written"}

    torch.cuda.synchronize()
    stage("gpu phases done")
    if rank == 0:
        # ---- parity of the benched arithmetic where the headline lives: deltas against the CPU oracle on a sample of C3
        stage("parity sample vs oracle")
        n_par = min(S, args.parity_scenes)
        delta = {args.prec: parity_sample(ops, synth, params, dev, prec, args.variant, n_par, N)}
        for name in modes:
            if name not in delta:
                delta[name] = parity_sample(ops, synth, params, dev, PRECS[name], args.variant, n_par, N)
        for name, m in modes.items():
            m["max_abs_d_ade_vs_oracle"] = delta[name]["max_abs_d_ade"]
            m["max_abs_d_fde_vs_oracle"] = delta[name]["max_abs_d_fde"]
            m["pos_rel_err_vs_oracle"] = delta[name]["pos_rel_err"]
            m["within_1e-3"] = delta[name]["within_1e-3"]
            m["frac_samples_within_1e-3"] = delta[name]["frac_samples_within_1e-3"]
            m["p9999_abs_d_fde_vs_oracle"] = delta[name]["quantiles_abs_d_fde"]["0.9999"]
            m["scenes_with_an_agent_over_1e-3"] = delta[name]["scenes_with_an_agent_over_1e-3"]
            m["scenes_with_a_flipped_neighbour"] = delta[name]["scenes_with_a_flipped_neighbour"]
            m["within_1e-3_where_adjacency_agrees"] = delta[name]["within_1e-3_where_adjacency_agrees"]
            m["max_abs_d_ade_where_adjacency_agrees"] = delta[name]["max_abs_d_ade_where_adjacency_agrees"]
            m["max_abs_d_fde_where_adjacency_agrees"] = delta[name]["max_abs_d_fde_where_adjacency_agrees"]
        stage("cpu baseline")
        cores = os.cpu_count() or 1
        # CPU baseline on a bounded sample of the same workload: a small probe sizes it for ~10 s of CPU work
        cpu_forecast_sample(8, N, args.variant)                      # warm-up (imports, allocator)
        n_cpu = 64
        t_cpu = cpu_forecast_sample(n_cpu, N, args.variant)          # probe: scenes per second
        n_big = max(64, min(4096 * 64 // N, int(10.0 * n_cpu / max(t_cpu, 1e-3))))
        if n_big > n_cpu:
            n_cpu = n_big
            t_cpu = cpu_forecast_sample(n_cpu, N, args.variant)
        line = {"metric": "agent-trajectories/sec (obs8->pred12, K=20)", "value": value, "unit": "agent-trajectories/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"f32": "f32", "f16": "f16 (tensor-core operands; fp32 accumulation and state)"}.get(args.prec, "bf16"),
                "data": "synthetic", "config": config(args, S),
                "modes": modes,
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "agent-trajectories/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h,
                        "d2h": "best trajectory [S,N,P,2] + best-of-K ADE and FDE [S,N] of every agent, every step"},
                "gpu_launches": int(launches), "roofline": roof, "roofline_pairwise": roof_pw, "roofline_decode": roof_dec,
                "cpu_baseline": {"value": n_cpu * N / t_cpu, "unit": "agent-trajectories/s", "cores": 1, "kind": "port",
                                 "sample": f"{n_cpu} scenes x {N} agents ({t_cpu:.1f} s of CPU work), numpy fp32 oracle of the whole path "
                                           f"(host has {cores} cores; TF 1.14 reference not installable offline)"},
                "ade_fde": {"best_of_k_ade": ade, "best_of_k_fde": fde, "note": "random-init weights, synthetic data",
                            "delta_vs_oracle": delta[args.prec],
                            "tolerance": "north_star: ADE/FDE within 1e-3, positions 1e-4 relative in fp32; reduced-precision operand modes (f16, bf16) stated separately in modes{}"},
                "lib": str(_lib.lib_path().relative_to(ROOT))}
        emit(line)
    stage("teardown")
    if world > 1:
        dist.barrier()               # rank 0 spent ~20 s in the CPU baseline: leave together
        dist.destroy_process_group()
    stage("done")


def run_config(args, rank, world, local_rank):
    """--config c1 | c5: the real-data configurations of BASELINE.json on the ETH / UCY tables shipped under data/.
      c1  g2k_lstm_mc forward + ADE/FDE on the ETH-univ split (configs[0]),
      c5  best-of-20 ADE/FDE over all five ETH/UCY splits, scenes sharded over the ranks (configs[4]).
    Per split: device-side scene batching of every obs 8 + pred 12 window -> forecaster on this rank's shard -> ONE
    all-reduce of (sum ADE, sum FDE, agents[, metres]).  Beside every score: the CPU oracle's value on the first
    --parity-scenes windows with the same fed noise, and the delta.  Weights are the seed-0 random init (the reference
    ships no trained model: SURVEY F2/F3), so the scores are parity evidence, not accuracy."""
    import types

    import torch
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from multimodaltraj_2_b200 import ops, realdata, synth
    a = types.SimpleNamespace(batch_size=16, seq_length=12, pred_len=P_PRED, obs_len=T_OBS, K=K_SAMPLES, data_root=None)
    prec = {"bf16": ops.PREC_BF16, "f32": ops.PREC_F32, "bf16-stepwise": ops.PREC_BF16_STEPWISE, "f16": ops.PREC_F16,
            "bf16x3": ops.PREC_BF16X3}[args.prec]
    relational = args.variant == "mcr"
    params = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
    splits = [1] if args.config == "c1" else [0, 1, 2, 3, 4]
    rows, total_agents, total_ms = {}, 0, 0.0
    for d in splits:
        stage(f"split {d}")
        t0 = time.perf_counter()
        sc = realdata.scene_windows(a, d, "all", dev)
        torch.cuda.synchronize()
        t_batch = time.perf_counter() - t0
        res = realdata.evaluate_split(a, d, params, part="all", prec=prec, relational=relational, rank=rank, world=world,
                                      device=dev, scenes=sc)                      # warm-up + the reported scores
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 3
        for _ in range(reps):
            realdata.evaluate_split(a, d, params, part="all", prec=prec, relational=relational, rank=rank, world=world,
                                    device=dev, scenes=sc)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        row = realdata.public(res)
        row.update(ms_forecast=float(t.item()), s_scene_batching=t_batch)
        if rank == 0:
            o_b = _oracle()
            n = min(args.parity_scenes, sc["pos"].shape[0])
            pos, vis, valid = (sc[k][:n].cpu().numpy() for k in ("pos", "vis", "valid"))
            eps = o_b.philox_eps(d, n, sc["N"], K_SAMPLES, P_PRED)
            want = o_b.forecast(pos, vis, valid, synth.init_params(seed=0), eps, T_OBS, P_PRED, R2, INV_2SIGMA2,
                                relational=relational)
            v = valid.astype(bool)
            pick = want["best_k"][..., None] == np.arange(K_SAMPLES)[None, None, :]
            w_ade, w_fde = (want["ade"] * pick).sum(-1)[v], (want["fde"] * pick).sum(-1)[v]
            sub = {k: sc[k][:n].contiguous() for k in ("pos", "vis", "valid")}
            sub["N"] = sc["N"]
            got = realdata.evaluate_split(a, d, params, part="all", prec=prec, relational=relational, device=dev, scenes=sub,
                                          eps=torch.from_numpy(eps).to(dev))
            g_ade = got["_out"]["best_ade"].cpu().numpy()[v]
            g_fde = got["_out"]["best_fde"].cpu().numpy()[v]
            row["oracle"] = {"sample": f"first {n} windows, fed noise", "ade": float(w_ade.mean()), "fde": float(w_fde.mean()),
                             "ours_ade": float(g_ade.mean()), "ours_fde": float(g_fde.mean()),
                             "max_abs_d_ade": float(np.abs(g_ade - w_ade).max()), "max_abs_d_fde": float(np.abs(g_fde - w_fde).max()),
                             "best_k_equal_frac": float((got["_out"]["best_k"].cpu().numpy()[v] == want["best_k"][v]).mean())}
            same = got["_out"]["best_k"].cpu().numpy()[v] == want["best_k"][v]       # a flipped near-tie compares two different samples' FDE
            row["oracle"]["max_abs_d_fde_same_best_k"] = float(np.abs(g_fde - w_fde)[same].max()) if same.any() else None
        rows[realdata.DATASET_NAMES[d]] = row
        total_agents += row["n_agents"]
        total_ms += row["ms_forecast"]
    if rank == 0:
        emit({"metric": "agent-trajectories/sec (obs8->pred12, K=20)", "value": total_agents / (total_ms * 1e-3),
              "unit": "agent-trajectories/s", "n_gpus": world, "steps": 3, "warmup": 1, "ms_per_step": total_ms,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": {"f32": "f32", "f16": "f16"}.get(args.prec, "bf16"), "data": "ETH/UCY tables under data/ (real), seed-0 random-init weights",
              "config": {"workload": {"c1": "C1: g2k_lstm forward + best-of-20 ADE/FDE on the ETH-univ split",
                                      "c5": "C5: best-of-20 ADE/FDE over all five ETH/UCY splits, scene-sharded"}[args.config],
                         "variant": f"g2k_lstm_{args.variant}", "precision_mode": args.prec,
                         "units": "ADE/FDE in the tables' units (normalised pixels / z-scores); *_m in metres (ETH homographies)"},
              "splits": rows})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c2(args, rank, world, local_rank):
    """--config c2 (BASELINE configs[1]): g2k_lstm_mcr training steps, UCY zara1 leave-one-out.  The reference's loop
    (train.py:28-41) trains on datasets {2,3,4,5} minus the left-out one; 5 (town_center.csv) is absent upstream, so zara02
    and ucy/univ train and zara01 is the held-out split.  Every step is one data-parallel step of the Trainer on ALL obs+pred
    windows of one training table's first 70 % of the columns (load_traj.py:125-134), scenes sharded over the ranks, ONE
    gradient all-reduce; the held-out best-of-20 ADE / FDE is evaluated before and after (the only accuracy datum this repo can
    produce: the reference ships no trained model)."""
    import types

    import torch
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from multimodaltraj_2_b200 import ops, realdata, synth
    from multimodaltraj_2_b200.train import Trainer
    a = types.SimpleNamespace(batch_size=16, seq_length=12, pred_len=P_PRED, obs_len=T_OBS, K=K_SAMPLES, data_root=None)
    relational = args.variant != "mc"            # configs[1] names g2k_lstm_mcr: the default of this config
    params = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
    leave, train_sets = 2, [3, 4]
    tr = Trainer(params, T_OBS, P_PRED, R2, INV_2SIGMA2, lr=1e-3, gemm=args.train_gemm, relational=relational,
                 graph=args.train_graph)
    shards = {}
    for d in train_sets:
        sc = realdata.scene_windows(a, d, "train", dev)
        lo, hi = realdata.shard_range(sc["pos"].shape[0], rank, world)
        shards[d] = tuple(sc[k][lo:hi].contiguous() for k in ("pos", "vis", "valid")) + (int(sc["valid"].sum()), sc["N"], sc["pos"].shape[0])
    held = realdata.scene_windows(a, leave, "all", dev)

    def evaluate():
        r = realdata.evaluate_split(a, leave, params, part="all", prec=ops.PREC_F16, relational=relational, rank=rank,
                                    world=world, device=dev, scenes=held, seed=11)
        return realdata.public(r)
    before = evaluate()
    losses, ms = [], {d: [] for d in train_sets}
    steps = max(1, args.steps)
    for it in range(args.warmup + steps):
        for d in train_sets:
            pos, vis, valid, n_valid, N, S = shards[d]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = tr.step(pos, vis, valid)
            e1.record()
            torch.cuda.synchronize()
            if it >= args.warmup:
                t = torch.tensor([e0.elapsed_time(e1)], device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms[d].append(float(t.item()))
            losses.append(float(loss))
    after = evaluate()
    if rank == 0:
        per = {realdata.DATASET_NAMES[d]: {"scenes": shards[d][5], "agents_per_scene": shards[d][4], "valid_agents": shards[d][3],
                                           "ms_per_step": float(np.mean(ms[d])),
                                           "agent_trajectories_per_s": shards[d][3] / (float(np.mean(ms[d])) * 1e-3)} for d in train_sets}
        tot_agents = sum(shards[d][3] for d in train_sets)
        tot_ms = sum(float(np.mean(ms[d])) for d in train_sets)
        emit({"mode": "train", "metric": "agent-trajectories/sec (training step: teacher-forced NLL + BPTT + gradient all-reduce + RMSProp)",
              "value": tot_agents / (tot_ms * 1e-3), "unit": "agent-trajectories/s", "n_gpus": world, "steps": steps,
              "warmup": args.warmup, "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "f32 state; contractions " + args.train_gemm, "data": "ETH/UCY tables under data/ (real)",
              "config": {"workload": "C2: g2k_lstm_%s training step, UCY zara1 leave-one-out (train: zara02 + ucy/univ, all obs+pred windows "
                                     "of the training columns per step; held out: zara01)" % ("mcr" if relational else "mc"),
                         "contractions": args.train_gemm, "cuda_graph": bool(args.train_graph), "lr": 1e-3, "collective": "one all-reduce (SUM) of the gradient bucket per step" if world > 1 else "none (1 GPU)"},
              "per_table": per, "loss_first": losses[0], "loss_last": losses[-1],
              "held_out_zara01_best_of_20": {"before": before, "after": after,
                                             "note": "normalised (z-scored) units; random init -> after %d steps per table" % (args.warmup + steps)}})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args, rank, world, local_rank):
    """--mode train: data-parallel training steps (teacher-forced NLL, BPTT through the fp32 kernels, ONE NCCL
    all-reduce of the flat gradient bucket, RMSProp) on this rank's scene shard.  Extra mode: the headline metric of
    BASELINE.json is the inference line printed by the default mode."""
    import torch
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from multimodaltraj_2_b200 import ops, synth
    from multimodaltraj_2_b200.train import Trainer, flatten_bucket
    S, N = args.scenes, args.agents
    pos_h, vis_h, valid_h = synth.make_crowd(S, N, seed=synth.SEED + rank)
    pos, vis, valid = (torch.from_numpy(a).to(dev) for a in (pos_h, vis_h, valid_h))
    params = ops.CellParams.from_numpy(synth.init_params(seed=0), dev)
    relational = getattr(args, "variant", "mc") == "mcr"
    tr = Trainer(params, T_OBS, P_PRED, R2, INV_2SIGMA2, lr=1e-3, gemm=args.train_gemm, relational=relational,
                 graph=args.train_graph)   # lr 0.005 (argParser.py:40) diverges on this synthetic set

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = [float(tr.step(pos, vis, valid)) for _ in range(max(args.warmup, 1))]
    barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = tr.step(pos, vis, valid)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    # all ranks hold identical weights after identical all-reduced updates
    w = flatten_bucket({k: getattr(params, k) for k in tr.keys})
    spread = torch.stack([w.min(), -w.max()])
    if world > 1:
        lo = spread.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(spread, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, spread))
    else:
        in_sync = True
    if rank == 0:
        emit({"mode": "train", "metric": "agent-trajectories/sec (training step: teacher-forced NLL + BPTT + gradient all-reduce + RMSProp)",
                          "value": int(valid_h.sum()) * world / (ms_step * 1e-3), "unit": "agent-trajectories/s", "n_gpus": world,
                          "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms_step, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32" if args.train_gemm == "fp32" else "f32 kernels, tf32 tensor-core contractions (fp32 accumulation)"
                                   + (": mmt_gemm_tf32 (TMA + tcgen05)" if args.train_gemm == "tc" else ": library GEMMs"),
                          "data": "synthetic",
                          "config": {"workload": f"{S} scenes x {N} agents per GPU, obs {T_OBS} / pred {P_PRED}, g2k_lstm_{'mcr' if relational else 'mc'} training step",
                                     "backward_gemm": args.train_gemm, "lr": 1e-3,
                                     "gradient_bucket_bytes": int(w.numel() * 4), "collective": "one NCCL all-reduce (SUM) per step" if world > 1 else "none (1 GPU)"},
                          "loss_first": losses[0], "loss_last": float(loss), "weights_identical_across_ranks": in_sync,
                          "cuda_graph": bool(args.train_graph),
                          "gpu_launches": int(ops.launch_count() - l0)})
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--prec", default=None, choices=["f16", "bf16", "f32", "bf16-stepwise", "bf16x3"],
                    help="default: f16 (fp16 tensor-core operands and state words, fp32 accumulation: inside the 1e-3 ADE/FDE bar "
                         "at the speed of bf16)")
    ap.add_argument("--variant", default=None, choices=["mc", "mcr"], help="default: mc (mcr for --config c2)")
    ap.add_argument("--scenes", type=int, default=4096)
    ap.add_argument("--agents", type=int, default=64)
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="launch kernels eagerly (no CUDA graph)")
    ap.add_argument("--no-modes", dest="modes", action="store_false", help="skip the short runs of the other precision modes")
    ap.add_argument("--parity-scenes", type=int, default=256, help="scenes of the batch compared with the CPU oracle")
    ap.add_argument("--config", default="c3", choices=["c3", "c1", "c2", "c5"],
                    help="c3 (default): synthetic crowds, the headline; c1 / c5: real-data evaluation on data/; c2: g2k_lstm_mcr "
                         "training steps on the real zara1 leave-one-out tables (extras)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"], help="train: data-parallel training steps (extra)")
    ap.add_argument("--train-graph", action="store_true",
                    help="--mode train: replay forward + BPTT of the shard as one CUDA graph (all-reduce / RMSProp outside it)")
    ap.add_argument("--train-gemm", default="fp32", choices=["fp32", "tf32", "tc"],
                    help="--mode train: contractions of the step: fp32 / tf32 library GEMMs, tc = mmt_gemm_tf32 (TMA + tcgen05) and mmt_aggregate_transpose_f32")
    args = ap.parse_args()
    if args.variant is None:
        args.variant = "mcr" if args.config == "c2" else "mc"
    if args.prec is None:
        args.prec = "f16" if (args.mode == "infer" and args.config != "c2" and args.impl == "ours") else "bf16"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    _claim_stdout()
    try:
        if args.impl == "reference":
            run_reference(args, rank, world)
        elif args.config in ("c1", "c5"):
            run_config(args, rank, world, local_rank)
        elif args.config == "c2":
            run_c2(args, rank, world, local_rank)
        elif args.mode == "train":
            run_train(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    except BaseException as exc:     # noqa: BLE001 -- post-mortem, then the original failure propagates
        if not isinstance(exc, SystemExit) or exc.code not in (0, None):
            record_failure(exc)
        raise


if __name__ == "__main__":
    try:
        from torch.distributed.elastic.multiprocessing.errors import record
        main = record(main)          # torchrun then shows this rank's traceback instead of "error_file: <N/A>"
    except Exception:                # noqa: BLE001
        pass
    main()
