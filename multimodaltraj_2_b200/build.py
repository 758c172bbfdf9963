"""Builds ``lib/libmmt.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a only."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libmmt.so"
SOURCES = ["capi.cu", "pairwise.cu", "aggregate.cu", "cell_f32.cu", "cell_tc.cu", "edge_mlp.cu",
           "decode_score.cu", "track_a.cu", "scene_batch.cu", "forecast.cu", "scores.cu", "graph_agg.cu", "graph_mma.cu", "rollout_tc.cu", "edge_mlp_tc.cu", "train_step.cu", "static_ctx.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "mmt_common.cuh", CSRC / "tc_common.cuh", PKG.parent / "include" / "mmt.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    objdir = PKG / "lib" / "obj"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for s in SOURCES:
        o = objdir / (s[:-3] + ".o")
        procs.append((s, o, subprocess.Popen([nvcc, *NVCC_FLAGS, "-c", str(CSRC / s), "-o", str(o)],
                                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, o, p in procs:
        out, _ = p.communicate()
        log.append(f"== {s}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    (PKG / "lib" / "build.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.check_call([nvcc, "-shared", "-o", str(LIB), *[str(o) for _, o, _ in procs], "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
