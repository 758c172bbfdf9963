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
           "decode_score.cu", "track_a.cu", "scene_batch.cu", "forecast.cu", "scores.cu", "graph_agg.cu", "graph_mma.cu", "rollout_tc.cu", "edge_mlp_tc.cu", "train_step.cu", "static_ctx.cu", "collective.cu", "gemm_tc.cu", "edge_mlp_bwd.cu", "edge_mlp_bwd_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    return [CSRC / s for s in SOURCES] + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "mmt.h"]


def source_hash() -> str:
    """sha256 over the names and contents of every source the library is built from (+ the nvcc flags)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        h.update(d.name.encode())
        h.update(d.read_bytes())
    return h.hexdigest()


STAMP = PKG / "lib" / "libmmt.sha256"


def needs_build() -> bool:
    """True unless lib/libmmt.so exists AND was built from exactly the sources in the tree (content hash, not mtime:
    a snapshot copy or a checkout does not preserve modification times)."""
    return not (LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == source_hash())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile and link lib/libmmt.so unless it is up to date with the sources.  Serialised across processes by a file
    lock: under torchrun every rank calls this at import time, and only one may run nvcc."""
    import fcntl
    LIB.parent.mkdir(parents=True, exist_ok=True)
    with open(LIB.parent / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    if not force and not needs_build():
        print(f"[multimodaltraj_2_b200.build] {LIB.relative_to(PKG.parent)} is up to date with the sources "
              f"(hash {source_hash()[:16]}): not rebuilt", file=sys.stderr)
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
    subprocess.check_call([nvcc, "-shared", "-o", str(LIB), *[str(o) for _, o, _ in procs], "-lcudart", "-ldl"])
    STAMP.write_text(source_hash() + "\n")
    print(f"[multimodaltraj_2_b200.build] nvcc compiled {len(SOURCES)} sources for sm_100a -> {LIB.relative_to(PKG.parent)} "
          f"(source hash {source_hash()[:16]})", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
