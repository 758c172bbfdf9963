"""build/libmmt_<name>.so = the in-tree library with rollout_tc.cu recompiled with extra nvcc flags (kernel experiments):
    python scratch/build_variant.py <name> [--src cell_tc.cu] [-DMACRO[=v] ...]"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from multimodaltraj_2_b200 import build as b  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
src = "rollout_tc.cu"
if flags and flags[0] == "--src":
    src, flags = flags[1], flags[2:]
b.build()
obj = ROOT / "build" / f"{src[:-3]}_{name}.o"
obj.parent.mkdir(exist_ok=True)
out = subprocess.run([b._nvcc(), *b.NVCC_FLAGS, *flags, "-c", str(b.CSRC / src), "-o", str(obj)],
                     capture_output=True, text=True)
if out.returncode:
    sys.exit(out.stdout + out.stderr)
lines = (out.stdout + out.stderr).splitlines()
for i, ln in enumerate(lines):
    if ("rollout_tc_kernelILb0" in ln or "gsk_cell_tc_kernelILi0ELb1" in ln) and "Function properties" in ln:
        print(name, "|", lines[i + 1].strip(), "|", lines[i + 2].strip())
objs = [str(obj if s == src else b.PKG / "lib" / "obj" / (s[:-3] + ".o")) for s in b.SOURCES]
subprocess.check_call([b._nvcc(), "-shared", "-o", str(ROOT / "build" / f"libmmt_{name}.so"), *objs, "-lcudart"])
