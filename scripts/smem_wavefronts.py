#!/usr/bin/env python
"""Shared-memory wavefronts of an ncu report per SASS opcode and per instruction (source page): where the LSU data pipe goes.
usage: smem_wavefronts.py rep [kernel-substring] [top-n]"""
import collections
import csv
import subprocess
import sys


def I(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main(rep, kfilter="", topn=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[])
            secs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for sec in secs:
        if kfilter not in sec["name"]:
            continue
        hdr = sec["rows"][0]
        data = [r for r in sec["rows"][1:] if len(r) == len(hdr)]
        isrc, iw, iex, iid = hdr.index("Source"), hdr.index("L1 Wavefronts Shared"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared Ideal")
        tot = sum(I(r[iw]) for r in data)
        print("==", sec["name"][:80], "shared wavefronts", tot)
        by = collections.defaultdict(lambda: [0, 0, 0])
        for r in data:
            w = I(r[iw])
            if not w:
                continue
            op = r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1]
            by[op][0] += w
            by[op][1] += I(r[iex])
            by[op][2] += I(r[iid])
        for op, (w, ex, idl) in sorted(by.items(), key=lambda x: -x[1][0]):
            print(f"  {op:28s} wavefronts {w:12d} ({100.0 * w / tot:5.1f} %)  executed {ex:11d}  wavefronts/instr {w / max(ex, 1):5.2f}  ideal {idl / max(ex, 1):5.2f}")
        print("  -- top instructions")
        for i, r in sorted(enumerate(data), key=lambda x: -I(x[1][iw]))[:topn]:
            print(f"  [{i:5d}] {I(r[iw]):11d} ex {I(r[iex]):10d} w/i {I(r[iw]) / max(I(r[iex]), 1):5.2f}  {r[isrc][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 25)
