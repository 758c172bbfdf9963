#!/usr/bin/env python
"""Warp-stall samples of an ncu report aggregated between landmark SASS instructions (barriers, TMEM loads,
MMAs, ...): a phase-level view of a long warp-specialised kernel.  usage: region_summary.py rep [kernel]"""
import csv
import re
import subprocess
import sys

LAND = re.compile(r"BAR\.SYNC|SYNCS|LDTM|UTCHMMA|UTCBAR|UBLKCP|FENCE|WARPSYNC|EXIT|BRA ")


def I(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main(rep, kfilter=None):
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
        if kfilter and kfilter not in sec["name"]:
            continue
        hdr = sec["rows"][0]
        data = [r for r in sec["rows"][1:] if len(r) == len(hdr)]
        isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        sidx = {s: hdr.index(s) for s in stalls}
        tot = sum(I(r[isamp]) for r in data)
        print("==", sec["name"][:80], "samples", tot)
        acc, accst, n0, nins, mufu = 0, {s: 0 for s in stalls}, 0, 0, 0
        for i, r in enumerate(data):
            acc += I(r[isamp])
            nins += 1
            mufu += "MUFU" in r[isrc]
            for s in stalls:
                accst[s] += I(r[sidx[s]])
            if LAND.search(r[isrc]):
                if acc > 0.004 * tot:
                    top = sorted(accst.items(), key=lambda x: -x[1])[:3]
                    print(f"  [{n0:5d}-{i:5d}] {acc * 100 / tot:5.1f}%  n={nins:4d} mufu={mufu:3d} ex={r[iex]:>9s}  "
                          f"{r[isrc].strip()[:48]:48s} " + " ".join(f"{k[6:]}={v * 100 // max(acc, 1)}%" for k, v in top))
                acc, accst, n0, nins, mufu = 0, {s: 0 for s in stalls}, i + 1, 0, 0


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
