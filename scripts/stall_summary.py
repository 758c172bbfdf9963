#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu report (needs `ncu -i rep --page source --csv --print-source sass`)."""
import csv
import subprocess
import sys


def I(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main(rep, kernel_filter=None, top=28):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # split per kernel: a row with "Kernel Name" starts a section
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[])
            secs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for sec in secs:
        if kernel_filter and kernel_filter not in sec["name"]:
            continue
        hdr = sec["rows"][0]
        data = [r for r in sec["rows"][1:] if len(r) == len(hdr)]
        isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
        stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(I(r[isamp]) for r in data)
        print("==", sec["name"][:80], "samples", tot, "instrs", len(data))
        agg = {s: sum(I(r[hdr.index(s)]) for r in data) for s in stalls}
        print("  ", ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
        for i, r in enumerate(data):
            if I(r[isamp]) > 0.012 * tot:
                st = {s[6:]: I(r[hdr.index(s)]) for s in stalls}
                st = {k: v for k, v in st.items() if v > 0.2 * I(r[isamp])}
                print(f"  {i:5d} {I(r[isamp]) * 100 / tot:5.1f}% ex={r[iex]:>9s} {r[isrc][:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
