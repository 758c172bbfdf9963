#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
            continue
        name = r[kn].split('(')[0][:56]
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':58s} {'n':>5s} {'avg us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
        print(f"{k:58s} {v[0]:5d} {v[1] / v[0] / 1e3:9.1f} {v[1] / tot * 100:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
