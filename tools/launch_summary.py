#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, with -v, every launch."""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("cnb::", "").replace("<unnamed>::", "")
        rows.append((name, r.get("Grid Size", ""), us))
    return rows


def main():
    verbose = "-v" in sys.argv
    path = [a for a in sys.argv[1:] if a != "-v"][0]
    rows = load(path)
    tot = sum(r[2] for r in rows)
    print("total %.0f us over %d launches\n" % (tot, len(rows)))
    agg = OrderedDict()
    for n, g, us in rows:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
    print("| kernel | launches | us | share |\n|---|---|---|---|")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (n, c, us, 100 * us / tot))
    if verbose:
        print()
        for i, (n, g, us) in enumerate(rows):
            print("%3d %-48s %-16s %8.1f" % (i, n[:48], g, us))


if __name__ == "__main__":
    main()
