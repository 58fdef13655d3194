#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, mean duration, share.
usage: python tools/ncu_launches.py launches.csv"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, ig, ib, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size", "Metric Unit"))
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[ik] == "Kernel Name": continue
    try: v = float(r[iv].replace(",", ""))
    except ValueError: continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0, r[ig], r[ib]])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("# launches %d, total %.3f ms (per-launch times under ncu are cold-cache and serialised: shares, not absolutes)" % (sum(a[0] for a in agg.values()), tot / 1e3))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s n=%4d  avg %10.1f us  share %5.1f%%  grid %s block %s" % (name[:72], a[0], a[1] / a[0], 100 * a[1] / tot, a[2], a[3]))
