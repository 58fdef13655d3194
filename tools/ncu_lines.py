#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: share of executed warp instructions, of stall samples, and
the dominant stall reasons.  usage: python tools/ncu_lines.py report.ncu-rep [top_n] [kernel-regex]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = None; hdr = None; recs = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":      # per-source-line aggregate row
        recs.append((fname, r))
ci = {}
for i, h in enumerate(hdr):
    ci.setdefault(h, i)
def f(r, k):
    try: return float(r[ci[k]])
    except Exception: return 0.0
ti = sum(f(r, "Instructions Executed") for _, r in recs); ts = sum(f(r, "# Samples") for _, r in recs)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total warp-instructions %.4e, samples %d" % (ti, ts))
agg = collections.Counter()
for _, r in recs:
    for s_ in stalls: agg[s_] += f(r, s_)
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / ts) for k, v in agg.most_common(9)))
print("%-16s %5s %6s %6s  %-28s %s" % ("file", "line", "inst%", "samp%", "top stalls", "source"))
for fn, r in sorted(recs, key=lambda t: -f(t[1], "# Samples"))[:topn]:
    st = sorted(((f(r, s_), s_[6:]) for s_ in stalls), reverse=True)[:2]
    sts = " ".join("%s:%.0f%%" % (n, 100 * v / max(f(r, "# Samples"), 1)) for v, n in st)
    print("%-16s %5s %6.2f %6.2f  %-28s %s" % (fn[:16], r[0], 100 * f(r, "Instructions Executed") / ti,
                                              100 * f(r, "# Samples") / ts, sts, r[1].strip()[:90]))
