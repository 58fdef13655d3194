#!/usr/bin/env python
"""Share of executed warp instructions and of stall samples per kernel phase of cgb_sweep.cu, from an ncu
report captured with --import-source on.  Phase boundaries are found from the "// ----------------" section
comments of the source file.   usage: python tools/ncu_phases.py report.ncu-rep [evals] [source file name]"""
import csv, subprocess, sys, io, collections, os, re
rep = sys.argv[1]
evals = float(sys.argv[2]) if len(sys.argv) > 2 else None
srcname = sys.argv[3] if len(sys.argv) > 3 else "cgb_sweep.cu"
src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "coulombgas_b200", "csrc", srcname)
marks = []
for ln, text in enumerate(open(src), 1):
    m = re.match(r"\s*// -{8,} ([^-:(]+)", text)
    if m and ln > 150: marks.append((ln, m.group(1).strip()[:22]))
kstart = next(ln for ln, t in enumerate(open(src), 1) if "sweep_kernel(const SweepParams" in t)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = None; hdr = None; recs = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-": recs.append((fname, r))
ci = {}
for i_, h in enumerate(hdr): ci.setdefault(h, i_)
def f(r, k):
    try: return float(r[ci[k]])
    except Exception: return 0.0
ti = sum(f(r, "Instructions Executed") for _, r in recs); ts = sum(f(r, "# Samples") for _, r in recs)
agg = collections.OrderedDict([("setup", [0, 0])] + [(m[1], [0, 0]) for m in marks] + [("helpers/inlined", [0, 0])])
for fn, r in recs:
    ln = int(r[0]); ie = f(r, "Instructions Executed"); sm = f(r, "# Samples")
    if fn == srcname and ln >= kstart:
        name = "setup"
        for a, nm in marks:
            if ln >= a: name = nm
        agg[name][0] += ie; agg[name][1] += sm
    else:
        agg["helpers/inlined"][0] += ie; agg["helpers/inlined"][1] += sm
print("total warp-instructions %.4e, stall samples %d" % (ti, ts))
for k, v in agg.items():
    extra = "  %7.0f warp-inst/eval" % (v[0] / evals) if evals else ""
    print("%-24s inst %6.2f%%  samples %6.2f%%%s" % (k, 100 * v[0] / ti, 100 * v[1] / ts, extra))
