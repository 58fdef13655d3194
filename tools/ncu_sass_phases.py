#!/usr/bin/env python
"""Per-phase and per-loop shares of executed warp instructions and stall samples of a sweep kernel from an ncu report
(--import-source on): the report's SASS listing (execution counts, samples, stall reasons per instruction) is joined, by
instruction order, with the line table of the object file the kernel was built from, so that inlined helpers (packed FMA
intrinsics, MUFU wrappers) are attributed to the phase whose code surrounds them.
usage: python tools/ncu_sass_phases.py report.ncu-rep obj kernel-substring evals [--src=cgb_wsweep.cu] [--loops] [--dump=LO:HI]"""
import csv, subprocess, sys, io, collections, os, re, tempfile
rep, obj, kern, evals = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
srcname = "cgb_wsweep.cu"; loops = False; dump = None
for a in sys.argv[5:]:
    if a.startswith("--src="): srcname = a[6:]
    if a == "--loops": loops = True
    if a.startswith("--dump="): dump = tuple(int(v) for v in a[7:].split(":"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "coulombgas_b200", "csrc", srcname)
marks = []
kstart = next(ln for ln, t in enumerate(open(src), 1) if "sweep_kernel(const SweepParams" in t)
for ln, text in enumerate(open(src), 1):
    m = re.match(r"\s*// -{8,} ([^-:(]+)", text)
    if m and ln > kstart: marks.append((ln, m.group(1).strip()[:26]))
# --- line table from the object
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if kern in l and l.strip().startswith(".text."))
lines = []; cur = ("", 0); last_src = 0
for l in txt[start + 1:]:
    if l.strip().startswith(".text."): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", l)
    if m:
        if cur[0] == srcname and cur[1] >= kstart: last_src = cur[1]
        lines.append((cur, last_src, m.group(1).strip()))
# --- executed counts from the report
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; recs = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
isrc, ii, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sc = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
if len(recs) != len(lines):
    print("warning: report has %d instructions, object has %d (different build?)" % (len(recs), len(lines)))
n = min(len(recs), len(lines))
ti = sum(float(r[ii]) for r in recs); ts = sum(float(r[ism]) for r in recs)
agg = collections.OrderedDict()
def phase_of(ln):
    nm = "setup"
    for a, t in marks:
        if ln >= a: nm = t
    return nm
for k in range(n):
    (fl, ln), last, text = lines[k]
    ph = phase_of(ln if (fl == srcname and ln >= kstart) else last)
    a = agg.setdefault(ph, [0.0, 0.0, collections.Counter()])
    a[0] += float(recs[k][ii]); a[1] += float(recs[k][ism])
    for i, nm in sc: a[2][nm] += float(recs[k][i])
print("total %.4e warp-instructions = %.0f per evaluation, %d stall samples" % (ti, ti / evals, ts))
for ph, (i_, s_, st) in agg.items():
    top = " ".join("%s:%.0f%%" % (nm, 100 * v / max(s_, 1)) for nm, v in st.most_common(4))
    print("%-28s inst %6.2f%% (%6.0f/eval)  samples %6.2f%%  cyc/inst %.2f  %s" % (ph, 100 * i_ / ti, i_ / evals, 100 * s_ / ts, (s_ / ts) / max(i_ / ti, 1e-9), top))
if dump:
    for k in range(n):
        (fl, ln), last, text = lines[k]
        l2 = ln if (fl == srcname and ln >= kstart) else last
        if dump[0] <= l2 <= dump[1]:
            r = recs[k]
            st = sorted([(int(r[i]), nm) for i, nm in sc if int(r[i]) > 0], reverse=True)[:3]
            print("%5d %5s %-64s %9s %6s  %s" % (k, ln if fl == srcname else "h", text[:64], r[ii], r[ism], " ".join("%s:%d" % (nm, v) for v, nm in st)))
