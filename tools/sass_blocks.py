#!/usr/bin/env python
"""Loop structure of one kernel in an object file: every backward branch (= loop) with its static body size, opcode
histogram and the source lines it covers.  Static instruction budgeting of the sweep kernels without a GPU.
usage: python tools/sass_blocks.py obj kernel-substring [--src=cgb_wsweep.cu] [--min=20] [--list=LABEL]"""
import subprocess, re, os, sys, collections, tempfile

obj, kern = sys.argv[1], sys.argv[2]
src = "cgb_wsweep.cu"
minsize, want = 20, None
for a in sys.argv[3:]:
    if a.startswith("--src="): src = a[6:]
    if a.startswith("--min="): minsize = int(a[6:])
    if a.startswith("--list="): want = a[7:]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if kern in l and l.strip().startswith(".text."))
rows = []          # (kind, text, (file, line))
cur = ("", 0)
for l in txt[start + 1:]:
    if l.strip().startswith(".text."): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", l)
    if m: rows.append(("i", m.group(1).strip(), cur)); continue
    m = re.match(r"\s*(\.L_x_\d+):", l)
    if m: rows.append(("l", m.group(1), cur))
labels = {t: i for i, (k, t, c) in enumerate(rows) if k == "l"}
print("kernel has %d instructions" % sum(1 for r in rows if r[0] == "i"))
loops = []
for i, (k, t, c) in enumerate(rows):
    if k != "i": continue
    m = re.search(r"\bBRA\b.*?(\.L_x_\d+)", t)
    if m and m.group(1) in labels and labels[m.group(1)] < i:
        loops.append((labels[m.group(1)], i, m.group(1)))
loops.sort()
for a, b, lab in loops:
    body = [r for r in rows[a:b + 1] if r[0] == "i"]
    if len(body) < minsize: continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t, _ in body)
    lines = [c[1] for _, _, c in body if c[0] == src]
    inner = [l2 for a2, b2, l2 in loops if a2 > a and b2 < b]
    print("%-10s %5d instr  lines %s..%s  inner loops %s" % (lab, len(body), min(lines) if lines else "?", max(lines) if lines else "?", inner))
    print("           ", ops.most_common(12))
    if want == lab:
        for k, t, c in rows[a:b + 1]: print("    %5s %s" % (c[1] if c[0] == src else "h", t[:110]))
