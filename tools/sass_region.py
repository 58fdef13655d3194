#!/usr/bin/env python
"""Opcode histogram / listing of the SASS attributed to a source-line range of one kernel in an object file.
usage: python tools/sass_region.py obj kernel-substring first_line last_line [--list]"""
import subprocess, re, os, sys, collections, tempfile
obj, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if kern in l and l.strip().startswith(".text."))
rows = []; cur = ("", 0)
for l in txt[start + 1:]:
    if l.strip().startswith(".text."): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", l)
    if m: rows.append((cur, m.group(1).strip()))
    elif l.strip().startswith(".L_"): rows.append((cur, l.strip()))
main = os.path.basename([a for a in sys.argv if a.endswith(".o")][0]).replace(".o", ".cu")
for a in sys.argv:
    if a.startswith("--src="): main = a[6:]
idx = [i for i, (c, t) in enumerate(rows) if c[0] == main and lo <= c[1] <= hi]
a, b = min(idx), max(idx)
ops = collections.Counter()
for c, t in rows[a:b + 1]:
    if t.startswith(".L_"): continue
    t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
    ops[t2.split()[0].split(".")[0]] += 1
print("instructions in [%d..%d] address span: %d" % (lo, hi, sum(ops.values())))
print(ops.most_common(40))
if "--list" in sys.argv:
    for c, t in rows[a:b + 1]: print(c[1], t[:100])
