#!/usr/bin/env python
"""Static SASS size per source line / per kernel phase of one sweep_kernel instantiation (instruction-cache
budgeting).  usage: python tools/sass_static.py [NW]   (needs coulombgas_b200/csrc/cgb_sweep.o)"""
import subprocess, sys, re, os, collections, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NW = sys.argv[1] if len(sys.argv) > 1 else "4"
obj = os.path.join(ROOT, "coulombgas_b200", "csrc", "cgb_sweep.o")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
name = ".text._ZN3cgb12sweep_kernelILi%sEEEvNS_11SweepParamsE:" % NW
start = next(i for i, l in enumerate(txt) if l.strip() == name)
src = os.path.join(ROOT, "coulombgas_b200", "csrc", "cgb_sweep.cu")
marks = []
for ln, text in enumerate(open(src), 1):
    m = re.match(r"\s*// -{8,} ([^-:(]+)", text)
    if m and ln > 150: marks.append((ln, m.group(1).strip()[:22]))
cur = ("?", 0); per = collections.Counter(); perfile = collections.Counter(); total = 0
ops = collections.Counter()
for l in txt[start + 1:]:
    if l.startswith("//---") or ".section" in l and ".text." in l: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        total += 1; per[cur] += 1; perfile[cur[0]] += 1; ops[m.group(2).split(".")[0]] += 1
print("sweep_kernel<%s>: %d SASS instructions = %.1f KiB" % (NW, total, total * 16 / 1024))
agg = collections.OrderedDict([("setup", 0)] + [(m[1], 0) for m in marks] + [("inlined helpers", 0)])
for (fn, ln), c in per.items():
    if fn == "cgb_sweep.cu" and ln >= 150:
        nm = "setup"
        for a, t in marks:
            if ln >= a: nm = t
        agg[nm] += c
    else: agg["inlined helpers"] += c
for k, v in agg.items(): print("  %-24s %6d" % (k, v))
print("  top inlined lines:", [(k, v) for k, v in per.most_common(60) if not (k[0] == "cgb_sweep.cu" and k[1] >= 150)][:12])
print("  top opcodes:", ops.most_common(14))
