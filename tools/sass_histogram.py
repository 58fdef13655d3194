#!/usr/bin/env python
"""SASS opcode histogram per kernel of libcgb200.so (cuobjdump -sass): the repository's own evidence of what the kernels are
made of -- packed FP32 FMAs (FFMA2 / FMUL2 / FADD2), float64 pipe (DFMA / DADD / DMUL), SFU (MUFU), warp collectives
(SHFL / REDUX / CREDUX), cluster barriers (UCGABAR_ARV / UCGABAR_WAIT), TMA bulk copies of the weight blocks (UBLKCP, with
SYNCS mbarrier operations), spills (LDL / STL), and the absence of tensor-core instructions (UTC*MMA / HMMA: the contractions are
K = 5/9/48, N = 16).
usage: python tools/sass_histogram.py [lib.so] > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "coulombgas_b200", "libcgb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern = None
hist = collections.OrderedDict()
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", l)
    if m and kern:
        hist[kern][m.group(1)] += 1
WATCH = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "DFMA", "DADD", "DMUL", "MUFU", "SHFL", "REDUX", "CREDUX", "LDS", "STS",
         "LDG", "STG", "ATOMG", "LDC", "LDCU", "LDL", "STL", "BAR", "UCGABAR_ARV", "UCGABAR_WAIT", "MEMBAR", "HMMA", "UTCHMMA", "UTMALDG",
         "UBLKCP", "SYNCS", "LDGSTS"]
print("# SASS opcode histogram of %s (static instruction counts per kernel)" % os.path.basename(lib))
tot = collections.Counter()
for k, h in hist.items():
    n = sum(h.values())
    tot.update(h)
    short = re.sub(r"\(cgb::\w+\)$|\(.*\)$", "", k).replace("void ", "")
    print("\n%s: %d instructions (%.1f KiB)" % (short, n, n * 16 / 1024))
    print("   " + "  ".join("%s %d" % (o, h[o]) for o in WATCH if h[o]))
    rest = [(o, c) for o, c in h.most_common() if o not in WATCH][:10]
    print("   other: " + "  ".join("%s %d" % oc for oc in rest))
print("\n# whole library: " + "  ".join("%s %d" % (o, tot[o]) for o in WATCH))
