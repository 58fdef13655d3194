#!/usr/bin/env python
"""Timing of the VAN sampler kernel (draw of the n occupied orbitals + log_prob) at the shipped shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_van

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

for tag in ("n13_pretrained", "n29_rs1", "n57_rs0.25"):
    c = load_van(tag)
    v = cg.VanSampler(c["ktable"], c["n"], c["M"])
    p = torch.as_tensor(cg.pack_van_params(c["params"])).cuda()
    for B in (256, 2048, 16384):
        ms = timeit(lambda: v.sampler(p, 5, B, return_log_prob=True))
        s = v.sampler(p, 5, B)
        ml = timeit(lambda: v.log_prob(p, s))
        print("%-15s n=%2d M=%3d B=%5d  sample %8.3f ms (%.3e walkers/s)  log_prob %8.3f ms" % (tag, c["n"], c["M"], B, ms, B / ms * 1e3, ml), flush=True)
