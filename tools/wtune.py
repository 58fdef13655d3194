#!/usr/bin/env python
"""Timing sweep of the warp-per-walker kernel over batch sizes and launch knobs (CGB_W_WPC / CGB_W_GRID).
usage: python tools/wtune.py case B [B ...]   (env knobs are read by the library at launch time)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

tag = sys.argv[1]
c = load_case(tag)
logp = cg.make_logp(c["ktable"], c["L"])
for B in [int(a) for a in sys.argv[2:]]:
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
    si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
    bound = logp.bind(c["params"], si)
    for _ in range(2):
        cg.mcmc(bound, x.clone(), 1, 50, 0.1)
    torch.cuda.synchronize()
    best = 1e9
    for r in range(3):
        xx = x.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _, acc = cg.mcmc(bound, xx, 2, 50, 0.1); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ws = logp.workspace(B, c["n"])
    spins = int(ws[4096 + 80:4096 + 88].view(torch.int64).item())
    print("%s B=%d wpc=%s grid=%s: %.3f ms %.3e walker-steps/s acc %.3f spins %d" % (
        tag, B, os.environ.get("CGB_W_WPC", "-"), os.environ.get("CGB_W_GRID", "-"), best, B * 50 / best * 1e3, float(acc), spins), flush=True)
