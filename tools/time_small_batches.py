#!/usr/bin/env python
"""A/B timing of the sweep kernels (CGB_SWEEP_KERNEL=warp|cta|spec) over batch sizes: where the CTA-per-walker kernel
(4-8 warps cooperate on one walker: low latency) and the speculative mode of the warp kernel (one CTA per walker evaluates
the next 2-3 steps' proposals at once) beat the ready-queue warp-per-walker kernel (throughput).  50-proposal sweeps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import coulombgas_b200 as cg
from conftest import load_case

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

for tag, Bs in (("n13_rs10", (64, 148, 256, 444, 512, 768)), ("n29_rs1", (64, 148, 200, 256, 296, 444, 512))):
    c = load_case(tag)
    logp = cg.make_logp(c["ktable"], c["L"])
    for B in Bs:
        reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
        x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
        si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
        bound = logp.bind(c["params"], si)
        out = []
        names = ("warp", "cta", "spec4", "spec8")
        for k in names:
            os.environ["CGB_SWEEP_KERNEL"] = k
            out.append(timeit(lambda: cg.mcmc(bound, x.clone(), 3, 50, 0.1)))
        print("%-9s B=%5d  warp %7.3f ms  cta %7.3f ms  spec4 %7.3f ms  spec8 %7.3f ms  -> %s" % (tag, B, out[0], out[1], out[2], out[3], names[int(np.argmin(out))]), flush=True)
