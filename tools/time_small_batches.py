#!/usr/bin/env python
"""A/B timing of the sweep kernels (CGB_SWEEP_KERNEL=warp|cta|spec) over batch sizes: where the CTA-per-walker kernel
(4-8 warps cooperate on one walker: low latency) and the speculative mode of the warp kernel (one CTA per walker evaluates
the next 2-3 steps' proposals at once) beat the ready-queue warp-per-walker kernel (throughput).  50-proposal sweeps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

KERNELS = ("warp", "spec4", "spec8", "cta", "cta4", "cta8", "auto")
for tag, Bs in (("n13_rs10", (32, 64, 100, 148, 256, 444, 512, 768, 888, 1024)), ("n29_rs1", (32, 64, 92, 148, 256, 444, 512, 640, 740, 888)),
                ("n49_rs0.5", (16, 32, 64, 74, 128, 512)), ("n57_rs0.25", (16, 32, 64, 74, 128, 512))):
    c = load_case(tag)
    logp = cg.make_logp(c["ktable"], c["L"])
    for B in Bs:
        reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
        x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
        si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
        bound = logp.bind(c["params"], si)
        out = []
        names = [k for k in KERNELS if c["n"] <= 32 or not (k.startswith("spec") or k == "warp")]
        for k in names:
            if k == "auto":
                cg.set_sweep_kernel(None)
            else:
                cg.set_sweep_kernel(k)
            out.append(timeit(lambda: cg.mcmc(bound, x.clone(), 3, 50, 0.1)))
        print("%-10s B=%5d  " % (tag, B) + "  ".join("%s %6.3f" % (k, v) for k, v in zip(names, out)) +
              "  ms -> %s" % names[int(np.argmin(out[:-1]))], flush=True)
