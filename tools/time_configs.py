#!/usr/bin/env python
"""Time the sweep and Ewald kernels on every BASELINE.json config shape (named per-GPU batch and a saturating batch)."""
import os, sys, json
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

rows = []
for tag, batches in (("n13_rs10", (256, 512, 8192)), ("n29_rs1", (256, 2048, 8192)), ("n49_rs0.5", (64, 512, 2048)),
                     ("n57_rs0.25", (64, 512, 2048))):
    c = load_case(tag)
    logp = cg.make_logp(c["ktable"], c["L"])
    for B in batches:
        reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
        x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
        si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
        bound = logp.bind(c["params"], si)
        ms = timeit(lambda: cg.mcmc(bound, x.clone(), 3, 50, 0.1))
        R = 100                                   # back-to-back launches: one call per event pair would time the launch latency
        me = timeit(lambda: [cg.potential_energy(x, c["kappa"], c["G"], c["L"], c["rs"]) for _ in range(R)], 3) / R
        n = c["n"]
        fl = cg.flops.flop_sweep(n, 50) * B / (ms * 1e-3) / 1e12
        fe = cg.flops.flop_ewald(n) * B / (me * 1e-3) / 1e12
        rows.append(dict(case=tag, n=n, B=B, sweep_ms=ms, walker_steps_per_s=B * 50 / ms * 1e3, sweep_tflops=fl,
                         ewald_ms=me, ewald_evals_per_s=B / me * 1e3, ewald_tflops=fe, fallbacks=logp.fallback_count()))
        print("%-11s B=%5d sweep %8.3f ms %.3e w-steps/s %5.2f TF | ewald %7.4f ms %.3e ev/s %5.2f TF | fb %d"
              % (tag, B, ms, rows[-1]["walker_steps_per_s"], fl, me, rows[-1]["ewald_evals_per_s"], fe, rows[-1]["fallbacks"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "time_configs.json"), "w"), indent=1)
