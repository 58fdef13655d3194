#!/usr/bin/env python
"""Randomised stress of the speculative small-batch modes against the sequential kernels: random (case, B, mc_steps, stddev,
seed), throughput-mode Philox streams; accept masks and final walkers must be identical (spec4/spec8 vs warp, cta4/cta8 vs cta).
usage: python tools/stress_spec.py [trials]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(2026)
cases = {t: load_case(t) for t in ("n13_rs10", "n29_rs1", "n29_rs10", "n49_rs0.5", "n57_rs0.25")}
bad = 0
for it in range(trials):
    tag = list(cases)[rng.integers(len(cases))]
    c = cases[tag]
    B = int(rng.choice([1, 2, 5, 17, 64, 65, 100, 148, 149, 300, 444, 445, 600]))
    S = int(rng.choice([1, 2, 3, 4, 5, 9, 16, 33]))
    sd = float(rng.choice([0.02, 0.1, 0.3]))
    seed = int(rng.integers(1 << 40))
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    x0 = np.tile(c["x"], (reps, 1, 1))[:B] + 0.05 * rng.standard_normal((B, c["n"], 2))
    si = np.tile(c["state_idx"], (reps, 1))[:B]
    logp = cg.make_logp(c["ktable"], c["L"])
    groups = [("cta", ("cta4", "cta8"))] + ([("warp", ("spec4", "spec8"))] if c["n"] <= 32 else [])
    for ref, others in groups:
        cg.set_sweep_kernel(ref)
        xr, ar, tr = cg.mcmc(logp.bind(c["params"], si), x0.copy(), seed, S, sd, return_trace=True)
        for k in others:
            cg.set_sweep_kernel(k)
            xk, ak, tk = cg.mcmc(logp.bind(c["params"], si), x0.copy(), seed, S, sd, return_trace=True)
            ok = np.array_equal(tr["accept"], tk["accept"]) and np.array_equal(xr, xk) and np.array_equal(tr["naccept"], tk["naccept"]) \
                and np.abs(tr["logp_prop"] - tk["logp_prop"]).max() <= 1e-11 * np.abs(tr["logp_prop"]).max()
            if not ok:
                bad += 1
                print("MISMATCH", tag, B, S, sd, seed, ref, k, flush=True)
    cg.set_sweep_kernel(None)
    xa, aa, ta = cg.mcmc(logp.bind(c["params"], si), x0.copy(), seed, S, sd, return_trace=True)      # the library's own choice
    if not (np.array_equal(ta["accept"], tr["accept"]) or c["n"] <= 32):
        bad += 1
        print("MISMATCH auto", tag, B, S, sd, seed, flush=True)
print("stress_spec: %d trials, %d mismatches" % (trials, bad))
sys.exit(1 if bad else 0)
