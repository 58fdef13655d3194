#!/usr/bin/env python
"""Stress test of the segmented chains of the CTA-per-walker kernel (flag hand-over of a walker between CTAs): random
(n, batch, steps), forced 2-step segments and the launcher's own choice against whole chains; everything must be bit-identical.
usage: python tools/stress_segments.py [trials]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(2024)
cg.set_sweep_kernel("cta")
bad = 0
for t in range(trials):
    tag = ("n57_rs0.25", "n49_rs0.5", "n29_rs1")[t % 3]
    c = load_case(tag)
    B = int(rng.integers(200, 900)); S = int(rng.integers(3, 40))
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    x0 = np.tile(c["x"], (reps, 1, 1))[:B] + 0.05 * rng.standard_normal((B, c["n"], 2))
    si = np.tile(c["state_idx"], (reps, 1))[:B]
    logp = cg.make_logp(c["ktable"], c["L"])
    out = {}
    for mode in (False, True, None):
        cg.set_debug(segments=mode)
        xs, acc, tr = cg.mcmc(logp.bind(c["params"], si), x0.copy(), 1000 + t, S, 0.1, wrap_L=c["L"], return_trace=True)
        out[mode] = (xs, tr["naccept"], tr["logp"])
    for mode in (True, None):
        ok = all(np.array_equal(a, b) for a, b in zip(out[False], out[mode]))
        bad += not ok
        if not ok:
            print("MISMATCH trial %d %s B=%d S=%d mode=%s" % (t, tag, B, S, mode), flush=True)
cg.set_debug(segments=None)
print("stress_segments: %d trials, %d mismatches" % (trials, bad))
sys.exit(1 if bad else 0)
