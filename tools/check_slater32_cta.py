#!/usr/bin/env python
"""FP32 Slater LU of the CTA-per-walker kernel against its float64 LU: log|det D| on many walkers per config, the fallback rate
under the error indicator, and the sweep timing with / without the FP32 path (CGB_FLAG_FORCE_F64_SLATER)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

cg.set_sweep_kernel("cta")
rng = np.random.default_rng(0)
for tag, B in (("n57_rs0.25", 4096), ("n49_rs0.5", 4096), ("n29_rs1", 8192), ("n13_rs10", 8192)):
    c = load_case(tag)
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    M = c["ktable"].shape[0]
    for mode in ("physical", "uniform"):
        x = rng.uniform(0, c["L"], (B, c["n"], 2)) if mode == "uniform" else np.tile(c["x"], (reps, 1, 1))[:B] + 0.1 * rng.standard_normal((B, c["n"], 2))
        si = np.stack([np.arange(M - c["n"], M) if b % 2 == 0 else np.sort(rng.choice(M, c["n"], replace=False)) for b in range(B)]).astype(np.int32)
        logp = cg.make_logp(c["ktable"], c["L"])
        f0 = logp.slater_fallback_count()
        lp32, p32 = logp(x, c["params"], si, return_parts=True)
        f1 = logp.slater_fallback_count()
        cg.set_debug(force_f64_slater=True)
        lp64, p64 = logp(x, c["params"], si, return_parts=True)
        cg.set_debug(force_f64_slater=False)
        ok = np.isfinite(p64[:, 0])
        d = np.abs(p32[:, 0] - p64[:, 0])[ok]
        rel = (np.abs(lp32 - lp64) / np.abs(lp64))[ok]
        print("%s %s: B=%d, float64 fallbacks %.2f %%, max |d log|det D|| %.2e (rms %.2e), max rel error of log p %.2e, |log p| median %.0f"
              % (tag, mode, B, 100.0 * (f1 - f0) / B, d.max(), np.sqrt((d ** 2).mean()), rel.max(), np.median(np.abs(lp64[ok]))), flush=True)
for tag, B in (("n57_rs0.25", 2048), ("n57_rs0.25", 512), ("n49_rs0.5", 2048), ("n57_rs0.25", 64)):
    c = load_case(tag)
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
    si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
    logp = cg.make_logp(c["ktable"], c["L"])
    bound = logp.bind(c["params"], si)
    cg.set_sweep_kernel(None)
    for f64 in (False, True):
        cg.set_debug(force_f64_slater=f64)
        ts = []
        for it in range(4):
            xx = x.clone()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); cg.mcmc(bound, xx, 3 + it, 50, 0.1); b.record(); torch.cuda.synchronize()
            if it: ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        print("%s B=%d sweep (%s Slater LU): %.3f ms  %.3e walker-steps/s  %.2f TFLOP/s" % (
            tag, B, "float64" if f64 else "FP32+monitor", ms, B * 50 / ms * 1e3, cg.flops.flop_sweep(c["n"], 50) * B / ms / 1e9), flush=True)
    cg.set_debug(force_f64_slater=False)
