#!/usr/bin/env python
"""Calibration of the FP32 Slater LU's error indicator (warp kernel): log|det D| of the FP32 factorisation against the
float64 LU on many walkers, binned by the indicator  es = sum_k k / |u_kk|^2  that the kernel accumulates.
Needs the diagnostic build:  tools/build_variant.sh diag -DCGBW_DIAG_ERRSUM=1 ;
CGB_LIB=build/variants/libcgb200_diag.so python tools/calib_slater32.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

cg.set_sweep_kernel("warp")
rng = np.random.default_rng(0)
edges = [0, 30, 100, 300, 1e3, 3e3, 1e4, 3e4, 1e5, 1e6, 1e30]
for tag in ("n29_rs1", "n29_rs10", "n13_rs10"):
    c = load_case(tag)
    B = 32768
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    M = c["ktable"].shape[0]
    for mode in ("ground+subset states, jitter 0.1", "uniform random walkers"):
        if mode.startswith("uniform"):
            x = rng.uniform(0, c["L"], (B, c["n"], 2))
        else:
            x = np.tile(c["x"], (reps, 1, 1))[:B] + 0.1 * rng.standard_normal((B, c["n"], 2))
        si = np.stack([np.arange(M - c["n"], M) if b % 2 == 0 else np.sort(rng.choice(M, c["n"], replace=False)) for b in range(B)]).astype(np.int32)
        logp = cg.make_logp(c["ktable"], c["L"])
        cg.set_debug(force_f64_slater=False)
        lp32, parts32 = logp(x, c["params"], si, return_parts=True)
        cg.set_debug(force_f64_slater=True)
        lp64, parts64 = logp(x, c["params"], si, return_parts=True)
        cg.set_debug(force_f64_slater=False)
        es = parts32[:, 1]
        d = np.abs(parts32[:, 0] - parts64[:, 0])
        ok = np.isfinite(d)
        print("%s, %s: B=%d, |logp| median %.1f; max |d log|det D|| = %.2e; indicator quantiles 50/90/99/99.9%% = %s" % (
            tag, mode, B, np.median(np.abs(lp64)), d[ok].max(), np.round(np.quantile(es[ok], [0.5, 0.9, 0.99, 0.999]), 1)))
        for lo, hi in zip(edges[:-1], edges[1:]):
            m = ok & (es >= lo) & (es < hi)
            if m.any():
                print("   es in [%8.0f, %8.0f): %6d walkers (%.3f%%)  max err %.2e  rms %.2e  max err/(6e-8 sqrt(es)) = %.2f" % (
                    lo, hi, m.sum(), 100 * m.mean(), d[m].max(), np.sqrt((d[m] ** 2).mean()), (d[m] / (6e-8 * np.sqrt(es[m]))).max()))
