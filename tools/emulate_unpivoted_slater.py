#!/usr/bin/env python
"""Developer experiment (CPU, NumPy emulation of FP32 arithmetic; imports the oracle flow -- not a product path): would an UNPIVOTED
FP32 Slater LU (rows retire as in the Jacobian LU, no pivot search) pass the error monitor often enough to pay?  Prints the error of
log|det D| against LAPACK float64 and the pass rate under thresholds of the indicator es = sum_k k / |u_kk|^2.
Result (profiles/r02_unpivoted_slater_emulation.txt): at es < 300 only 86 % (n=29) / 67 % (n=49) / 52 % (n=57) of equilibrated walkers
pass -- the fallback to the pivoted LU would eat the saving; rejected."""
import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coulombgas_b200.fixtures import load_case
from oracle import flow as oflow
rng = np.random.default_rng(0)
def lu_unpiv32(D):
    A = D.astype(np.complex64).copy(); n = A.shape[0]
    logdet = 0.0; es = 0.0; maxl = 0.0; minpiv = 1e9
    for k in range(n):
        p = A[k, k]
        pn = np.float32(abs(p) ** 2)
        logdet += 0.5 * np.log(float(pn)); es += (k + 1) / float(pn); minpiv = min(minpiv, abs(p))
        if k == n - 1: break
        l = (A[k + 1:, k] / p).astype(np.complex64)
        maxl = max(maxl, np.abs(l).max())
        A[k + 1:, k + 1:] = (A[k + 1:, k + 1:] - np.outer(l, A[k, k + 1:])).astype(np.complex64)
    return logdet, es, maxl, minpiv
def lu_piv32(D):
    A = D.astype(np.complex64).copy(); n = A.shape[0]
    logdet = 0.0; es = 0.0
    for k in range(n):
        r = k + np.argmax(np.abs(A[k:, k].real) + np.abs(A[k:, k].imag))
        A[[k, r]] = A[[r, k]]
        p = A[k, k]; pn = np.float32(abs(p) ** 2)
        logdet += 0.5 * np.log(float(pn)); es += (k + 1) / float(pn)
        if k == n - 1: break
        l = (A[k + 1:, k] / p).astype(np.complex64)
        A[k + 1:, k + 1:] = (A[k + 1:, k + 1:] - np.outer(l, A[k, k + 1:])).astype(np.complex64)
    return logdet, es
for tag, nw in (("n29_rs1", 200), ("n29_rs10", 64), ("n57_rs0.25", 100), ("n49_rs0.5", 100), ("n13_rs10", 64)):
    c = load_case(tag); n = c["n"]; L = c["L"]
    params = oflow.to_torch(c["params"])
    X = c["x"]; reps = (nw + X.shape[0] - 1) // X.shape[0]
    X = np.tile(X, (reps, 1, 1))[:nw] + (0.1 * rng.standard_normal((nw, n, 2)) if reps > 1 else 0)
    errs_u, errs_p, ess, maxls, minp = [], [], [], [], []
    for b in range(nw):
        z = oflow.flow(params, torch.as_tensor(X[b]), L).numpy()
        si = c["state_idx"][b % c["state_idx"].shape[0]]
        k = 2 * np.pi / L * c["ktable"][si]
        w = z - z.mean(0)      # common displacement dropped (phase)
        D = np.exp(1j * (w @ k.T))
        ref = np.linalg.slogdet(D)[1]
        ld, es, ml, mp = lu_unpiv32(D)
        ldp, esp = lu_piv32(D)
        errs_u.append(abs(ld - ref)); errs_p.append(abs(ldp - ref)); ess.append(es); maxls.append(ml); minp.append(mp)
    errs_u = np.array(errs_u); ess = np.array(ess); maxls = np.array(maxls)
    print(tag, "unpivoted: err quantiles 50/90/99/max", np.quantile(errs_u, [.5, .9, .99, 1]), " pivoted max", max(errs_p))
    print("   es quantiles", np.quantile(ess, [.5, .9, .99, 1]).round(0), "max|l| q", np.quantile(maxls, [.5, .9, .99, 1]).round(1), "min piv q", np.quantile(minp, [0, .01, .1, .5]).round(3))
    for thr in (300, 1200, 5000, 20000, 1e5):
        ok = ess < thr
        print("   es<%g: %.0f%% pass, max err among passing %.2e" % (thr, 100 * ok.mean(), errs_u[ok].max() if ok.any() else 0))
