#!/usr/bin/env python
"""Per-phase cycle shares of the warp-per-walker sweep kernel from a developer build with -DCGBW_PROF=1
(tools/build_variant.sh prof -DCGBW_PROF=1; CGB_LIB=build/variants/libcgb200_prof.so python tools/phase_clocks.py [case] [B] [S]).
Every warp accumulates clock64() differences at the phase boundaries; the sums land in the workspace."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

tag = sys.argv[1] if len(sys.argv) > 1 else "n29_rs1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
S = int(sys.argv[3]) if len(sys.argv) > 3 else 50
kern = sys.argv[4] if len(sys.argv) > 4 else "warp"
c = load_case(tag)
reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
rng = np.random.default_rng(0)
x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B] + 0.05 * rng.standard_normal((B, c["n"], 2)), device="cuda")
si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
logp = cg.make_logp(c["ktable"], c["L"])
bound = logp.bind(c["params"], si)
cg.set_sweep_kernel(kern)
names = ["pop+proposal+setup", "pass 1", "f64 stream", "pass 2b", "contractions", "pass 2+3", "Jacobian LU", "Slater build",
         "Slater LU", "accept+push", "-", "-"]
for it in range(2):
    xx = x.clone()
    ws = logp.workspace(B, c["n"])
    ws[4096 + 128:4096 + 128 + 96].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cg.mcmc(bound, xx, 1 + it, S, 0.1); e1.record()
    torch.cuda.synchronize()
prof = ws[4096 + 128:4096 + 128 + 96].view(torch.int64).cpu().numpy().astype(np.float64)
evals = B * (S + 1)
print("%s B=%d S=%d kernel=%s: %.3f ms; cycles per evaluation and warp (sum over phases = %.0f)" % (tag, B, S, kern, e0.elapsed_time(e1), prof.sum() / evals))
for nm, v in zip(names, prof):
    if v: print("  %-20s %9.0f cycles/eval  %5.1f %%" % (nm, v / evals, 100 * v / prof.sum()))
