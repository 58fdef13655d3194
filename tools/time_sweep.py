#!/usr/bin/env python
"""Quick A/B timing of one sweep shape (CUDA events, L2 flushed between launches): python tools/time_sweep.py [case] [B] [S] [kernel] [reps]
CGB_LIB selects a variant build."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

tag = sys.argv[1] if len(sys.argv) > 1 else "n29_rs1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
S = int(sys.argv[3]) if len(sys.argv) > 3 else 50
kern = sys.argv[4] if len(sys.argv) > 4 else "auto"
reps_t = int(sys.argv[5]) if len(sys.argv) > 5 else 7
c = load_case(tag)
reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
rng = np.random.default_rng(0)
x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B] + 0.05 * rng.standard_normal((B, c["n"], 2)), device="cuda")
si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
logp = cg.make_logp(c["ktable"], c["L"])
bound = logp.bind(c["params"], si)
cg.set_sweep_kernel(kern)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(reps_t + 2):
    xx = x.clone(); flush.fill_(it)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _, acc = cg.mcmc(bound, xx, 1 + it, S, 0.1); e1.record()
    torch.cuda.synchronize()
    if it >= 2: ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
fl = cg.flops.flop_sweep(c["n"], S) * B
print("%-28s %s B=%d S=%d %s: %.3f ms (min %.3f)  %.3e walker-steps/s  %.2f TFLOP/s  acc %.3f  fallbacks %d" % (
    os.path.basename(os.environ.get("CGB_LIB", "libcgb200.so")), tag, B, S, kern, ms, min(ts), B * S / ms * 1e3, fl / ms / 1e9, float(acc), logp.fallback_count()))
