#!/usr/bin/env python
"""Tiny driver for ncu: one warm-up + one measured launch of the sweep / logp / Ewald kernels at a given shape.
usage: python tools/profile_sweep.py [case] [B] [mc_steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

tag = sys.argv[1] if len(sys.argv) > 1 else "n29_rs1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
S = int(sys.argv[3]) if len(sys.argv) > 3 else 10
if len(sys.argv) > 4: cg.set_sweep_kernel(sys.argv[4])
c = load_case(tag)
reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
logp = cg.make_logp(c["ktable"], c["L"])
bound = logp.bind(c["params"], si)
for it in range(2):
    xx = x.clone()
    _, acc = cg.mcmc(bound, xx, 1 + it, S, 0.1)
    V = cg.potential_energy(xx, c["kappa"], c["G"], c["L"], c["rs"])
torch.cuda.synchronize()
print("ok", float(acc), float(V.mean()))
