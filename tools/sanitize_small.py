#!/usr/bin/env python
"""Small-shape driver for compute-sanitizer (memcheck / racecheck / synccheck): every kernel, every NW variant once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case
for tag, B in (("n13_rs10", 3), ("n29_rs1", 3), ("n49_rs0.5", 2), ("n57_rs0.25", 2)):
    c = load_case(tag)
    logp = cg.make_logp(c["ktable"], c["L"])
    x, si = c["x"][:B].copy(), c["state_idx"][:B]
    for kern in ("cta", "warp", "spec4", "spec8", "cta4", "cta8"):     # both sweep kernels and their speculative modes
        cg.set_sweep_kernel(kern)
        lp = logp(x, c["params"], si)
        xs, acc = cg.mcmc(logp.bind(c["params"], si), x.copy(), 5, 4, 0.1)
    cg.set_sweep_kernel(None)
    cg.set_debug(force_pivot=True)
    lp2 = logp(x, c["params"], si)
    cg.set_debug(force_pivot=False)
    V = cg.potential_energy(x, c["kappa"], c["G"], c["L"], c["rs"])
    print(tag, lp[:2], lp2[:2], acc, V[:2])
torch.cuda.synchronize()
print("sanitize driver done")
