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
    cg.set_sweep_kernel("cta")
    cg.set_debug(segments=True)                                    # chains cut into ticketed segments (flag hand-over through global memory)
    xs, acc = cg.mcmc(logp.bind(c["params"], si), x.copy(), 5, 5, 0.1, wrap_L=c["L"])
    cg.set_debug(segments=None)
    cg.set_sweep_kernel(None)
    cg.set_debug(force_pivot=True)
    lp2 = logp(x, c["params"], si)
    cg.set_debug(force_pivot=False, force_f64_slater=True)         # float64 Slater fallback of the warp kernel
    cg.set_sweep_kernel("warp")
    lp3 = logp(x, c["params"], si)
    cg.set_debug(force_f64_slater=False)
    cg.set_sweep_kernel(None)
    V = cg.potential_energy(x, c["kappa"], c["G"], c["L"], c["rs"])
    # fused sampling step (prologue + sweep + Ewald with the statistics epilogue, device key), two-warps-per-walker Ewald
    xd = torch.as_tensor(x, device="cuda")
    st = cg.SamplingStep(logp, torch.as_tensor(si, device="cuda"), torch.as_tensor(cg.pack_params(c["params"]), device="cuda"),
                         xd, 7, 3, 0.1, c["kappa"], cg.kpoints(2, 15), c["rs"], 0.0)
    st.step(); print(tag, st.summary())
    cg.debug_rng(3, 4, 2, c["n"], kat=np.zeros((2, 6), np.uint32))
    print(tag, lp[:2], lp2[:2], acc, V[:2])
torch.cuda.synchronize()
print("sanitize driver done")
