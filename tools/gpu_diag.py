#!/usr/bin/env python
"""First-contact diagnostics on the GPU box: per-case error tables for Ewald / logp / Jacobian / sweep, plus
FFMA / MUFU / DFMA micro-benchmarks.  Writes gpurun_out/diag.txt."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case, CASES_2D
from oracle import logpsi as ol

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); out.write(s + "\n"); out.flush()

P(torch.cuda.get_device_name(0))
for tag in CASES_2D:
    c = load_case(tag)
    try:
        ref = c["ref_potential_energy"]
        V = cg.potential_energy(c["x"][:ref.shape[0]], c["kappa"], c["G"], c["L"], c["rs"])
        P(tag, "ewald rel err max %.3e" % (np.abs(V - ref) / np.abs(ref)).max())
    except Exception as e:
        P(tag, "ewald FAILED", repr(e))
    try:
        nb = c["ref_logp"].shape[0]
        logp = cg.make_logp(c["ktable"], c["L"])
        lp, parts, jac = logp(c["x"][:nb], c["params"], c["state_idx"][:nb], return_parts=True, return_jac=True)
        nj = c["ref_jac"].shape[0]
        ldj = np.array([np.linalg.slogdet(j)[1] for j in c["ref_jac"]])
        P(tag, "logp abs err max %.3e rel %.3e | jac err %.3e | ldJ err %.3e | ldD err %.3e" % (
            np.abs(lp - c["ref_logp"]).max(), (np.abs(lp - c["ref_logp"]) / np.abs(c["ref_logp"])).max(),
            np.abs(jac[:nj] - c["ref_jac"]).max(), np.abs(parts[:nj, 1] - ldj).max(),
            np.abs(parts[:, 0] - c["ref_logslater"].real).max()))
        P("   gpu", lp[:4], "ref", c["ref_logp"][:4])
    except Exception as e:
        P(tag, "logp FAILED", repr(e))
    if "mc_x0" in c:
        try:
            nb = c["mc_x0"].shape[0]
            x, acc, tr = cg.mcmc(logp.bind(c["params"], c["state_idx"][:nb]), c["mc_x0"].copy(),
                                 cg.ExplicitNoise(c["mc_noise"], c["mc_uniforms"]), c["mc_steps"], c["mc_stddev"],
                                 return_trace=True)
            P(tag, "mcmc x err %.3e acc %.4f ref %.4f" % (np.abs(x - c["ref_mc_x"]).max(), acc, c["ref_mc_accept_rate"]))
        except Exception as e:
            P(tag, "mcmc FAILED", repr(e))

# micro-benchmarks
blocks = 148 * 8
for kind, name, per in ((0, "FFMA TFLOP/s", 2 * 8), (1, "MUFU.EX2 Top/s", 8), (2, "DFMA TFLOP/s", 2 * 8)):
    iters = 20000 if kind != 2 else 4000
    best = min(cg.microbench(kind, blocks, iters) for _ in range(3))
    P("microbench", name, "%.2f" % (per * iters * blocks * 256 / best / 1e12), "time %.3f ms" % (best * 1e3))

# quick timing of the sweep at the headline shape
c = load_case("n29_rs1")
logp = cg.make_logp(c["ktable"], c["L"])
for B in (256, 2048, 8192):
    reps = (B + 255) // 256
    x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
    si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
    bound = logp.bind(c["params"], si)
    for _ in range(2):
        cg.mcmc(bound, x.clone(), 1, 50, 0.1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xx = x.clone()
    e0.record(); _, acc = cg.mcmc(bound, xx, 2, 50, 0.1); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    P("sweep n=29 B=%d: %.3f ms, %.3e walker-steps/s, acc %.3f" % (B, ms, B * 50 / ms * 1e3, float(acc)))
    e0.record(); V = cg.potential_energy(xx, 10, c["G"], c["L"], 1.0); e1.record(); torch.cuda.synchronize()
    P("ewald n=29 B=%d: %.3f ms, %.3e evals/s" % (B, e0.elapsed_time(e1), B / e0.elapsed_time(e1) * 1e3))
