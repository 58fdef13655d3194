import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case
c = load_case("n29_rs1"); B = 2048; S = 6
reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
x0 = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
params = torch.as_tensor(cg.pack_params(c["params"]), device="cuda")
G = cg.kpoints(2, 15)
logp = cg.make_logp(c["ktable"], c["L"])
bound = logp.bind(params, si)
key = cg.device_key(5)
stats = torch.zeros(8, dtype=torch.float64, device="cuda")
def try_capture(name, fn):
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "capture OK")
    except Exception as e:
        print(name, "capture FAILED:", str(e).splitlines()[0])
x = x0.clone()
try_capture("ewald", lambda: cg.potential_energy(x, 10, G, c["L"], c["rs"]))
try_capture("ewald+stats", lambda: cg.potential_energy(x, 10, G, c["L"], c["rs"], stats=stats))
try_capture("logp", lambda: logp(x, params, si))
try_capture("mcmc hostkey", lambda: cg.mcmc(bound, x, 7, S, 0.1, _want_rate=False))
try_capture("mcmc devkey", lambda: cg.mcmc(bound, x, key, S, 0.1, advance_key=True, stats=stats, _want_rate=False))
cg.set_sweep_kernel("cta")
try_capture("mcmc cta", lambda: cg.mcmc(bound, x, 7, S, 0.1, _want_rate=False))
