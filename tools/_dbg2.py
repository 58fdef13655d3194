import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case
c = load_case("n29_rs1"); S = 50
G = cg.kpoints(2, 15)
logp = cg.make_logp(c["ktable"], c["L"])
params = torch.as_tensor(cg.pack_params(c["params"]), device="cuda")
def mk(B, seed):
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda").clone()
    si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
    st = cg.SamplingStep(logp, si, params, x, cg.device_key(seed), S, 0.1, 10, G, c["rs"], 0.0)
    st.capture(warmup=2)
    return st, si
s1, si1 = mk(2048, 1)
for _ in range(3): s1.step()
torch.cuda.synchronize(); print("graph 1 ok", s1.summary())
s2, si2 = mk(1024, 2)
for _ in range(3): s2.step()
torch.cuda.synchronize(); print("graph 2 ok", s2.summary())
xk = s1.x.clone()
try:
    cg.mcmc(logp.bind(params, si1), xk, cg.device_key(5), S, 0.1, wrap_L=c["L"], advance_key=True, _want_rate=False)
    torch.cuda.synchronize(); print("eager ok")
except Exception as e:
    print("eager FAILED", e)
