#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES
(/root/reference/src/{orbitals,flow,slater,logpsi,MCMC,potential,VMC}.py, unmodified)
under the NumPy-backed jax/haiku stand-in in oracle/jaxshim, on the reference's
shipped checkpoints (walkers + trained flow parameters).

Run in the build container only (needs /root/reference):
    python tools/make_golden.py
The outputs are small and committed; nothing at test/bench time reads /root/reference.
"""
import glob
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=SyntaxWarning)

from oracle.jaxshim import load_reference  # noqa: E402
from oracle import checkpoint  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
ref = load_reference(REF)
import jax  # noqa: E402  (the stand-in)
import haiku as hk  # noqa: E402


def setup(n, dim, Emax, twist):
    """main.py:63-90 with the reference's own orbitals.py."""
    L = float(np.sqrt(np.pi * n)) if dim == 2 else float((4 / 3 * np.pi * n) ** (1 / 3))
    sp_indices, Es = ref["orbitals"].sp_orbitals(dim, Emax)
    idx_t, Es_t = ref["orbitals"].twist_sort(sp_indices, np.array(twist))
    return L, np.ascontiguousarray(idx_t[::-1]), np.ascontiguousarray(Es_t[::-1])


def make_state_idx(B, n, M, seed=43):
    """even walkers ground state [M-n..M-1]; odd walkers a sorted random subset."""
    rng = np.random.default_rng(seed)
    si = np.empty((B, n), dtype=np.int32)
    for b in range(B):
        si[b] = np.arange(M - n, M) if b % 2 == 0 else np.sort(rng.choice(M, size=n, replace=False))
    return si


def ref_case(tag, n, dim, Emax, rs, params, x_all, depth=2, n_logpsi=8, n_jac=2, n_mc=4, mc_steps=6,
             kappa=10, Gmax=15, twist=(0.25, 0.25), n_ewald=None):
    L, ktable, Es = setup(n, dim, Emax, twist)
    M = ktable.shape[0]
    flow = hk.transform(lambda x: ref["flow"].FermiNet(depth, 16, 16, L)(x))
    logpsi = ref["logpsi"].make_logpsi(flow, ktable, L)
    logp = ref["logpsi"].make_logp(logpsi)
    state_idx = make_state_idx(x_all.shape[0], n, M)

    out = dict(n=n, dim=dim, Emax=Emax, rs=rs, L=L, depth=depth, ktable=ktable, Es_twist=Es,
               x=x_all, state_idx=state_idx, kappa=kappa, Gmax=Gmax, twist=np.array(twist))
    for k, v in params.items():
        out["param:%s:w" % k] = v["w"]
        out["param:%s:b" % k] = v["b"]

    # --- flow, logpsi, jacobian (src/flow.py, src/logpsi.py, src/slater.py)
    xs = x_all[:n_logpsi]
    out["ref_z"] = np.stack([flow.apply(params, None, x) for x in xs])
    out["ref_logpsi"] = np.stack([logpsi(x, params, si) for x, si in zip(xs, state_idx)])
    out["ref_logp"] = logp(xs, params, state_idx[:n_logpsi])
    flat = lambda x: flow.apply(params, None, x.reshape(n, dim)).reshape(-1)
    out["ref_jac"] = np.stack([jax.jacfwd(flat)(x.reshape(-1)) for x in xs[:n_jac]])
    out["ref_logslater"] = np.array([ref["slater"].logslaterdet(ktable[si], z, L)
                                     for si, z in zip(state_idx, out["ref_z"])])

    # --- Ewald (src/potential.py)
    if dim == 2:
        G = np.asarray(ref["potential"].kpoints(dim, Gmax))
        out["G"] = G
        out["ref_madelung"] = float(ref["potential"].Madelung(dim, kappa, G))
        ne = x_all.shape[0] if n_ewald is None else n_ewald
        out["ref_potential_energy"] = ref["potential"].potential_energy(x_all[:ne], kappa, G, L, rs)

    # --- Metropolis (src/MCMC.py) and the sampling-step wrapper (src/VMC.py:8-25)
    if n_mc:
        sig = 0.1
        x0 = x_all[:n_mc].copy()
        si0 = state_idx[:n_mc]
        key = jax.random.PRNGKey(44)
        jax.random.log.clear()
        x_fin, acc = ref["MCMC"].mcmc(lambda x: logp(x, params, si0), x0.copy(), key, mc_steps, sig)
        draws = list(jax.random.log)
        out["mc_x0"], out["mc_steps"], out["mc_stddev"] = x0, mc_steps, sig
        out["mc_noise"] = np.stack([v for k, v in draws if k == "normal"])
        out["mc_uniforms"] = np.stack([v for k, v in draws if k == "uniform"])
        out["ref_mc_x"], out["ref_mc_accept_rate"] = x_fin, float(acc)

        # VMC.py wrapper with 2 "devices" of n_mc/2 walkers and a stub sampler returning fixed states
        if n_mc % 2 == 0:
            h = n_mc // 2
            keys = np.stack([jax.random.PRNGKey(7), jax.random.PRNGKey(8)])
            xdev = x_all[:n_mc].reshape(2, h, n, dim).copy() + 3.0 * L   # off-cell so the wrap matters
            call = [0]

            def sampler(params_van, key_state, batch):
                s = state_idx[call[0] * h:(call[0] + 1) * h]
                call[0] += 1
                return s
            jax.random.log.clear()
            keys_out, si_out, x_out, acc_out = ref["VMC"].sample_stateindices_and_x(
                keys, sampler, None, logp, xdev.copy(),
                jax.tree_map(lambda a: np.stack([a, a]), params),      # src/utils.py:6-8 replicate
                mc_steps, sig, L)
            draws = list(jax.random.log)
            out["vmc_x0"] = xdev
            out["vmc_noise"] = np.stack([v for k, v in draws if k == "normal"]).reshape(2, mc_steps, h, n, dim)
            out["vmc_uniforms"] = np.stack([v for k, v in draws if k == "uniform"]).reshape(2, mc_steps, h)
            out["ref_vmc_x"], out["ref_vmc_accept_rate"] = x_out, np.asarray(acc_out)
    path = os.path.join(OUT, "ref_%s.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB",
          "logp[0]=%.10f" % out["ref_logp"][0], flush=True)


def ckpt(pattern):
    f = glob.glob(os.path.join(REF, "data", pattern, "epoch_*.pkl"))
    assert len(f) == 1, (pattern, f)
    return checkpoint.load_data(f[0])


def main():
    os.makedirs(OUT, exist_ok=True)
    # C3: n=29 rs=1 (config 3) — device-0 shard of the shipped checkpoint (256 walkers)
    d = ckpt("n_29_dim_2_rs_1.0_*")
    ref_case("n29_rs1", 29, 2, 25, 1.0, d["params_flow"], d["x"][0], n_logpsi=16)
    d = ckpt("n_29_dim_2_rs_10.0_*")
    ref_case("n29_rs10", 29, 2, 25, 10.0, d["params_flow"], d["x"][0][:64], n_logpsi=8, n_mc=0)
    # C4 / C5
    d = ckpt("n_49_dim_2_rs_0.5_*")
    ref_case("n49_rs0.5", 49, 2, 36, 0.5, d["params_flow"], d["x"][0], n_logpsi=6, n_mc=16, mc_steps=10)
    d = ckpt("n_57_dim_2_rs_0.25_*")
    ref_case("n57_rs0.25", 57, 2, 49, 0.25, d["params_flow"], d["x"][0], n_logpsi=6, n_mc=16, mc_steps=10)
    # C1/C2: n=13 — no shipped flow checkpoint; haiku init (RandomNormal(0.01), zero bias) via the
    # stand-in, scaled up x30 so the flow is not a near-identity, walkers uniform in [0,L)
    n = 13
    L = float(np.sqrt(np.pi * n))
    flow = hk.transform(lambda x: ref["flow"].FermiNet(2, 16, 16, L)(x))
    rng = np.random.default_rng(42)
    x13 = rng.uniform(0, L, (64, n, 2))
    p13 = flow.init(jax.random.PRNGKey(13), x13[0])
    p13 = {k: {"w": v["w"] * 30.0, "b": rng.normal(0, 0.1, v["b"].shape)} for k, v in p13.items()}
    ref_case("n13_rs10", 13, 2, 25, 10.0, p13, x13, n_logpsi=16, n_mc=8, mc_steps=10)
    # the reference tests' own shape: n=7, dim=3, depth=3, L=1.234 (tests/test_logpsi.py:30-31)
    n, L = 7, 1.234
    x7 = rng.uniform(0, L, (4, n, 3))
    flow3 = hk.transform(lambda x: ref["flow"].FermiNet(3, 16, 16, L)(x))
    p7 = flow3.init(jax.random.PRNGKey(7), x7[0])
    p7 = {k: {"w": v["w"] * 30.0, "b": rng.normal(0, 0.1, v["b"].shape)} for k, v in p7.items()}
    sp_indices, _ = ref["orbitals"].sp_orbitals(3)
    logpsi = ref["logpsi"].make_logpsi(flow3, sp_indices.astype(np.float64), L)
    si = np.stack([np.sort(rng.choice(sp_indices.shape[0], n, replace=False)) for _ in range(4)])
    out = dict(n=n, dim=3, L=L, depth=3, x=x7, state_idx=si, ktable=sp_indices.astype(np.float64),
               ref_logpsi=np.stack([logpsi(x, p7, s) for x, s in zip(x7, si)]),
               ref_z=np.stack([flow3.apply(p7, None, x) for x in x7]))
    for k, v in p7.items():
        out["param:%s:w" % k], out["param:%s:b" % k] = v["w"], v["b"]
    np.savez_compressed(os.path.join(OUT, "ref_n7_dim3_depth3.npz"), **out)

    # statistical pins: last row of each shipped data.txt (main.py:369-374)
    pins = {}
    for run in sorted(glob.glob(os.path.join(REF, "data", "n_*"))):
        last = open(os.path.join(run, "data.txt")).read().strip().splitlines()[-1].split()
        name = os.path.basename(run)
        nn = int(name.split("_")[1]); rs = float(name.split("_rs_")[1].split("_")[0])
        pins["n%d_rs%g" % (nn, rs)] = dict(epoch=int(last[0]), V=float(last[7]), V_std=float(last[8]),
                                           accept_rate=float(last[11]), run=name)
    json.dump(pins, open(os.path.join(OUT, "data_txt_pins.json"), "w"), indent=1)
    print("done")


if __name__ == "__main__":
    main()
