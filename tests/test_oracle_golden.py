"""Pin the CPU oracle (oracle/) against golden vectors produced by executing the reference's own
source files (tools/make_golden.py), the reference's known answers, and the property tests the
reference's test-suite asserts (tests/test_flow.py, test_slater.py, test_logpsi.py, test_potential.py
of the reference).  CPU only."""
import json
import os
import numpy as np
import pytest
import torch

from conftest import CASES_2D, GOLDEN
from oracle import tables, flow as oflow, logpsi as ologpsi, mcmc as omcmc, potential as opot

RTOL = 1e-10   # float64 restatement vs reference source (numpy LAPACK both sides)


@pytest.mark.parametrize("tag", CASES_2D)
def test_tables_match_reference(cases, tag):
    c = cases(tag)
    kt, Es = tables.ktable(c["dim"], c["Emax"], c["twist"])
    assert np.array_equal(kt, c["ktable"])                      # src/orbitals.py + main.py:90
    assert np.allclose(Es, c["Es_twist"], rtol=0, atol=0)
    G = tables.kpoints(2, c["Gmax"])
    assert np.array_equal(G, c["G"])                            # src/potential.py:7-17
    assert abs(tables.madelung(2, c["kappa"], G) - c["ref_madelung"]) < 1e-13
    assert abs(tables.box_length(c["n"], 2) - c["L"]) < 1e-15


def test_known_answers():
    # SURVEY §8(c)(3), App. B.2; BASELINE.md §2
    G = tables.kpoints(2, 15)
    assert G.shape == (708, 2)
    assert abs(tables.madelung(2, 10, G) - (-3.9002649201)) < 1e-9
    for Emax, M in ((25, 81), (36, 113), (49, 149)):
        assert tables.sp_orbitals(2, Emax)[0].shape == (M, 2)
    for dim in (2, 3):                                          # reference tests/test_potential.py:9-16
        assert tables.kpoints(dim, 4).shape[1] == dim


@pytest.mark.parametrize("tag", CASES_2D)
def test_flow_logpsi_match_reference(cases, tag):
    c = cases(tag)
    nb = c["ref_logpsi"].shape[0]
    out, parts = ologpsi.logpsi_batch(c["x"][:nb], c["params"], c["state_idx"][:nb], c["ktable"], c["L"])
    assert np.allclose(parts["z"], c["ref_z"], rtol=RTOL, atol=1e-11)
    assert np.allclose(out[:, 0], c["ref_logpsi"][:, 0], rtol=RTOL, atol=1e-10)
    dphi = np.angle(np.exp(1j * (out[:, 1] - c["ref_logpsi"][:, 1])))
    assert np.abs(dphi).max() < 1e-9
    nj = c["ref_jac"].shape[0]
    assert np.abs(parts["jac"][:nj] - c["ref_jac"]).max() < 1e-11   # torch jacfwd vs complex-step
    lp = ologpsi.logp_batch(c["x"][:nb], c["params"], c["state_idx"][:nb], c["ktable"], c["L"])
    assert np.allclose(lp, c["ref_logp"], rtol=RTOL, atol=1e-10)
    assert np.allclose(parts["logabsdet_slater"], c["ref_logslater"].real, rtol=RTOL, atol=1e-10)


def test_general_depth_dim_match_reference():
    """n=7, dim=3, depth=3 — the shape the reference's own tests use (tests/test_logpsi.py:30-31)."""
    from conftest import load_case
    c = load_case("n7_dim3_depth3")
    out, parts = ologpsi.logpsi_batch(c["x"], c["params"], c["state_idx"], c["ktable"], c["L"])
    assert np.allclose(parts["z"], c["ref_z"], rtol=RTOL, atol=1e-12)
    assert np.allclose(out[:, 0], c["ref_logpsi"][:, 0], rtol=RTOL, atol=1e-10)


@pytest.mark.parametrize("tag", CASES_2D)
def test_ewald_matches_reference(cases, tag):
    c = cases(tag)
    ne = min(32, c["ref_potential_energy"].shape[0])
    V = opot.potential_energy(c["x"][:ne], c["kappa"], c["G"], c["L"], c["rs"])
    assert np.allclose(V, c["ref_potential_energy"][:ne], rtol=1e-12, atol=0)


@pytest.mark.parametrize("tag", ["n13_rs10", "n29_rs1", "n49_rs0.5", "n57_rs0.25"])
def test_mcmc_matches_reference(cases, tag):
    c = cases(tag)
    nb = c["mc_x0"].shape[0]
    lp = lambda x: ologpsi.logp_batch(x, c["params"], c["state_idx"][:nb], c["ktable"], c["L"])
    x, acc = omcmc.mcmc(lp, c["mc_x0"], c["mc_noise"], c["mc_uniforms"], c["mc_steps"], c["mc_stddev"])
    assert np.array_equal(x, c["ref_mc_x"])            # identical decisions -> bit-identical walkers
    assert acc == c["ref_mc_accept_rate"]
    # the pmapped wrapper incl. the final wrap into [0,L)  (src/VMC.py:23-24)
    for dev in range(2):
        h = nb // 2
        lpd = lambda x: ologpsi.logp_batch(x, c["params"], c["state_idx"][dev * h:(dev + 1) * h],
                                           c["ktable"], c["L"])
        xd, accd = omcmc.mcmc(lpd, c["vmc_x0"][dev], c["vmc_noise"][dev], c["vmc_uniforms"][dev],
                              c["mc_steps"], c["mc_stddev"])
        assert np.allclose(omcmc.wrap(xd, c["L"]), c["ref_vmc_x"][dev], rtol=0, atol=1e-9)
        assert abs(accd - c["ref_vmc_accept_rate"][dev]) < 1e-15


def test_statistical_pins_from_shipped_logs(cases):
    """V/rs^2 recomputed from the shipped walkers agrees with the last data.txt row."""
    pins = json.load(open(os.path.join(GOLDEN, "data_txt_pins.json")))
    for tag, key in (("n29_rs1", "n29_rs1"), ("n29_rs10", "n29_rs10"), ("n49_rs0.5", "n49_rs0.5"),
                     ("n57_rs0.25", "n57_rs0.25")):
        c = cases(tag)
        V = c["ref_potential_energy"] + c["n"] * c["rs"] / c["L"] * c["ref_madelung"]
        v = V / c["rs"] ** 2
        err = np.hypot(v.std() / np.sqrt(v.size), pins[key]["V_std"])
        assert abs(v.mean() - pins[key]["V"]) < 4 * err, (tag, v.mean(), pins[key]["V"], err)


# ---- property tests mirrored from the reference's suite (float64, jnp.allclose defaults) ----
def _rand_case(n=7, dim=3, depth=3, L=1.234, seed=0):
    rng = np.random.default_rng(seed)
    p = oflow.init_params(depth, 16, 16, dim, seed=seed, stddev=0.3)
    return rng, p, rng.uniform(0, L, (n, dim)), L


def test_flow_equivariances():
    """reference tests/test_flow.py:19-38."""
    rng, p, x, L = _rand_case()
    pt = oflow.to_torch(p)
    f = lambda x: oflow.flow(pt, torch.as_tensor(x), L).numpy()
    z = f(x)
    image = rng.integers(-5, 6, size=x.shape) * L
    assert np.allclose(f(x + image), z + image)
    shift = rng.normal(size=x.shape[1])
    assert np.allclose(f(x + shift), z + shift)
    P = rng.permutation(x.shape[0])
    assert np.allclose(f(x[P]), z[P])


def test_slater_antisymmetry_and_translation():
    """reference tests/test_slater.py:27-37."""
    rng = np.random.default_rng(1)
    n, dim, L = 7, 3, 1.234
    sp, _ = tables.sp_orbitals(dim)
    ind = torch.as_tensor(sp[rng.choice(sp.shape[0], n, replace=False)].astype(np.float64))
    x = rng.normal(size=(n, dim))
    det = np.exp(ologpsi.logslaterdet(ind, torch.as_tensor(x), L).numpy())
    P = rng.permutation(n)
    Pdet = np.exp(ologpsi.logslaterdet(ind, torch.as_tensor(x[P]), L).numpy())
    assert np.allclose(Pdet, det) or np.allclose(Pdet, -det)
    shift = rng.normal(size=dim)
    sdet = np.exp(ologpsi.logslaterdet(ind, torch.as_tensor(x + shift), L).numpy())
    phase = np.exp(1j * (2 * np.pi / L * ind.numpy()).dot(shift).sum())
    assert np.allclose(sdet, phase * det)


def test_logp_invariances():
    """reference tests/test_logpsi.py:40-77 (PBC images, global shift, permutation +-1)."""
    rng, p, x, L = _rand_case(seed=3)
    sp, _ = tables.sp_orbitals(3)
    kt = sp.astype(np.float64)
    si = np.sort(rng.choice(sp.shape[0], x.shape[0], replace=False))[None]
    lp = lambda x: ologpsi.logp_batch(x[None], p, si, kt, L)[0]
    base = lp(x)
    image = rng.integers(-5, 6, size=x.shape) * L
    assert np.allclose(lp(x + image), base)
    assert np.allclose(lp(x + rng.normal(size=3)), base)
    assert np.allclose(lp(x[rng.permutation(x.shape[0])]), base)
