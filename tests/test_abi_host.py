"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/cgb200.h
declares, rejects bad arguments without touching a GPU, and its host-side helpers agree with NumPy."""
import ctypes
import os
import re
import numpy as np
import pytest

from conftest import ROOT
import coulombgas_b200 as cg
from coulombgas_b200 import _lib
from oracle import tables


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "cgb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cgb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.cgb_abi_version() == 2


def test_bad_arguments_return_codes():
    lib = _lib.lib
    tw = (ctypes.c_double * 2)(0.25, 0.25)
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)      # non-null dummy; never dereferenced because validation fails first
    assert lib.cgb_logp_f64io(null, null, one, one, 81, tw, one, 4, 13, 6.39, one, null, null, one, 8192, 0) == -1
    assert lib.cgb_logp_f64io(null, one, one, one, 81, tw, one, 4, 65, 6.39, one, null, null, one, 8192, 0) == -2   # n > 64
    assert lib.cgb_logp_f64io(null, one, one, one, 81, tw, one, 4, 13, 6.39, one, null, null, null, 0, 0) == -3     # workspace
    assert lib.cgb_logp_f64io(null, one, one, one, 81, tw, one, 4, 13, 6.39, one, null, null, one, 1 << 20, 9) == -1   # kernel id
    sweep = lambda noise, unif, key, flags: lib.cgb_mcmc_sweep_f64io(
        null, one, one, one, 81, tw, one, noise, unif, 0, key, 5, 0.1, 6.39, 4, 13, 1, one, one, null, null, null, one, 8192, flags)
    assert sweep(one, null, null, 0) == -1                               # noise without uniforms
    assert sweep(null, null, null, _lib.FLAG_ADVANCE_KEY) == -1          # advancing needs a device key
    assert sweep(null, null, ctypes.c_void_p(18), 0) == -1               # misaligned device key
    assert sweep(null, null, null, 0) == -3                              # valid arguments, workspace too small for B = 4
    assert lib.cgb_ewald_f64io(null, one, 4, 13, 6.39, 10.0, 10.0, null, 10, 15, one, null) == -1
    assert lib.cgb_van_sample_f64io(null, one, 18 * 81 + 4368, one, 81, 13, 4, null, 0, null, null, null, null) == -1
    assert lib.cgb_debug_rng(null, 0, null, 4, 4, 13, null, null, null, null, 3) == -1
    assert lib.cgb_microbench(null, 7, 1, 1, one) == -1
    assert b"unsupported" in lib.cgb_error_string(-2)
    nb = ctypes.c_size_t()
    assert lib.cgb_logp_workspace_bytes(256, 29, ctypes.byref(nb)) == 0 and nb.value >= 4032


def test_ktable_twist_helper():
    kt, _ = tables.ktable(2, 25, (0.25, 0.25))
    lg_tw = (ctypes.c_double * 2)()
    assert _lib.lib.cgb_ktable_twist_host(kt.ctypes.data_as(ctypes.c_void_p), kt.shape[0], lg_tw) == 0
    assert abs(lg_tw[0] - 0.25) < 1e-15 and abs(lg_tw[1] - 0.25) < 1e-15
    bad = kt.copy(); bad[3, 0] += 0.1
    assert _lib.lib.cgb_ktable_twist_host(bad.ctypes.data_as(ctypes.c_void_p), bad.shape[0], lg_tw) == -2
    with pytest.raises(cg.CgbError):
        cg.make_logp(bad, 6.39)


def test_ewald_item_table_reproduces_pairwise_sum():
    """The merged (gx>=0, gy>=0) item table must give the reference's pairwise-cos reciprocal sum
    (src/potential.py:61) through the structure-factor identity — evaluated here in NumPy float64."""
    G = tables.kpoints(2, 15)
    kappa = 10
    items = np.empty((G.shape[0], 4), dtype=np.float32)
    n_items, gmax = ctypes.c_int(), ctypes.c_int()
    Gi = np.ascontiguousarray(G, dtype=np.int64)
    assert _lib.lib.cgb_ewald_prepare_host(Gi.ctypes.data_as(ctypes.c_void_p), G.shape[0], float(kappa),
                                           items.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n_items),
                                           ctypes.byref(gmax)) == 0
    items = items[:n_items.value].astype(np.float64)
    assert gmax.value == 15 and n_items.value < G.shape[0] // 2 + 32
    g_k, _ = tables.ewald_gk(2, kappa, G)
    assert abs(items[:, 2:].sum() - g_k.sum()) < 1e-6 * g_k.sum()
    rng = np.random.default_rng(5)
    n = 11
    u = rng.random((n, 2))
    i, j = np.triu_indices(n, k=1)
    rij = (u[:, None] - u)[i, j]
    ref = (g_k * np.cos(2 * np.pi * G.dot(rij.T)).sum(-1)).sum()
    ex, ey = np.exp(2j * np.pi * u[:, 0]), np.exp(2j * np.pi * u[:, 1])
    tot = 0.0
    for gx, gy, wp, wm in items:
        sp = (ex ** gx * ey ** gy).sum()
        sm = (ex ** gx * ey ** (-gy)).sum()
        tot += 0.5 * wp * (abs(sp) ** 2 - n) + 0.5 * wm * (abs(sm) ** 2 - n)
    assert abs(tot - ref) < 1e-6 * abs(ref)


def test_host_tables_match_oracle():
    assert np.array_equal(cg.kpoints(2, 15), tables.kpoints(2, 15))
    assert abs(cg.madelung(2, 10, cg.kpoints(2, 15)) - tables.madelung(2, 10, tables.kpoints(2, 15))) < 1e-12


def test_pack_params_layout(cases):
    c = cases("n29_rs1")
    flat = cg.pack_params(c["params"])
    assert flat.shape == (cg.NPARAMS,)
    assert np.array_equal(flat[:80].reshape(5, 16), c["params"]["fermi_net/~/linear_2"]["w"])
    assert np.array_equal(flat[-2:], c["params"]["fermi_net/linear"]["b"])
    with pytest.raises(ValueError):
        cg.pack_params({"x": {}})


def test_kernel_model_matches_reference(cases):
    """The algorithm the kernels implement (analytic Jacobian, twist-free Slater matrix, common-mode removal;
    tests/kernel_model.py) equals the as-written reference in float64, and its float32 arithmetic stays
    inside the 1e-5 relative budget on the shipped walkers."""
    from kernel_model import logp_model, pack_params, split_ktable
    for tag in ("n13_rs10", "n29_rs1", "n57_rs0.25"):
        c = cases(tag)
        fp = pack_params(c["params"])
        m, tw = split_ktable(c["ktable"])
        for b in range(3):
            lp, parts = logp_model(c["x"][b], fp, m[c["state_idx"][b]], c["L"], np.float64, True)
            assert abs(lp - c["ref_logp"][b]) < 1e-10
            if b < c["ref_jac"].shape[0]:
                assert np.abs(parts["J"] - c["ref_jac"][b]).max() < 1e-12
            lp32 = logp_model(c["x"][b], fp, m[c["state_idx"][b]], c["L"], np.float32)
            assert abs(lp32 - c["ref_logp"][b]) < 1e-5 * abs(c["ref_logp"][b])
