"""Interop with the reference's on-disk formats and setup tables (coulombgas_b200/interop.py; SURVEY §8 a9/a10, f/N1).
CPU only.  Golden data: the k-tables in tests/golden/ref_*.npz were produced by the reference's own orbitals.py
(tools/make_golden.py); run-directory names and data.txt rows are the shipped ones (tests/golden/data_txt_pins.json)."""
import json
import os
import pickle
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN, CASES_2D
from coulombgas_b200 import interop


@pytest.mark.parametrize("tag", CASES_2D)
def test_sp_indices_twist_matches_reference_tables(cases, tag):
    c = cases(tag)
    kt, es = interop.sp_indices_twist(2, c["Emax"], tuple(c["twist"]))
    assert kt.dtype == np.float64 and kt.flags["C_CONTIGUOUS"]
    assert np.array_equal(kt, c["ktable"])              # bit-exact, including the order of degenerate orbitals
    assert np.array_equal(es, c["Es_twist"])
    assert abs(interop.box_length(c["n"], 2) - c["L"]) < 1e-15
    gs = interop.ground_state_indices(kt.shape[0], c["n"])
    assert np.array_equal(gs, c["state_idx"][0])        # even walkers of the fixtures hold the ground state


def test_sp_orbitals_counts_and_order():
    for Emax, M in ((25, 81), (36, 113), (49, 149)):
        idx, es = interop.sp_orbitals(2, Emax)
        assert idx.shape == (M, 2) and np.all(np.diff(es) >= 0) and es[-1] <= Emax
        assert len({tuple(r) for r in idx}) == M
    idx3, es3 = interop.sp_orbitals(3, 4)
    assert idx3.shape[1] == 3 and np.all((idx3 ** 2).sum(1) == es3)


def test_run_dirname_reproduces_shipped_names():
    pins = json.load(open(os.path.join(GOLDEN, "data_txt_pins.json")))
    for tag, pin in pins.items():
        f = interop.parse_run_dirname(pin["run"])
        assert f["dim"] == 2 and f["Gmax"] == 15 and f["kappa"] == 10 and f["twist"] == (0.25, 0.25)
        assert (f["depth"], f["spsize"], f["tpsize"]) == (2, 16, 16)
        assert (f["mc_therm"], f["mc_steps"], f["mc_stddev"]) == (10, 50, 0.1)
        name = interop.run_dirname("", nlayers=2, modelsize=16, nheads=4, nhidden=32, hutchinson=True, sr=True,
                                   damping=1e-3, max_norm=1e-3, **f)
        if float("%.1f" % f["rs"]) == f["rs"]:
            assert name == pin["run"], tag
        else:       # rs=0.25: main.py:194 prints "%.1f" (-> rs_0.2); the shipped directory was named by hand / an older script
            assert name == pin["run"].replace("rs_0.25", "rs_0.2"), tag
    assert interop.ckpt_filename(3000, "/d") == "/d/epoch_003000.pkl"
    assert interop.pretrained_model_filename("/p") == "/p/params_van.pkl"


def test_checkpoint_roundtrip_and_sharding(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.random((8, 4, 5, 2))
    ck = {"keys": rng.integers(0, 2 ** 32, (8, 2), dtype=np.uint32), "x": x,
          "params_van": {"m": {"w": rng.random((3, 3))}}, "params_flow": {"fermi_net/linear": {"w": rng.random((16, 2)), "b": np.zeros(2)}},
          "opt_state": ()}
    fn = interop.ckpt_filename(100, str(tmp_path))
    interop.save_data(ck, fn)
    back = interop.load_data(fn)
    assert set(back) == set(ck)
    assert np.array_equal(back["x"], x) and back["keys"].dtype == np.uint32
    assert np.array_equal(back["params_flow"]["fermi_net/linear"]["w"], ck["params_flow"]["fermi_net/linear"]["w"])
    with open(fn, "rb") as f:                      # the reference reads it with a bare pickle.load (checkpoint.py:10-13)
        assert np.array_equal(pickle.load(f)["x"], x)
    # sharding is a pure partition in the reference's device-major order (main.py:234-238)
    for world in (1, 2, 4, 8, 16):
        parts = [interop.shard_walkers(x, r, world) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), x.reshape(-1, 5, 2))
    assert np.array_equal(interop.gather_walkers([interop.shard_walkers(x, r, 8) for r in range(8)]), x)
    with pytest.raises(ValueError, match="divisible"):
        interop.shard_walkers(x, 0, 3)


def test_load_data_without_jax(tmp_path):
    """A checkpoint written by the reference pickles jax Arrays through jax._src.array._reconstruct_array and optax
    states by class; neither library exists here.  Fabricate such a pickle with throw-away modules, drop the modules,
    and read it back."""
    jmod, omod = types.ModuleType("jax._src.array"), types.ModuleType("optax._src.base")

    def _reconstruct_array(fun, args, arr_state, aval_state):
        raise AssertionError("the real jax function must not be needed")
    _reconstruct_array.__module__ = "jax._src.array"
    _reconstruct_array.__qualname__ = "_reconstruct_array"
    jmod._reconstruct_array = _reconstruct_array

    class EmptyState(tuple):
        pass
    EmptyState.__module__ = "optax._src.base"
    EmptyState.__qualname__ = "EmptyState"
    omod.EmptyState = EmptyState

    class FakeArray:
        def __init__(self, a):
            self.a = a

        def __reduce__(self):
            fun, args, state = self.a.__reduce__()
            return (_reconstruct_array, (fun, args, state, {"weak_type": False}))
    saved = {k: sys.modules.get(k) for k in ("jax", "jax._src", "jax._src.array", "optax", "optax._src", "optax._src.base")}
    try:
        for k in ("jax", "jax._src", "optax", "optax._src"):
            sys.modules[k] = types.ModuleType(k)
        sys.modules["jax._src.array"], sys.modules["optax._src.base"] = jmod, omod
        x = np.arange(24, dtype=np.float64).reshape(2, 2, 3, 2)
        blob = pickle.dumps({"x": FakeArray(x), "keys": FakeArray(np.array([[1, 2]], dtype=np.uint32)),
                             "opt_state": (EmptyState(), EmptyState())})
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    with pytest.raises(Exception):
        pickle.loads(blob)                         # plain pickle needs jax / optax
    d = interop.loads(blob)
    assert isinstance(d["x"], np.ndarray) and np.array_equal(d["x"], x) and d["keys"].dtype == np.uint32
    assert len(d["opt_state"]) == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("tag", ["n29_rs1", "n49_rs0.5", "n57_rs0.25"])
def test_shipped_checkpoints_load(cases, tag):
    pins = json.load(open(os.path.join(GOLDEN, "data_txt_pins.json")))
    run = os.path.join("/root/reference/data", pins[tag]["run"])
    f = interop.parse_run_dirname(run)
    ck = interop.load_data(interop.ckpt_filename(pins[tag]["epoch"], run))
    nd = f["num_devices"]
    assert ck["x"].shape == (nd, f["batch"] // nd, f["n"], 2) and ck["x"].dtype == np.float64
    assert ck["keys"].shape == (nd, 2) and ck["keys"].dtype == np.uint32
    c = cases(tag)
    shard0 = interop.shard_walkers(ck["x"], 0, nd)
    assert np.array_equal(shard0[:c["x"].shape[0]], c["x"])            # the fixtures are the first walkers of device 0
    for mod, leaves in c["params"].items():
        for leaf, v in leaves.items():
            assert np.array_equal(ck["params_flow"][mod][leaf], v)
    log = interop.read_data_txt(os.path.join(run, "data.txt"))
    row = np.where(log["epoch"] == pins[tag]["epoch"])[0][-1]
    assert abs(log["V"][row] - pins[tag]["V"]) < 1e-9 and abs(log["accept_rate"][row] - pins[tag]["accept_rate"]) < 1e-9


def test_data_txt_line_format(tmp_path):
    # the last row of the shipped n=29 rs=1 log (main.py:369-374 format), rebuilt from unscaled values
    want = "  3000  4.388876  0.026395  10.436681  0.034511  61.474450  0.047871  -51.037769  0.038738  10.079675  0.037649  0.5231\n"
    got = interop.format_data_line(3000, F=4.388876, F_std=0.026395, E=10.436681, E_std=0.034511, K=61.474450, K_std=0.047871,
                                   V=-51.037769, V_std=0.038738, S=10.079675, S_std=0.037649, accept_rate=0.5231, rs=1.0)
    assert got == want
    half = interop.format_data_line(7, F=4.0, F_std=0.4, E=8.0, E_std=0.8, K=12.0, K_std=1.2, V=-16.0, V_std=1.6, S=3.0,
                                    S_std=0.3, accept_rate=0.5, rs=2.0)
    p = tmp_path / "data.txt"
    p.write_text(want + half)
    d = interop.read_data_txt(str(p))
    assert list(d["epoch"]) == [3000, 7] and d["V"][1] == -4.0 and d["S"][1] == 3.0 and d["F_std"][1] == 0.1
    assert abs(interop.observable_std(2.0, 5.0, 2048, 4) - np.sqrt(1.0 / 8192)) < 1e-15


def test_shipped_tie_order_tables():
    """interop.sp_indices_twist(order="shipped"): the k-tables with the degenerate-orbital order the shipped runs were trained
    with (tools/recover_tie_order.py).  Same orbitals and energies as today's table row by row; on the shells the n=13 fixture
    pins it agrees with the order recovered independently there; and the recovery report says the shipped parameters reach
    the spread of log p their own runs logged."""
    import json
    path = os.path.join(os.path.dirname(interop.__file__), "data", "shipped_ktables.npz")
    rep = json.loads(str(np.load(path)["report"]))
    for Emax, M in ((25, 81), (36, 113), (49, 149)):
        kt, es = interop.sp_indices_twist(2, Emax, (0.25, 0.25))
        ks, es2 = interop.sp_indices_twist(2, Emax, (0.25, 0.25), order="shipped")
        assert ks.shape == (M, 2) and np.array_equal(es, es2) and np.allclose((ks ** 2).sum(1), es)
        assert sorted(map(tuple, ks)) == sorted(map(tuple, kt)) and not np.array_equal(ks, kt)
    z = np.load(os.path.join(GOLDEN, "van_n13_pretrained.npz"))
    ks, _ = interop.sp_indices_twist(2, 25, (0.25, 0.25), order="shipped")
    shells = list(range(66, 81))                                 # the well populated shells of the n = 13 ensemble: the n = 13 and
                                                                 # n = 29 pretraining runs pin them independently and agree
    assert np.array_equal(ks[shells], z["ktable_trained_order"][shells])
    for key in ("n13", "n29", "n49"):
        r = rep[key]
        assert r["std_recovered"] < 1.15 * r["std_logged_run"] and r["std_today"] > 10 * r["std_recovered"], r
    assert rep["n57"]["std_today"] > 10 * rep["n57"]["std_recovered"]
    with pytest.raises(ValueError):
        interop.sp_indices_twist(2, 16, (0.25, 0.25), order="shipped")
