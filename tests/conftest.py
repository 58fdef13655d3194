import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_case(tag):
    """Load tests/golden/ref_<tag>.npz -> dict with `params` rebuilt in the haiku layout."""
    z = np.load(os.path.join(GOLDEN, "ref_%s.npz" % tag))
    d = {k: z[k] for k in z.files if not k.startswith("param:")}
    params = {}
    for k in z.files:
        if k.startswith("param:"):
            _, mod, leaf = k.split(":")
            params.setdefault(mod, {})[leaf] = z[k]
    d["params"] = params
    for k in ("n", "dim", "Emax", "depth", "kappa", "Gmax", "mc_steps"):
        if k in d:
            d[k] = int(d[k])
    for k in ("L", "rs", "mc_stddev", "ref_madelung", "ref_mc_accept_rate"):
        if k in d:
            d[k] = float(d[k])
    return d


CASES_2D = ["n13_rs10", "n29_rs1", "n29_rs10", "n49_rs0.5", "n57_rs0.25"]


@pytest.fixture(scope="session")
def cases():
    cache = {}

    def get(tag):
        if tag not in cache:
            cache[tag] = load_case(tag)
        return cache[tag]
    return get
