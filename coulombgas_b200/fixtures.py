"""Loaders of the committed fixtures under tests/golden (outputs of the reference's own sources on the shipped
checkpoints, tools/make_golden*.py): walkers, trained parameters, k-tables and reference results per configuration.
Used by the tests, by bench.py (synthetic inputs of the named shapes) and by the developer tools; not on the compute path."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES_2D = ["n13_rs10", "n29_rs1", "n29_rs10", "n49_rs0.5", "n57_rs0.25"]


def _params(z):
    p = {}
    for k in z.files:
        if k.startswith("param:"):
            _, mod, leaf = k.split(":")
            p.setdefault(mod, {})[leaf] = z[k]
    return p


def load_case(tag):
    """tests/golden/ref_<tag>.npz -> dict with `params` rebuilt in the haiku layout."""
    z = np.load(os.path.join(GOLDEN, "ref_%s.npz" % tag))
    d = {k: z[k] for k in z.files if not k.startswith("param:")}
    d["params"] = _params(z)
    for k in ("n", "dim", "Emax", "depth", "kappa", "Gmax", "mc_steps"):
        if k in d:
            d[k] = int(d[k])
    for k in ("L", "rs", "mc_stddev", "ref_madelung", "ref_mc_accept_rate"):
        if k in d:
            d[k] = float(d[k])
    return d


def load_van(tag):
    """tests/golden/van_<tag>.npz (momentum sampler fixtures)."""
    z = np.load(os.path.join(GOLDEN, "van_%s.npz" % tag))
    d = {k: z[k] for k in z.files if not k.startswith("param:")}
    d["params"] = _params(z)
    for k in ("n", "M", "Emax"):
        d[k] = int(d[k])
    return d
