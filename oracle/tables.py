"""Host-side table builders, restated from the reference (test infrastructure only).

sp_orbitals / twist_sort : src/orbitals.py:22-55
kpoints                  : src/potential.py:7-17
madelung                 : src/potential.py:19-34
box_length, k-table      : main.py:63-69, 79-90
"""
import numpy as np
from scipy.special import erfc


def box_length(n, dim):
    """main.py:64,67 — L for unit density parameter rs=1."""
    if dim == 3:
        return float((4.0 / 3.0 * np.pi * n) ** (1.0 / 3.0))
    if dim == 2:
        return float(np.sqrt(np.pi * n))
    raise ValueError("dim must be 2 or 3")


def sp_orbitals(dim, Emax=60):
    """src/orbitals.py:22-44 — integer plane-wave indices with |m|^2<=Emax, by energy."""
    n_max = int(np.floor(np.sqrt(Emax)))
    rng = np.arange(-n_max, n_max + 1)
    grids = np.meshgrid(*([rng] * dim))
    indices = np.array([g.flatten() for g in grids]).T
    Es = (indices ** 2).sum(axis=-1)
    keep = Es <= Emax
    indices, Es = indices[keep], Es[keep]
    order = Es.argsort()
    return indices[order], Es[order]


def twist_sort(indices, twist):
    """src/orbitals.py:46-55 — add the twist and re-sort by twisted energy."""
    shifted = indices + np.asarray(twist)
    Es = (shifted ** 2).sum(axis=-1)
    order = Es.argsort()
    return shifted[order], Es[order]


def ktable(dim, Emax, twist):
    """main.py:79-90 — the (M, dim) float64 table `sp_indices_twist`, descending energy."""
    idx, _ = sp_orbitals(dim, Emax)
    idx_t, Es_t = twist_sort(idx, twist)
    return np.ascontiguousarray(idx_t[::-1], dtype=np.float64), np.ascontiguousarray(Es_t[::-1])


def kpoints(dim, Gmax):
    """src/potential.py:7-17 — integer G with 0 < |G|^2 <= Gmax^2, meshgrid order."""
    rng = np.arange(-Gmax, Gmax + 1)
    grids = np.meshgrid(*([rng] * dim))
    G = np.array([g.flatten() for g in grids]).T
    G2 = (G ** 2).sum(axis=-1)
    return G[(G2 <= Gmax ** 2) * (G2 > 0)]


def ewald_gk(dim, kappa, G):
    """src/potential.py:25-32 / :52-59 — reciprocal weights g_G and the G=0 term g_0."""
    Gnorm = np.linalg.norm(np.asarray(G, dtype=np.float64), axis=-1)
    if dim == 3:
        g_k = np.exp(-np.pi ** 2 * Gnorm ** 2 / kappa ** 2) / (np.pi * Gnorm ** 2)
        g_0 = -np.pi / kappa ** 2
    elif dim == 2:
        g_k = erfc(np.pi * Gnorm / kappa) / Gnorm
        g_0 = -2 * np.sqrt(np.pi) / kappa
    else:
        raise ValueError("dim must be 2 or 3")
    return g_k, g_0


def madelung(dim, kappa, G):
    """src/potential.py:19-34."""
    g_k, g_0 = ewald_gk(dim, kappa, G)
    return float(g_k.sum() + g_0 - 2 * kappa / np.sqrt(np.pi))
