"""Ewald Coulomb energy, restated as written from src/potential.py:36-77.

Test infrastructure only.  NumPy float64 + SciPy erfc; pairwise-cosine
reciprocal sum exactly as the reference does it (O(|G| n^2)).
"""
import numpy as np
from scipy.special import erfc
from .tables import ewald_gk


def psi(x, kappa, G):
    """src/potential.py:36-65.  x: (n,dim) in units of L."""
    n, dim = x.shape
    i, j = np.triu_indices(n, k=1)
    rij = (x[:, None, :] - x)[i, j]
    rij = rij - np.rint(rij)
    dij = np.linalg.norm(rij, axis=-1)
    V_shortrange = (erfc(kappa * dij) / dij).sum()
    g_k, g_0 = ewald_gk(dim, kappa, G)
    V_longrange = (g_k * np.cos(2 * np.pi * np.asarray(G).dot(rij.T)).sum(axis=-1)).sum() \
        + g_0 * rij.shape[0]
    return V_shortrange + V_longrange


def potential_energy(x, kappa, G, L, rs):
    """src/potential.py:69-77 (vmapped over the batch)."""
    x = np.asarray(x, dtype=np.float64)
    return np.array([2 * rs / L * psi(xb / L, kappa, G) for xb in x])
