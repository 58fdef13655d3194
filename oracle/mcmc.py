"""Metropolis sweep, restated from src/MCMC.py:6-40 and src/VMC.py:23-24.

Test infrastructure only.  The reference draws `normal` and `uniform` from JAX's
threefry stream inside the loop (MCMC.py:24-29); that PRNG boundary is parity-
unpinned here, so the oracle (like the kernels' parity mode) takes the proposal
noise (mc_steps,B,n,dim) and uniforms (mc_steps,B) as explicit inputs.
"""
import numpy as np


def mcmc(logp_fn, x_init, noise, uniforms, mc_steps, mc_stddev, trace=False):
    """MCMC.py:22-40.  Returns (x, accept_rate[, trace dict])."""
    x = np.array(x_init, dtype=np.float64, copy=True)
    logp = logp_fn(x)                                        # :36
    num_accepts = 0.0
    tr = dict(logp_init=logp.copy(), accept=[], logp_prop=[], logp=[])
    for i in range(mc_steps):                                # :37
        x_proposal = x + mc_stddev * noise[i]                # :26
        logp_proposal = logp_fn(x_proposal)                  # :27
        with np.errstate(over="ignore", invalid="ignore"):
            ratio = np.exp(logp_proposal - logp)             # :28
        accept = uniforms[i] < ratio                         # :29 (NaN -> reject)
        x = np.where(accept[:, None, None], x_proposal, x)   # :31
        logp = np.where(accept, logp_proposal, logp)         # :32
        num_accepts += accept.sum()                          # :33
        if trace:
            tr["accept"].append(accept.copy()); tr["logp_prop"].append(logp_proposal.copy())
            tr["logp"].append(logp.copy())
    accept_rate = num_accepts / (mc_steps * x.shape[0])      # :39 (pmean over 1 device)
    if trace:
        tr = {k: np.array(v) for k, v in tr.items()}
        tr["logp_final"] = logp
        return x, accept_rate, tr
    return x, accept_rate


def wrap(x, L):
    """src/VMC.py:24."""
    return x - L * np.floor(x / L)
