"""Frozen algorithmic-work model used by bench.py and the reports.
dim 2, depth 2, spsize = tpsize = 16; FMA = 2 FLOP.  This counts the *useful* analytic work, not the
reference's jacfwd / pairwise-cos formulation."""


def flop_logp(n):
    return 8 * n ** 3 + 895 * n ** 2 + 1760 * n


def flop_sweep(n, mc_steps):
    """one walker, one sweep = mc_steps + 1 evaluations of log p (src/MCMC.py:36-37)."""
    return (mc_steps + 1) * flop_logp(n)


def flop_ewald(n, nG=708):
    return 12 * n * (nG // 2) + 40 * n * (n - 1) // 2


def mufu_logp(n):
    """SFU ops per evaluation in this implementation: pass 1 ex2+lg2, pass 2 ex2+rcp per pair-channel,
    single-particle softplus/sigmoid (3 per channel, 2 layers), sincos for positions and the Slater matrix."""
    return 4 * 16 * n * n + 2 * 3 * 16 * n + 2 * (n * n + 2 * n) + 2 * n * n


def hbm_bytes_sweep(n, mc_steps, parity_mode):
    b = n * 2 * 8 * 2 + n * 4 + 8 + 4
    if parity_mode:
        b += mc_steps * (n * 2 * 8 + 8)
    return b
