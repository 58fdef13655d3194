"""coulombgas_b200 — B200 (sm_100a) kernels for the forward-sampling hot path of fermiflow/CoulombGas,
behind the reference's own call signatures:

    logp = make_logp(sp_indices_twist, L)            # src/logpsi.py:7-33,174-181 (+ src/flow.py, src/slater.py)
    logp(x, params_flow, state_indices)              # -> (batch,)
    mcmc(logp.bind(params_flow, state_indices), x_init, key, mc_steps, mc_stddev)   # src/MCMC.py:7
    sample_x(key, state_indices, logp, x, params_flow, mc_steps, mc_stddev, L)      # src/VMC.py:8-25 (minus the VAN)
    potential_energy(x, kappa, G, L, rs)             # src/potential.py:70

The compute path is libcgb200.so (hand-written CUDA, C ABI in include/cgb200.h).  torch is used only for
device memory, streams and torch.distributed.  There is no CPU fallback.
"""
from ._lib import CgbError, LIB_PATH, NPARAMS  # noqa: F401
from .api import (ExplicitNoise, Logp, BoundLogp, make_logp, mcmc, sample_x, potential_energy,  # noqa: F401
                  potential_stats, pack_params, ewald_tables, kpoints, madelung, Config, microbench,
                  make_autoregressive_sampler, sample_stateindices_and_x, pack_van_params, ExplicitGumbel, VanSampler,
                  set_sweep_kernel, set_debug, device_key, debug_rng, SamplingStep)
from . import flops, dist, interop  # noqa: F401
