"""log psi / log p, restated as written (test infrastructure only).

logslaterdet : src/slater.py:4-19
logpsi       : src/logpsi.py:7-33   (flow -> Slater logdet + 0.5*slogdet(jacfwd(flow)))
logp         : src/logpsi.py:174-181 (2*Re logpsi, vmapped over the batch)
"""
import numpy as np
import torch
from . import flow as _flow


def logslaterdet(indices, z, L):
    """src/slater.py:14-19.  indices (n,dim) = sp_indices[state_idx]; returns complex."""
    k = 2 * np.pi / L * indices
    k_dot_x = (k * z[:, None, :]).sum(dim=-1)
    dim = z.shape[1]
    D = torch.exp(1j * k_dot_x) / L ** (dim / 2)
    phase, logabsdet = torch.linalg.slogdet(D)
    return logabsdet + torch.log(phase)


def flow_jacobian(params, x, L):
    """src/logpsi.py:24-28 — forward-mode Jacobian of the flattened flow, (n*dim, n*dim)."""
    n, dim = x.shape
    f = lambda xf: _flow.flow(params, xf.reshape(n, dim), L).reshape(-1)
    return torch.func.jacfwd(f)(x.reshape(-1))


def logpsi(x, params, indices, L, parts=False):
    """src/logpsi.py:22-31.  Returns (Re, Im) of log psi; parts=True also returns pieces."""
    z = _flow.flow(params, x, L)
    log_phi = logslaterdet(indices, z, L)
    jac = flow_jacobian(params, x, L)
    _, logjacdet = torch.linalg.slogdet(jac)
    out = torch.stack([log_phi.real + 0.5 * logjacdet, log_phi.imag])
    if parts:
        return out, dict(z=z, jac=jac, logabsdet_slater=log_phi.real, logjacdet=logjacdet)
    return out


def logp_batch(x, params, state_idx, ktable, L, chunk=32):
    """src/logpsi.py:176-179 on numpy inputs (jax.vmap -> torch.func.vmap, chunked).
    x (B,n,dim), state_idx (B,n) -> (B,) float64."""
    p = _flow.to_torch(params)
    xt = torch.as_tensor(np.asarray(x), dtype=torch.float64)
    kt = torch.as_tensor(np.asarray(ktable), dtype=torch.float64)
    ind = kt[torch.as_tensor(np.asarray(state_idx), dtype=torch.long)]
    one = lambda xb, ib: 2.0 * logpsi(xb, p, ib, L)[0]
    out = [torch.func.vmap(one)(xt[s:s + chunk], ind[s:s + chunk]) for s in range(0, xt.shape[0], chunk)]
    return torch.cat(out).numpy() if out else np.empty(0)


def logpsi_batch(x, params, state_idx, ktable, L):
    """(B,2) array of (Re, Im) log psi plus a dict of per-walker pieces (z, jac, logdets)."""
    p = _flow.to_torch(params)
    xt = torch.as_tensor(np.asarray(x), dtype=torch.float64)
    kt = torch.as_tensor(np.asarray(ktable), dtype=torch.float64)
    si = np.asarray(state_idx)
    B = xt.shape[0]
    out = np.empty((B, 2))
    zs, jacs, ld_s, ld_j = [], [], [], []
    for b in range(B):
        o, parts = logpsi(xt[b], p, kt[si[b]], L, parts=True)
        out[b] = o.numpy()
        zs.append(parts["z"].numpy()); jacs.append(parts["jac"].numpy())
        ld_s.append(float(parts["logabsdet_slater"])); ld_j.append(float(parts["logjacdet"]))
    return out, dict(z=np.stack(zs), jac=np.stack(jacs),
                     logabsdet_slater=np.array(ld_s), logjacdet=np.array(ld_j))
