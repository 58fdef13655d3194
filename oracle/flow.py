"""FermiNet-style backflow x -> z, restated from src/flow.py:5-55 in torch float64.

Test infrastructure only.  Parameters use the haiku dict layout of the shipped
checkpoints:
    "fermi_net/~/linear"   .. "fermi_net/~/linear_{depth-1}"      splayers[0..depth-1]
    "fermi_net/~/linear_{depth}" .. "fermi_net/~/linear_{2depth-2}"  tplayers[0..depth-2]
    "fermi_net/linear"                                            final
(src/flow.py:11-14 creates the single-particle layers first, then the two-particle
layers; haiku numbers them in creation order.)
"""
import numpy as np
import torch


def _name(i):
    return "fermi_net/~/linear" if i == 0 else "fermi_net/~/linear_%d" % i


def depth_of(params):
    nlin = sum(1 for k in params if k.startswith("fermi_net/~/linear"))
    return (nlin + 1) // 2


def to_torch(params):
    return {k: {kk: torch.as_tensor(np.asarray(vv), dtype=torch.float64) for kk, vv in v.items()}
            for k, v in params.items()}


def init_params(depth, spsize, tpsize, dim, seed=0, stddev=0.01):
    """Haiku-equivalent init (src/flow.py:11-14,54): N(0, stddev) weights, zero biases."""
    rng = np.random.default_rng(seed)
    p = {}
    f_sp = [4 * dim + 1 if i == 0 else 2 * spsize + tpsize for i in range(depth)]
    f_tp = [2 * dim + 1 if i == 0 else tpsize for i in range(depth - 1)]
    for i in range(depth):
        p[_name(i)] = {"w": rng.normal(0, stddev, (f_sp[i], spsize)), "b": np.zeros(spsize)}
    for i in range(depth - 1):
        p[_name(depth + i)] = {"w": rng.normal(0, stddev, (f_tp[i], tpsize)), "b": np.zeros(tpsize)}
    p["fermi_net/linear"] = {"w": rng.normal(0, stddev, (spsize, dim)), "b": np.zeros(dim)}
    return p


def _linear(p, x):
    return x @ p["w"] + p["b"]          # hk.Linear


def _tpstream0(x, L):
    """src/flow.py:20-26."""
    n = x.shape[0]
    rij = x[:, None, :] - x
    cos_rij, sin_rij = torch.cos(2 * np.pi / L * rij), torch.sin(2 * np.pi / L * rij)
    eye = torch.eye(n, dtype=x.dtype)
    dij = torch.linalg.norm(torch.sin(np.pi / L * rij) + eye[..., None], dim=-1) * (1.0 - eye)
    return torch.cat((cos_rij, sin_rij, dij[..., None]), dim=-1)


def _f(sp, tp):
    """src/flow.py:28-37."""
    n = sp.shape[0]
    return torch.cat((sp, sp.mean(dim=0, keepdim=True).repeat(n, 1), tp.mean(dim=1)), dim=-1)


def flow(params, x, L):
    """src/flow.py:39-55.  x: (n, dim) torch float64; params: torch dict."""
    depth = depth_of(params)
    sp, tp = torch.zeros_like(x), _tpstream0(x, L)
    softplus = torch.nn.functional.softplus
    for i in range(depth - 1):
        f = _f(sp, tp)
        if i == 0:
            sp = softplus(_linear(params[_name(i)], f))
            tp = softplus(_linear(params[_name(depth + i)], tp))
        else:
            sp = sp + softplus(_linear(params[_name(i)], f))
            tp = tp + softplus(_linear(params[_name(depth + i)], tp))
    f = _f(sp, tp)
    sp = sp + softplus(_linear(params[_name(depth - 1)], f))
    return x + _linear(params["fermi_net/linear"], sp)
