"""Host-side mirror of the reference's hot-path interface (same names, argument order and meaning),
dispatching to the C ABI of libcgb200.so.  Reference file:line citations are relative to the reference tree.

Arrays: torch CUDA float64 tensors are used in place (zero copy; results are CUDA tensors on the current
stream); NumPy arrays are copied host->device->host around the call (results are NumPy).  Everything is
float64 at the boundary, as in the reference (main.py:2-3 enables x64)."""
import ctypes
import dataclasses
from typing import Optional

import numpy as np
import torch

import os

from . import _lib
from ._lib import lib, check, NPARAMS

# ----------------------------------------------------------------------------------------------------
# developer switches: kernel choice / debug paths are per-call `flags` of the C ABI (include/cgb200.h).  The host layer
# keeps a process-wide default for tests and A/B timings; the environment is read ONCE, here, at import.
# ----------------------------------------------------------------------------------------------------
_SWITCH = {"kernel": os.environ.get("CGB_SWEEP_KERNEL") or None, "force_pivot": os.environ.get("CGB_FORCE_PIVOT") == "1",
           "force_f64_slater": os.environ.get("CGB_FORCE_F64_SLATER") == "1",
           "segments": {"1": True, "0": False}.get(os.environ.get("CGB_SEGMENTS"))}
if _SWITCH["kernel"] not in _lib.KERNELS:
    raise ValueError("CGB_SWEEP_KERNEL must be one of %s" % sorted(k for k in _lib.KERNELS if k))


def set_sweep_kernel(name=None):
    """Force a sweep kernel for A/B timings and tests: None/'auto', 'cta', 'cta4', 'cta8', 'warp', 'spec4', 'spec8'.
    The results do not depend on it.  Returns the previous choice."""
    if name not in _lib.KERNELS:
        raise ValueError("unknown sweep kernel %r" % (name,))
    prev, _SWITCH["kernel"] = _SWITCH["kernel"], (None if name in (None, "auto") else name)
    return prev


def set_debug(force_pivot=None, force_f64_slater=None, segments="keep"):
    """Debug paths: the partially pivoted fallback LU of the Jacobian / the float64 Slater LU for every evaluation;
    segments=True cuts every chain of the CTA-per-walker kernel into 2-step segments, False never, None = the launcher's choice."""
    if segments != "keep":
        _SWITCH["segments"] = None if segments is None else bool(segments)
    if force_pivot is not None:
        _SWITCH["force_pivot"] = bool(force_pivot)
    if force_f64_slater is not None:
        _SWITCH["force_f64_slater"] = bool(force_f64_slater)


def _flags(kernel=None):
    f = _lib.KERNELS[kernel if kernel is not None else _SWITCH["kernel"]]
    if _SWITCH["force_pivot"]:
        f |= _lib.FLAG_FORCE_PIVOT
    if _SWITCH["force_f64_slater"]:
        f |= _lib.FLAG_FORCE_F64_SLATER
    if _SWITCH["segments"] is not None:
        f |= _lib.FLAG_FORCE_SEGMENTS if _SWITCH["segments"] else _lib.FLAG_NO_SEGMENTS
    return f


# ----------------------------------------------------------------------------------------------------
# parameters / tables
# ----------------------------------------------------------------------------------------------------
_ORDER = ("fermi_net/~/linear_2", "fermi_net/~/linear", "fermi_net/~/linear_1", "fermi_net/linear")
_SHAPES = ((5, 16), (9, 16), (48, 16), (16, 2))


def pack_params(params_flow):
    """haiku dict {module: {"w","b"}} of FermiNet(depth=2, spsize=16, tpsize=16), dim 2 (src/flow.py:5-14,54;
    SURVEY §8 a13) -> flat float64 vector [Wt,bt,W0,b0,W1,b1,Wf,bf] (include/cgb200.h).  A flat vector of
    length 1074 (NumPy or torch) passes through."""
    if isinstance(params_flow, torch.Tensor):
        if params_flow.numel() != NPARAMS:
            raise ValueError("flat flow parameter vector must have %d entries" % NPARAMS)
        return params_flow.reshape(-1).to(torch.float64)
    if isinstance(params_flow, np.ndarray):
        if params_flow.size != NPARAMS:
            raise ValueError("flat flow parameter vector must have %d entries" % NPARAMS)
        return np.ascontiguousarray(params_flow.reshape(-1), dtype=np.float64)
    if set(params_flow.keys()) != set(_ORDER):
        raise ValueError("unsupported flow: expected depth=2 FermiNet modules %s, got %s"
                         % (sorted(_ORDER), sorted(params_flow.keys())))
    parts = []
    for name, shape in zip(_ORDER, _SHAPES):
        w, b = np.asarray(params_flow[name]["w"], dtype=np.float64), np.asarray(params_flow[name]["b"], dtype=np.float64)
        if w.shape != shape or b.shape != (shape[1],):
            raise ValueError("unsupported flow shape for %s: w %s b %s (need spsize=tpsize=16, dim=2)"
                             % (name, w.shape, b.shape))
        parts += [w.ravel(), b.ravel()]
    return np.concatenate(parts)


def kpoints(dim, Gmax):
    """src/potential.py:7-17 — host table builder (integer G, 0 < |G|^2 <= Gmax^2, meshgrid order)."""
    rng = np.arange(-Gmax, Gmax + 1)
    grids = np.meshgrid(*([rng] * dim))
    G = np.array([g.flatten() for g in grids]).T
    G2 = (G ** 2).sum(axis=-1)
    G = np.ascontiguousarray(G[(G2 <= Gmax ** 2) * (G2 > 0)])
    G.setflags(write=False)           # a constant table (main.py:170); lets potential_energy skip re-hashing it
    return G


def madelung(dim, kappa, G):
    """src/potential.py:19-34 (dim 2 branch) — host, once per run (main.py:171)."""
    if dim != 2:
        raise NotImplementedError("only dim=2 is supported")
    import math
    Gn = np.linalg.norm(np.asarray(G, dtype=np.float64), axis=-1)
    g_k = np.array([math.erfc(math.pi * g / kappa) / g for g in Gn])
    return float(g_k.sum() - 2 * math.sqrt(math.pi) / kappa - 2 * kappa / math.sqrt(math.pi))


def _dev():
    if not torch.cuda.is_available():
        raise _lib.CgbError("coulombgas_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _to_dev(a, dtype):
    """-> (contiguous CUDA tensor, was_numpy)"""
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            return a.to(_dev(), dtype).contiguous(), False
        if a.dtype != dtype or not a.is_contiguous():
            a = a.to(dtype).contiguous()
        return a, False
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(_dev(), non_blocking=False), True


# ----------------------------------------------------------------------------------------------------
# logp
# ----------------------------------------------------------------------------------------------------
class Logp:
    """The callable returned by make_logp: logp(x, params, state_idx) -> (batch,)  (src/logpsi.py:176-179)."""

    def __init__(self, sp_indices, L):
        kt = np.ascontiguousarray(np.asarray(sp_indices, dtype=np.float64))
        if kt.ndim != 2 or kt.shape[1] != 2:
            raise ValueError("sp_indices must be (num_states, 2): only dim=2 is supported")
        tw = (ctypes.c_double * 2)()
        check(lib.cgb_ktable_twist_host(kt.ctypes.data_as(ctypes.c_void_p), kt.shape[0], tw), "cgb_ktable_twist_host")
        self.ktable_host, self.M, self.L = kt, kt.shape[0], float(L)
        self.twist = tw
        self._ktable_dev = {}
        self._ws = {}

    def ktable_dev(self):
        d = torch.cuda.current_device()
        if d not in self._ktable_dev:
            self._ktable_dev[d] = torch.as_tensor(self.ktable_host).to(_dev())
        return self._ktable_dev[d]

    def workspace(self, B, n):
        d = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        nbytes = ctypes.c_size_t()
        check(lib.cgb_logp_workspace_bytes(B, n, ctypes.byref(nbytes)), "cgb_logp_workspace_bytes")
        if d not in self._ws or self._ws[d].numel() < nbytes.value:      # grows with the batch (per-walker scheduler state)
            new = torch.zeros(nbytes.value, dtype=torch.uint8, device=_dev())
            if d in self._ws:
                new[4096:4112].copy_(self._ws[d][4096:4112])              # the fallback counters are cumulative (device-side copy:
            self._ws[d] = new                                             # no host sync, legal under graph capture)
        return self._ws[d]

    def fallback_count(self):
        """Cumulative number of log p evaluations (this device/stream workspace) whose Jacobian left the safe range
        of the unpivoted 2x2-block LU and were redone with the partially pivoted LU."""
        d = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        if d not in self._ws:
            return 0
        return int(self._ws[d][4096:4104].view(torch.int64).item())

    def slater_fallback_count(self):
        """Cumulative number of evaluations (warp kernel) whose FP32 Slater LU exceeded its error indicator, or whose wave
        vectors exceed the power tables, and that were redone by the float64 LU."""
        d = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        if d not in self._ws:
            return 0
        return int(self._ws[d][4104:4112].view(torch.int64).item())

    def __call__(self, x, params, state_idx, return_parts=False, return_jac=False):
        xd, was_np = _to_dev(x, torch.float64)
        if xd.ndim != 3 or xd.shape[2] != 2:
            raise ValueError("x must be (batch, n, 2)")
        B, n = int(xd.shape[0]), int(xd.shape[1])
        pd, _ = _to_dev(pack_params(params), torch.float64)
        sd, _ = _to_dev(state_idx, torch.int32)
        if tuple(sd.shape) != (B, n):
            raise ValueError("state_idx must be (batch, n)")
        out = torch.empty(B, dtype=torch.float64, device=xd.device)
        parts = torch.empty((B, 2), dtype=torch.float64, device=xd.device) if return_parts else None
        jac = torch.empty((B, 2 * n, 2 * n), dtype=torch.float32, device=xd.device) if return_jac else None
        ws = self.workspace(B, n)
        check(lib.cgb_logp_f64io(_stream(), _ptr(xd), _ptr(sd), _ptr(self.ktable_dev()), self.M, self.twist, _ptr(pd),
                                 B, n, self.L, _ptr(out), _ptr(parts), _ptr(jac), _ptr(ws), ws.numel(), _flags()),
              "cgb_logp_f64io")
        res = [out] + ([parts] if return_parts else []) + ([jac] if return_jac else [])
        if was_np:
            res = [r.cpu().numpy() for r in res]
        return res[0] if len(res) == 1 else tuple(res)

    def bind(self, params, state_idx):
        """The closure `lambda x: logp(x, params_flow, state_indices)` of src/VMC.py:23, in a form the fused
        sweep kernel can consume."""
        return BoundLogp(self, params, state_idx)


class BoundLogp:
    def __init__(self, logp, params, state_idx):
        self.logp = logp
        self.params, _ = _to_dev(pack_params(params), torch.float64)
        self.state_idx, _ = _to_dev(state_idx, torch.int32)

    def __call__(self, x):
        return self.logp(x, self.params, self.state_idx)


def make_logp(sp_indices, L, depth=2, spsize=16, tpsize=16):
    """Replaces make_logp(make_logpsi(flow, sp_indices, L)) (src/logpsi.py:7,174; main.py:162-164) for
    flow = FermiNet(depth=2, spsize=16, tpsize=16, L) in 2D.  sp_indices = sp_indices_twist (main.py:90)."""
    if (depth, spsize, tpsize) != (2, 16, 16):
        raise NotImplementedError("kernels are built for depth=2, spsize=tpsize=16 (every BASELINE config)")
    return Logp(sp_indices, L)


# ----------------------------------------------------------------------------------------------------
# mcmc
# ----------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class ExplicitNoise:
    """Parity-mode `key`: the proposal noise and acceptance uniforms that jax.random.normal/uniform would have
    produced (src/MCMC.py:26,29).  noise (mc_steps, batch, n, 2) ~ N(0,1); uniforms (mc_steps, batch) in [0,1)."""
    noise: object
    uniforms: object


def _seed_of(key):
    if isinstance(key, (int, np.integer)):
        return int(key) & 0xFFFFFFFFFFFFFFFF
    k = np.asarray(key.cpu() if isinstance(key, torch.Tensor) else key).astype(np.uint64).reshape(-1)
    if k.size != 2:
        raise ValueError("key must be an int seed, a (2,) uint32 PRNG key, or ExplicitNoise")
    return int((k[0] << np.uint64(32)) | k[1])


def _is_device_key(key):
    """A CUDA tensor holding the reference's per-device PRNG key, (2,) uint32 (src/VMC.py:20): the kernels read it
    themselves (no device->host sync, graph-capturable)."""
    return isinstance(key, torch.Tensor) and key.is_cuda


def _device_key(key):
    if key.numel() != 2 or key.dtype not in (torch.uint32, torch.int32) or not key.is_contiguous():
        raise ValueError("a device key must be a contiguous (2,) uint32 (or int32 bit pattern) CUDA tensor")
    return key


def device_key(seed, device=None):
    """(2,) uint32 CUDA tensor = [seed >> 32, seed & 0xffffffff]: the device-resident form of an integer key."""
    seed = _seed_of(seed)
    k = np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)
    return torch.as_tensor(k.view(np.int32)).to(device or _dev()).view(torch.uint32)


def mcmc(logp_fn, x_init, key, mc_steps, mc_stddev=0.02, wrap_L=None, return_trace=False, stats=None, advance_key=False,
         _want_rate=True):
    """Metropolis sweep with the reference's signature (src/MCMC.py:7).

    logp_fn : a BoundLogp (`logp.bind(params_flow, state_indices)`); arbitrary Python callables are not supported.
    key     : int / host (2,) uint32 key -> in-kernel Philox stream keyed by it (throughput mode); a CUDA (2,) uint32
              tensor -> the same, with the key read on the device (the traced per-device key of src/VMC.py:20; no host
              sync; with advance_key=True it is replaced by key + 0x9E3779B97F4A7C15 before use, in place);
              ExplicitNoise -> parity mode.
    Returns (x, accept_rate) — accept_rate is this process's mean (MCMC.py:38); the cross-device pmean of
    MCMC.py:39 is done by sample_x.  x is updated IN PLACE when x_init is a contiguous CUDA float64 tensor (donated
    buffer, src/VMC.py:11); any other input is copied and the advanced copy is returned.  wrap_L applies src/VMC.py:24
    inside the kernel.  return_trace adds a dict with per-step accept flags, proposal log p, final log p and per-walker
    accept counts.  stats: optional CUDA float64 vector (>= 8) that receives [accepted moves, proposals] (include/cgb200.h)."""
    if not isinstance(logp_fn, BoundLogp):
        raise TypeError("mcmc needs logp.bind(params_flow, state_indices); generic callables have no kernel")
    lg = logp_fn.logp
    # every argument is validated BEFORE the launch: the kernel advances x in place
    if wrap_L is not None and abs(float(wrap_L) - lg.L) > 1e-12 * lg.L:
        raise ValueError("wrap_L must equal the box length the logp was built with")
    shape = tuple(x_init.shape)
    if len(shape) != 3 or shape[2] != 2:
        raise ValueError("x_init must be (batch, n, 2)")
    B, n = int(shape[0]), int(shape[1])
    S = int(mc_steps)
    if S < 1:
        raise ValueError("mc_steps must be >= 1")
    if tuple(logp_fn.state_idx.shape) != (B, n):
        raise ValueError("state_indices bound to logp_fn must be (batch, n)")
    xd, was_np = _to_dev(x_init, torch.float64)
    noise = unif = key_dev = None
    seed = 0
    if isinstance(key, ExplicitNoise):
        noise, _ = _to_dev(key.noise, torch.float64)
        unif, _ = _to_dev(key.uniforms, torch.float64)
        if tuple(noise.shape) != (S, B, n, 2) or tuple(unif.shape) != (S, B):
            raise ValueError("ExplicitNoise shapes must be (mc_steps,batch,n,2) and (mc_steps,batch)")
    elif _is_device_key(key):
        key_dev = _device_key(key)
    else:
        seed = _seed_of(key)
    if advance_key and key_dev is None:
        raise ValueError("advance_key needs a device key")
    if stats is not None and (not stats.is_cuda or stats.dtype != torch.float64 or stats.numel() < _lib.STATS_LEN):
        raise ValueError("stats must be a CUDA float64 tensor with >= %d entries" % _lib.STATS_LEN)
    dev = xd.device
    out_logp = torch.empty(B, dtype=torch.float64, device=dev)
    nacc = torch.empty(B, dtype=torch.int32, device=dev)
    acc_tr = torch.empty((S, B), dtype=torch.uint8, device=dev) if return_trace else None
    lp_tr = torch.empty((S + 1, B), dtype=torch.float64, device=dev) if return_trace else None
    ws = lg.workspace(B, n)
    flags = _flags() | (_lib.FLAG_ADVANCE_KEY if advance_key else 0)
    check(lib.cgb_mcmc_sweep_f64io(_stream(), _ptr(xd), _ptr(logp_fn.state_idx), _ptr(lg.ktable_dev()), lg.M, lg.twist,
                                   _ptr(logp_fn.params), _ptr(noise), _ptr(unif), seed, _ptr(key_dev), S, float(mc_stddev),
                                   lg.L, B, n, 0 if wrap_L is None else 1, _ptr(out_logp), _ptr(nacc), _ptr(acc_tr),
                                   _ptr(lp_tr), _ptr(stats), _ptr(ws), ws.numel(), flags), "cgb_mcmc_sweep_f64io")
    if not _want_rate:                       # fused sampling step: the rate comes from the statistics vector
        return xd, None
    accept_rate = nacc.sum(dtype=torch.float64) / (S * B)
    x_out = xd.cpu().numpy() if was_np else xd
    if was_np:
        accept_rate = float(accept_rate)
    if return_trace:
        tr = dict(accept=acc_tr, logp_prop=lp_tr, logp=out_logp, naccept=nacc)
        if was_np:
            tr = {k: v.cpu().numpy() for k, v in tr.items()}
        return x_out, accept_rate, tr
    return x_out, accept_rate


def _all_reduce_sum(vec):
    from .dist import psum
    return psum(vec)


def _world():
    from .dist import world
    return world()


def sample_x(key, state_indices, logp, x, params_flow, mc_steps, mc_stddev, L):
    """One sampling step of src/VMC.py:8-25 for THIS process's shard of walkers (one process per GPU replaces
    jax.pmap): the sweep (VMC.py:23), the wrap into [0,L) (VMC.py:24) and the cross-device mean of the accept
    rate (MCMC.py:39) as one all-reduce over the default process group (NCCL on GPUs).  The VAN draw of
    VMC.py:22 stays with the caller, which passes `state_indices`.  Returns (key', x, accept_rate) where key'
    is the advanced integer key; a device key (CUDA (2,) uint32, one per rank) is advanced in place by the sweep's prologue
    kernel and returned as is."""
    dev_key = _is_device_key(key)
    seed = _seed_of(key) if not (isinstance(key, ExplicitNoise) or dev_key) else None
    r, w = _world()
    # the Philox counters are (local walker, step): ranks must not share a key.  The reference carries one key per device
    # (keys of shape (ndev, 2), main.py:236); an integer key is therefore folded with the rank here, a device key is
    # per-rank by construction (the caller holds one per GPU).
    run_key = key if seed is None else _fold_rank(seed, r)
    x, acc = mcmc(logp.bind(params_flow, state_indices), x, run_key, mc_steps, mc_stddev, wrap_L=L, advance_key=dev_key)
    acc_t = acc if isinstance(acc, torch.Tensor) else torch.tensor(acc, dtype=torch.float64, device=_dev())
    rate = acc_t                               # equal shards (main.py:231-238): mean of per-device means == global mean
    if w > 1:
        rate = _all_reduce_sum(acc_t.reshape(1).clone())[0] / w
    new_key = key if seed is None else (seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
    return new_key, x, (rate if isinstance(acc, torch.Tensor) else float(rate))


def _fold_rank(seed, rank):
    """jax.random.fold_in analogue for integer keys: rank 0 keeps the key (single-process runs are unchanged)."""
    if rank == 0:
        return seed
    z = (seed + 0x9E3779B97F4A7C15 * (rank + 1)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


# ----------------------------------------------------------------------------------------------------
# autoregressive momentum sampler (SURVEY §8 f/N2)
# ----------------------------------------------------------------------------------------------------
def pack_van_params(params_van, nlayers=2, prefix="transformer"):
    """haiku dict of src/autoregressive.py:Transformer (modules `transformer`, `transformer/embedding_mlp`,
    `transformer/layer{i}_attn/{query,key,value,linear}`, `transformer/layer{i}_mlp/{linear,linear_1}`,
    `transformer/output_mlp`) -> the flat float64 vector of include/cgb200.h (cgb_van_*).  Flat vectors pass through."""
    if isinstance(params_van, torch.Tensor):
        return params_van.reshape(-1).to(torch.float64)
    if isinstance(params_van, np.ndarray):
        return np.ascontiguousarray(params_van.reshape(-1), dtype=np.float64)
    parts = [np.asarray(params_van[prefix]["x1hat"], dtype=np.float64).ravel()]

    def wb(name):
        parts.append(np.asarray(params_van[name]["w"], dtype=np.float64).ravel())
        parts.append(np.asarray(params_van[name]["b"], dtype=np.float64).ravel())
    wb(prefix + "/embedding_mlp")
    for i in range(nlayers):
        for leaf in ("query", "key", "value", "linear"):
            wb("%s/layer%d_attn/%s" % (prefix, i, leaf))
        wb("%s/layer%d_mlp/linear" % (prefix, i))
        wb("%s/layer%d_mlp/linear_1" % (prefix, i))
    wb(prefix + "/output_mlp")
    return np.concatenate(parts)


@dataclasses.dataclass
class ExplicitGumbel:
    """Parity-mode `key` of the sampler: the Gumbel noise jax.random.categorical would add to the masked logits of
    every draw (src/sampler.py:36-37), shape (batch, n, num_states)."""
    noise: object


class VanSampler:
    """`sampler` / `log_prob` of make_autoregressive_sampler (src/sampler.py:3-49) on the GPU."""

    def __init__(self, sp_indices, n, num_states):
        kt = np.ascontiguousarray(np.asarray(sp_indices, dtype=np.float64))
        if kt.shape != (num_states, 2):
            raise ValueError("sp_indices must be (num_states, 2): only dim=2 is supported")
        self.ktable_host, self.n, self.M = kt, int(n), int(num_states)
        self.nparams = int(lib.cgb_van_nparams(self.M))
        self._kt = {}

    def _ktable(self):
        d = torch.cuda.current_device()
        if d not in self._kt:
            self._kt[d] = torch.as_tensor(self.ktable_host).to(_dev())
        return self._kt[d]

    def _params(self, params):
        pd, _ = _to_dev(pack_van_params(params), torch.float64)
        if pd.numel() != self.nparams:
            raise ValueError("unsupported VAN shape: expected %d parameters (nlayers=2, modelsize=16, nheads=4, nhidden=32, "
                             "num_states=%d), got %d" % (self.nparams, self.M, pd.numel()))
        return pd

    def sampler(self, params, key, batch, return_log_prob=False, return_logits=False):
        """src/sampler.py:30-38: (batch, n) int32 strictly increasing state indices.  key: int / (2,) uint32 -> in-kernel
        Philox; ExplicitGumbel -> parity mode.  Results are CUDA tensors."""
        pd = self._params(params)
        B = int(batch)
        g, seed, key_dev = None, 0, None
        if isinstance(key, ExplicitGumbel):
            g, _ = _to_dev(key.noise, torch.float64)
            if tuple(g.shape) != (B, self.n, self.M):
                raise ValueError("ExplicitGumbel noise must be (batch, n, num_states)")
        elif _is_device_key(key):
            key_dev = _device_key(key)
        else:
            seed = _fold_rank(_seed_of(key), _world()[0])       # one stream per rank (src/VMC.py:20: per-device keys)
        dev = pd.device
        out = torch.empty((B, self.n), dtype=torch.int32, device=dev)
        lp = torch.empty(B, dtype=torch.float64, device=dev) if return_log_prob else None
        lg = torch.empty((B, self.n, self.M), dtype=torch.float64, device=dev) if return_logits else None
        check(lib.cgb_van_sample_f64io(_stream(), _ptr(pd), self.nparams, _ptr(self._ktable()), self.M, self.n, B, _ptr(g),
                                       seed, _ptr(key_dev), _ptr(out), _ptr(lp), _ptr(lg)), "cgb_van_sample_f64io")
        res = [out] + ([lp] if return_log_prob else []) + ([lg] if return_logits else [])
        return res[0] if len(res) == 1 else tuple(res)

    def log_prob(self, params, state_indices, return_logits=False):
        """vmap(log_prob_novmap, (None, 0)) of main.py:108 / src/sampler.py:40-44: (batch, n) -> (batch,)."""
        pd = self._params(params)
        sd, was_np = _to_dev(state_indices, torch.int32)
        if sd.ndim != 2 or sd.shape[1] != self.n:
            raise ValueError("state_indices must be (batch, n)")
        B = int(sd.shape[0])
        lp = torch.empty(B, dtype=torch.float64, device=pd.device)
        lg = torch.empty((B, self.n, self.M), dtype=torch.float64, device=pd.device) if return_logits else None
        check(lib.cgb_van_logprob_f64io(_stream(), _ptr(pd), self.nparams, _ptr(self._ktable()), self.M, self.n, B, _ptr(sd),
                                        _ptr(lp), _ptr(lg)), "cgb_van_logprob_f64io")
        if was_np:
            lp, lg = lp.cpu().numpy(), (lg.cpu().numpy() if lg is not None else None)
        return (lp, lg) if return_logits else lp


def make_autoregressive_sampler(sp_indices, n, num_states, nlayers=2, modelsize=16, nheads=4, nhidden=32):
    """src/sampler.py:3 with `network` = the Transformer of src/autoregressive.py at the reference's default shape
    (main.py:24-27, 96-99).  Returns (sampler(params, key, batch), log_prob(params, state_indices))."""
    if (nlayers, modelsize, nheads, nhidden) != (2, 16, 4, 32):
        raise NotImplementedError("the VAN kernel is built for nlayers=2, modelsize=16, nheads=4, nhidden=32")
    v = VanSampler(sp_indices, n, num_states)
    return v.sampler, v.log_prob


def sample_stateindices_and_x(key, sampler, params_van, logp, x, params_flow, mc_steps, mc_stddev, L):
    """src/VMC.py:8-25 for THIS process's shard (one process per GPU replaces jax.pmap): key split (VMC.py:20), VAN draw of
    the occupied orbitals (VMC.py:22), Metropolis sweep (VMC.py:23), wrap (VMC.py:24), cross-device mean of the accept rate
    (MCMC.py:39).  `key` is an integer; returns (key', state_indices, x, accept_rate) like the reference."""
    if _is_device_key(key):
        # device-resident key: the sampler and the sweep draw from different Philox stream ids of the same key; the sweep's
        # prologue advances the key in place for the next step (the analogue of VMC.py:20's split)
        state_indices = sampler(params_van, key, int(x.shape[0]))
        _, x, rate = sample_x(key, state_indices, logp, x, params_flow, mc_steps, mc_stddev, L)
        return key, state_indices, x, rate
    seed = _seed_of(key)
    mix = lambda s, c: (s * 6364136223846793005 + c) & 0xFFFFFFFFFFFFFFFF
    key_next, key_state, key_mcmc = mix(seed, 1442695040888963407), mix(seed, 0x9E3779B97F4A7C15), mix(seed, 0xD1B54A32D192ED03)
    state_indices = sampler(params_van, key_state, int(x.shape[0]))
    _, x, rate = sample_x(key_mcmc, state_indices, logp, x, params_flow, mc_steps, mc_stddev, L)
    return key_next, state_indices, x, rate


# ----------------------------------------------------------------------------------------------------
# Ewald
# ----------------------------------------------------------------------------------------------------
_EWALD_CACHE = {}
_EWALD_LAST = [None, None, None, None]       # (G object, kappa, device, tables): identity fast path for the hot loop


def ewald_tables(G, kappa):
    """Fold (G, kappa) into the merged reciprocal item table (cgb_ewald_prepare_host); cached per device.
    The training loop passes the same constant `G` array every step (main.py:170): a repeated call with the
    identical array object skips the content hash."""
    dev = torch.cuda.current_device()
    if G is _EWALD_LAST[0] and kappa == _EWALD_LAST[1] and dev == _EWALD_LAST[2]:
        return _EWALD_LAST[3]
    Gh = np.ascontiguousarray(np.asarray(G.cpu() if isinstance(G, torch.Tensor) else G, dtype=np.int64))
    if Gh.ndim != 2 or Gh.shape[1] != 2:
        raise NotImplementedError("only dim=2 is supported (G must be (nG, 2))")
    key = (Gh.tobytes(), float(kappa), dev)
    if key not in _EWALD_CACHE:
        items = np.empty((Gh.shape[0], 4), dtype=np.float32)
        n_items, gmax = ctypes.c_int(), ctypes.c_int()
        check(lib.cgb_ewald_prepare_host(Gh.ctypes.data_as(ctypes.c_void_p), Gh.shape[0], float(kappa),
                                         items.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n_items),
                                         ctypes.byref(gmax)), "cgb_ewald_prepare_host")
        dev_items = torch.as_tensor(items[:n_items.value].copy()).to(_dev())
        _EWALD_CACHE[key] = (dev_items, n_items.value, gmax.value)
    if isinstance(G, np.ndarray) and not G.flags.writeable:
        _EWALD_LAST[:] = [G, kappa, dev, _EWALD_CACHE[key]]        # only immutable arrays qualify for the identity path
    return _EWALD_CACHE[key]


def potential_energy(x, kappa, G, L, rs, stats=None):
    """src/potential.py:69-77 — Ewald energy (coordinate-dependent part) of a batch x (batch, n, 2) -> (batch,).
    The caller adds Vconst = n*rs/L*Madelung (src/VMC.py:40, main.py:171).  stats: optional CUDA float64 vector (>= 8)
    that receives [.., .., sum V, sum V^2, batch] (include/cgb200.h CGB_STATS_*) from the kernel's own epilogue."""
    xd, was_np = _to_dev(x, torch.float64)
    if xd.ndim != 3 or xd.shape[2] != 2:
        raise NotImplementedError("only dim=2 is supported (x must be (batch, n, 2))")
    if stats is not None and (not stats.is_cuda or stats.dtype != torch.float64 or stats.numel() < _lib.STATS_LEN):
        raise ValueError("stats must be a CUDA float64 tensor with >= %d entries" % _lib.STATS_LEN)
    items, n_items, gmax = ewald_tables(G, kappa)
    B, n = int(xd.shape[0]), int(xd.shape[1])
    out = torch.empty(B, dtype=torch.float64, device=xd.device)
    check(lib.cgb_ewald_f64io(_stream(), _ptr(xd), B, n, float(L), float(rs), float(kappa), _ptr(items), n_items, gmax,
                              _ptr(out), _ptr(stats)), "cgb_ewald_f64io")
    return out.cpu().numpy() if was_np else out


class SamplingStep:
    """One sampling step of the reference's training loop for THIS rank's shard (main.py:334-338 -> src/VMC.py:23-24,40,46-53
    -> src/MCMC.py): Metropolis sweep + wrap, Ewald energy of the final walkers, and ONE all-reduce of the 5-entry statistics
    vector [accepted, proposed, sum V, sum V^2, walkers] that the kernels' epilogues fill — 3 kernel launches
    (prologue, sweep, Ewald) and one collective, all on the current stream, no host synchronisation: the whole step can be
    captured in a CUDA graph (`capture()`), the per-device PRNG key lives on the device and is advanced by the prologue.

    step() returns the CUDA tensor `stats` (float64, 8 entries; after the all-reduce the sums are over all ranks):
    accept_rate = stats[0]/stats[1], V_mean = stats[2]/stats[4] + Vconst, ... (see `summary`)."""

    def __init__(self, logp, state_indices, params_flow, x, key, mc_steps, mc_stddev, kappa, G, rs, Vconst=0.0):
        self.bound = logp.bind(params_flow, state_indices)
        self.logp, self.x = logp, x
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()):
            raise ValueError("SamplingStep needs a contiguous CUDA float64 walker tensor (updated in place)")
        self.key = key if _is_device_key(key) else device_key(_fold_rank(_seed_of(key), _world()[0]), x.device)
        self.mc_steps, self.mc_stddev = int(mc_steps), float(mc_stddev)
        self.kappa, self.G, self.rs, self.Vconst = kappa, G, float(rs), float(Vconst)
        self.stats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=x.device)
        self.V = None
        self.graph = None

    def _enqueue(self):
        mcmc(self.bound, self.x, self.key, self.mc_steps, self.mc_stddev, wrap_L=self.logp.L, stats=self.stats,
             advance_key=True, _want_rate=False)
        self.V = potential_energy(self.x, self.kappa, self.G, self.logp.L, self.rs, stats=self.stats)
        _, w = _world()
        if w > 1:
            _all_reduce_sum(self.stats[:5])          # the ONE collective of the step (replaces the pmeans of MCMC.py:39, VMC.py:46-53)

    def capture(self, warmup=2):
        """Capture the step into a CUDA graph (collective included).  Call once; `step()` then replays it."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._enqueue()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue()
        return self

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        return self.stats

    def summary(self):
        """(accept_rate, V_mean, V2_mean) incl. Vconst, as Python floats (synchronises)."""
        s = self.stats.cpu().numpy()
        nw = s[4]
        vm = s[2] / nw + self.Vconst
        v2 = (s[3] + 2 * self.Vconst * s[2]) / nw + self.Vconst ** 2
        return float(s[0] / s[1]), float(vm), float(v2)


def potential_stats(V):
    """The V_mean / V2_mean entries of the pmean at src/VMC.py:46-53, as one all-reduce of this rank's [mean V, mean V^2]."""
    Vd, was_np = _to_dev(V, torch.float64)
    _, w = _world()
    vec = torch.stack([Vd.mean(), (Vd * Vd).mean()])       # equal shards: pmean of the per-device means (VMC.py:46)
    if w > 1:
        vec = _all_reduce_sum(vec) / w
    mean, mean2 = vec[0], vec[1]
    return (float(mean), float(mean2)) if was_np else (mean, mean2)


# ----------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Config:
    """Names and defaults of the reference's flags that shape the hot path (main.py:17-43)."""
    n: int = 13
    dim: int = 2
    rs: float = 10.0
    Theta: float = 0.15
    Emax: int = 25
    twist: tuple = (0.25, 0.25)
    depth: int = 2
    spsize: int = 16
    tpsize: int = 16
    Gmax: int = 15
    kappa: int = 10
    mc_therm: int = 10
    mc_steps: int = 50
    mc_stddev: float = 0.1
    batch: int = 2048

    @property
    def L(self):
        return float(np.sqrt(np.pi * self.n))      # main.py:67


def debug_rng(key, B, S, n, kat=None):
    """The throughput mode's random streams as the sweep kernels draw them (tests): returns (normal (S,B,n,2) f32,
    uniform (S,B) f64[, philox blocks (len(kat),4) u32 for kat = (k,6) uint32 rows of counter[4]+key[2]])."""
    key_dev, seed = (_device_key(key), 0) if _is_device_key(key) else (None, _seed_of(key))
    nrm = torch.empty((S, B, n, 2), dtype=torch.float32, device=_dev())
    uni = torch.empty((S, B), dtype=torch.float64, device=_dev())
    kin = kout = None
    nk = 0
    if kat is not None:
        k = np.ascontiguousarray(np.asarray(kat, dtype=np.uint32).reshape(-1, 6))
        nk = k.shape[0]
        kin = torch.as_tensor(k.view(np.int32)).to(_dev())
        kout = torch.empty((nk, 4), dtype=torch.int32, device=_dev())
    check(lib.cgb_debug_rng(_stream(), seed, _ptr(key_dev), B, S, n, _ptr(nrm), _ptr(uni), _ptr(kin), _ptr(kout), nk),
          "cgb_debug_rng")
    if kat is not None:
        return nrm, uni, kout.cpu().numpy().view(np.uint32)
    return nrm, uni


def microbench(kind, blocks, iters):
    """FFMA (0) / MUFU.EX2 (1) / DFMA (2) saturating micro-kernel; returns seconds (CUDA events)."""
    out = torch.zeros(1, dtype=torch.float32, device=_dev())
    check(lib.cgb_microbench(_stream(), kind, blocks, 64, _ptr(out)), "cgb_microbench")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.cgb_microbench(_stream(), kind, blocks, iters, _ptr(out)), "cgb_microbench")
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3
