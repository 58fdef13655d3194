"""Interop with the reference's on-disk formats and host-side setup tables (SURVEY §8 rows a9/a10 and f/N1):

* single-particle orbital table  sp_orbitals / twist_sort / sp_indices_twist   (src/orbitals.py:22-55, main.py:79-90)
* run directory name and checkpoint file names                                  (main.py:194-206, src/checkpoint.py:4-8)
* checkpoint reader / writer that needs neither jax, haiku nor optax            (src/checkpoint.py:10-17, main.py:376-384)
* walker / key sharding of a checkpoint over ranks                              (main.py:217-223, 231-238)
* `data.txt` line writer and reader                                             (main.py:355-374)

Pure NumPy + pickle: no device code here; everything the kernels consume from a reference run passes through
this module, so a run written by the reference's main.py can be replayed through libcgb200.so (tools/replay_checkpoint.py).
"""
import io
import os
import pickle

import numpy as np

# ----------------------------------------------------------------------------------------------------
# host-side setup tables
# ----------------------------------------------------------------------------------------------------


def box_length(n, dim=2):
    """main.py:63-67: box length at unit rs (the kernels work in these units; rs only scales the energy)."""
    if dim == 2:
        return float(np.sqrt(np.pi * n))
    if dim == 3:
        return float((4.0 / 3.0 * np.pi * n) ** (1.0 / 3.0))
    raise ValueError("dim must be 2 or 3")


def _lattice(dim, nmax):
    """All integer vectors of [-nmax, nmax]^dim in numpy.meshgrid('xy').flatten order — the enumeration order is part
    of the contract: ties in energy are broken by numpy's default argsort acting on THIS order (orbitals.py:36-42)."""
    axis = np.arange(-nmax, nmax + 1)
    mesh = np.meshgrid(*([axis] * dim))
    return np.stack([m.reshape(-1) for m in mesh], axis=1)


def sp_orbitals(dim, Emax=60, kind=None):
    """src/orbitals.py:22-44 -> (indices (M, dim) int, Es (M,) int), ascending |m|^2 <= Emax.

    CAVEAT: the reference orders DEGENERATE orbitals by numpy's default argsort, which is unstable and has
    changed between NumPy versions / CPU dispatch paths.  `kind=None` reproduces "whatever this NumPy does", like the
    reference; pass kind="stable" for a platform-independent table.  Parameters trained by the reference (the VAN in
    particular) are tied to the order of the NumPy they were trained with — keep the k-table of a run next to its
    checkpoints instead of regenerating it (tests/golden/van_n13_pretrained.npz carries the recovered order of the
    shipped Emax=25 runs)."""
    lat = _lattice(dim, int(np.floor(np.sqrt(Emax))))
    e = np.sum(lat * lat, axis=1)
    sel = e <= Emax
    lat, e = lat[sel], e[sel]
    perm = e.argsort(kind=kind)  # numpy default kind, as the reference: tie order must be reproduced, not improved
    return lat[perm], e[perm]


def twist_sort(indices, twist, kind="stable"):
    """src/orbitals.py:46-55 -> (indices + twist re-sorted by twisted energy, twisted energies).

    The reference calls this with `twist = jnp.array(args.twist)` (main.py:72,88): `indices + twist` is then a jax array and
    `Es_twist.argsort()` is jnp.argsort, which is STABLE -- unlike the first sort in sp_orbitals, which acts on a NumPy array.
    Hence kind="stable" by default here; kind=None gives NumPy's default (unstable) sort, which is what a caller passing a
    NumPy twist to the reference's function gets (the committed fixtures were generated that way, tools/make_golden.py)."""
    kt = np.asarray(indices) + np.asarray(twist, dtype=np.float64)
    e = np.sum(kt * kt, axis=1)
    perm = e.argsort(kind=kind)
    return kt[perm], e[perm]


def sp_indices_twist(dim, Emax, twist, kind=None, order=None, kind_twist="stable"):
    """main.py:79-90: the (M, dim) float64 k-table the kernels take as `ktable`, DESCENDING twisted energy
    (row M-1 is the lowest orbital; the ground state occupies rows M-n..M-1, main.py:101).  `kind`: sort of sp_orbitals
    (NumPy default, as in the reference); `kind_twist`: sort of twist_sort (stable, as jnp.argsort in a real reference run).

    order="shipped": the table with the order of degenerate orbitals the reference's SHIPPED runs (data/, twist
    (1/4, 1/4), Emax 25 / 36 / 49) were trained with, recovered by tools/recover_tie_order.py from the exact free-fermion
    answer of the shipped pretraining runs.  Use it to replay shipped checkpoints: with today's NumPy order
    the shipped VAN parameters are off by up to 13 in log p (n = 57)."""
    idx, _ = sp_orbitals(dim, Emax, kind)
    kt, e = twist_sort(idx, twist, kind_twist)
    kt, e = np.ascontiguousarray(kt[::-1], dtype=np.float64), np.ascontiguousarray(e[::-1], dtype=np.float64)
    if order is None:
        return kt, e
    if order != "shipped":
        raise ValueError("order must be None or 'shipped'")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "shipped_ktables.npz")
    z = np.load(path)
    key = "Emax%d" % Emax
    if dim != 2 or key not in z.files or not np.allclose(np.asarray(twist, np.float64), z["twist"]):
        raise ValueError("no recovered table for dim=%d Emax=%d twist=%s (shipped runs: dim 2, twist (1/4,1/4), Emax 25/36/49)"
                         % (dim, Emax, tuple(twist)))
    rec = np.ascontiguousarray(z[key], dtype=np.float64)
    # same orbitals, same energies row by row: only degenerate ones are permuted
    if rec.shape != kt.shape or not np.allclose((rec ** 2).sum(1), e) or sorted(map(tuple, rec)) != sorted(map(tuple, kt)):
        raise RuntimeError("shipped_ktables.npz is inconsistent with sp_orbitals/twist_sort")
    return rec, e


def ground_state_indices(M, n):
    """main.py:101 (`jnp.arange(num_states - n, num_states)`): the n lowest twisted orbitals."""
    return np.arange(M - n, M, dtype=np.int32)


# ----------------------------------------------------------------------------------------------------
# file names
# ----------------------------------------------------------------------------------------------------


def run_dirname(folder, *, n, dim, rs, Theta, Emax, twist, nlayers, modelsize, nheads, nhidden, depth, spsize, tpsize,
                Gmax, kappa, mc_therm, mc_steps, mc_stddev, batch, num_devices, acc_steps, hutchinson=False,
                sr=True, damping=1e-3, max_norm=1e-3, lr=1e-3):
    """The run directory main.py:194-206 derives from its flags (every hyper-parameter is encoded in the name)."""
    s = folder + "n_%d_dim_%d_rs_%.1f_Theta_%.2f" % (n, dim, rs, Theta)
    s += "_Emax_%d" % Emax
    s += ("_twist" + "_%.3f" * dim) % tuple(twist)
    s += "_nlayers_%d_modelsize_%d_nheads_%d_nhidden_%d" % (nlayers, modelsize, nheads, nhidden)
    s += "_depth_%d_spsize_%d_tpsize_%d" % (depth, spsize, tpsize)
    s += "_Gmax_%d_kappa_%d" % (Gmax, kappa)
    s += "_mctherm_%d_mcsteps_%d_mcstddev_%.2f" % (mc_therm, mc_steps, mc_stddev)
    s += "_hutchinson" if hutchinson else ""
    s += ("_damping_%.5f_maxnorm_%.5f" % (damping, max_norm)) if sr else ("_lr_%.3f" % lr)
    s += "_batch_%d_ndevices_%d_accsteps_%d" % (batch, num_devices, acc_steps)
    return s


def parse_run_dirname(path):
    """Inverse of run_dirname for the fields the hot path needs -> dict(n, dim, rs, Theta, Emax, twist, depth, spsize,
    tpsize, Gmax, kappa, mc_therm, mc_steps, mc_stddev, batch, num_devices, acc_steps)."""
    import re
    name = os.path.basename(os.path.normpath(path))
    out = {}

    def grab(key, pat, conv, field=None):
        m = re.search(r"(?:^|_)%s_(%s)(?=_|$)" % (key, pat), name)
        if m:
            out[field or key] = conv(m.group(1))
    num = r"-?\d+(?:\.\d+)?"
    for key, conv in (("n", int), ("dim", int), ("rs", float), ("Theta", float), ("Emax", int), ("depth", int),
                      ("spsize", int), ("tpsize", int), ("Gmax", int), ("kappa", int), ("batch", int), ("accsteps", int),
                      ("ndevices", int), ("mctherm", int), ("mcsteps", int), ("mcstddev", float)):
        grab(key, num, conv)
    for old, new in (("accsteps", "acc_steps"), ("ndevices", "num_devices"), ("mctherm", "mc_therm"),
                     ("mcsteps", "mc_steps"), ("mcstddev", "mc_stddev")):
        if old in out:
            out[new] = out.pop(old)
    m = re.search(r"_twist((?:_%s)+?)_nlayers" % num, name)
    if m:
        out["twist"] = tuple(float(t) for t in m.group(1).strip("_").split("_"))
    return out


def ckpt_filename(epoch, path):
    """src/checkpoint.py:7-8."""
    return os.path.join(path, "epoch_%06d.pkl" % epoch)


def pretrained_model_filename(freefermion_path):
    """src/checkpoint.py:4-5."""
    return os.path.join(freefermion_path, "params_van.pkl")


# ----------------------------------------------------------------------------------------------------
# checkpoints
# ----------------------------------------------------------------------------------------------------


class _Opaque(tuple):
    """Stand-in for pickled objects of libraries that are not installed (optax optimiser states): kept as an empty tuple
    subclass so that the dict loads; the hot path never looks inside `opt_state`.  save_data refuses to write one (the
    re-saved checkpoint could not be resumed by the reference's optimiser): pass a real opt_state or drop the key."""

    def __new__(cls, *a, **k):
        return super().__new__(cls)

    def __setstate__(self, state):
        pass


def _rebuild_jax_array(fun, args, arr_state, aval_state=None):
    # jax pickles an Array as (numpy reconstruct fn, args, ndarray state, aval state): rebuild the ndarray, drop the aval
    arr = fun(*args)
    arr.__setstate__(arr_state)
    return arr


def _flat_mapping(*args, **kwargs):
    """haiku's FlatMapping (immutable dict of parameters) -> plain dict."""
    if args and isinstance(args[0], dict):
        return dict(args[0], **kwargs)
    return dict(*args, **kwargs)


# what a reference checkpoint legitimately contains: numpy arrays / scalars / dtypes and plain containers.  Everything else
# is refused -- unpickling must not be able to import arbitrary callables (a pickle is code).
_SAFE = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
    ("collections", "OrderedDict"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"),
    ("builtins", "frozenset"), ("builtins", "complex"), ("builtins", "slice"), ("builtins", "bytearray"),
}


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split(".")[0]
        if root in ("jax", "jaxlib"):
            return _rebuild_jax_array if name == "_reconstruct_array" else _Opaque
        if root == "haiku" and name == "FlatMapping":
            return _flat_mapping
        if root in ("optax", "haiku", "flax", "chex"):
            return _Opaque
        if (module, name) in _SAFE:
            return super().find_class(module, name)
        raise pickle.UnpicklingError("refusing to unpickle %s.%s: not part of a CoulombGas checkpoint "
                                     "(numpy arrays, plain containers, jax/optax/haiku stand-ins)" % (module, name))


def _numpyfy(t):
    if isinstance(t, dict):
        return {k: _numpyfy(v) for k, v in t.items()}
    if isinstance(t, _Opaque):
        return t
    if isinstance(t, (list, tuple)):
        return type(t)(_numpyfy(v) for v in t) if type(t) in (list, tuple) else t
    if hasattr(t, "__array__"):
        return np.asarray(t)
    return t


def load_data(filename):
    """src/checkpoint.py:10-13 without jax/optax/haiku: returns the pickled object with every array as numpy.
    For `epoch_*.pkl` that is {keys (ndev,2) u32, x (ndev,B/ndev,n,dim) f64, params_van, params_flow, opt_state}
    (main.py:377-381); for `params_van.pkl` the VAN parameter dict."""
    with open(filename, "rb") as f:
        return _numpyfy(_RefUnpickler(f).load())


def loads(buf):
    return _numpyfy(_RefUnpickler(io.BytesIO(buf)).load())


def save_data(data, filename):
    """src/checkpoint.py:15-17.  Arrays are written as numpy (torch tensors are moved to host first), which the
    reference's `pickle.load` + `src.shard` / `src.replicate` accept unchanged (main.py:218-223)."""
    def host(t):
        if isinstance(t, _Opaque):
            raise ValueError("save_data: the object holds a placeholder for a pickled optax/haiku object that was dropped on "
                             "load (e.g. opt_state): pass the real object, an empty tuple, or remove the key explicitly")
        if isinstance(t, dict):
            return {k: host(v) for k, v in t.items()}
        if type(t) in (list, tuple):
            return type(t)(host(v) for v in t)
        if hasattr(t, "detach") and hasattr(t, "cpu"):
            return t.detach().cpu().numpy()
        return t
    tmp = filename + ".tmp"
    with open(tmp, "wb") as f:
        pickle.dump(host(data), f)
    os.replace(tmp, filename)


def shard_walkers(x, rank, world_size):
    """The walker shard of one rank.  The reference stores x as (ndev, B/ndev, n, dim) (main.py:236-238); the global
    batch is its row-major flattening, and rank r of `world_size` takes the contiguous block r (main.py:234-238 with
    ndev = world_size).  Re-sharding a checkpoint written with a different device count is therefore a reshape."""
    x = np.asarray(x)
    if x.ndim == 4:
        x = x.reshape(-1, x.shape[2], x.shape[3])
    if x.ndim != 3:
        raise ValueError("x must be (ndev, B/ndev, n, dim) or (B, n, dim)")
    B = x.shape[0]
    if B % world_size != 0:
        raise ValueError("Batch size must be divisible by the number of GPU devices. Got batch = %d for %d devices now."
                         % (B, world_size))           # main.py:231-233
    per = B // world_size
    return np.ascontiguousarray(x[rank * per:(rank + 1) * per])


def gather_walkers(shards):
    """Inverse of shard_walkers for writing a checkpoint: list of per-rank (B/N, n, dim) -> (N, B/N, n, dim)."""
    return np.stack([np.asarray(s) for s in shards], axis=0)


# ----------------------------------------------------------------------------------------------------
# data.txt
# ----------------------------------------------------------------------------------------------------
_OBS = ("F", "E", "K", "V", "S")


def observable_std(mean, mean2, batch, acc_steps):
    """main.py:355-359."""
    return float(np.sqrt((mean2 - mean ** 2) / (batch * acc_steps)))


def format_data_line(epoch, *, F, F_std, E, E_std, K, K_std, V, V_std, S, S_std, accept_rate, rs):
    """One line of data.txt (main.py:369-374): energies in Ry/rs^2 are divided by rs^2 here, as main.py does."""
    r2 = rs ** 2
    return ("%6d" + "  %.6f" * 10 + "  %.4f" + "\n") % (epoch, F / r2, F_std / r2, E / r2, E_std / r2, K / r2, K_std / r2,
                                                        V / r2, V_std / r2, S, S_std, accept_rate)


def read_data_txt(filename):
    """-> dict of columns: epoch, F, F_std, E, E_std, K, K_std, V, V_std, S, S_std, accept_rate (as logged, i.e. /rs^2)."""
    rows = np.loadtxt(filename, ndmin=2)
    names = ["epoch"] + [n + s for n in _OBS for s in ("", "_std")] + ["accept_rate"]
    if rows.shape[1] != len(names):
        raise ValueError("%s: expected %d columns, found %d" % (filename, len(names), rows.shape[1]))
    return {k: rows[:, i] for i, k in enumerate(names)}
