"""Reader for the reference's pickled checkpoints (test infrastructure only).

src/checkpoint.py:10-13 is a bare ``pickle.load``; the shipped files reference
``jax._src.array._reconstruct_array`` and ``optax._src.base.EmptyState`` which
are not importable here, so those two symbols are stubbed.
"""
import pickle
import numpy as np


class _Empty(tuple):
    def __new__(cls, *a, **k):
        return super().__new__(cls)


def _reconstruct_array(fun, args, arr_state, aval_state):
    arr = fun(*args)
    arr.__setstate__(arr_state)
    return arr


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("jax") and name == "_reconstruct_array":
            return _reconstruct_array
        if module.startswith("optax"):
            return _Empty
        if module.startswith("jax") or module.startswith("haiku"):
            return _Empty
        return super().find_class(module, name)


def load_data(filename):
    """Returns the dict {keys, x, params_van, params_flow, opt_state} (main.py:377-381)."""
    with open(filename, "rb") as f:
        data = _Unpickler(f).load()

    def to_np(t):
        if isinstance(t, dict):
            return {k: to_np(v) for k, v in t.items()}
        if isinstance(t, (list, tuple)) and not isinstance(t, _Empty):
            return type(t)(to_np(v) for v in t)
        if hasattr(t, "__array__"):
            return np.asarray(t)
        return t
    return to_np(data)
