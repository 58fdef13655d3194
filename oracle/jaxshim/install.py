import sys
import types
import importlib
import numpy as np
from scipy import special as _sp


# ----------------------------------------------------------------------------- jax.numpy
def _make_jnp():
    jnp = types.ModuleType("jax.numpy")
    for name in ("cos", "sin", "exp", "log", "pi", "eye", "concatenate", "stack", "where", "zeros_like",
                 "zeros", "ones", "arange", "sqrt", "floor", "rint", "abs", "triu_indices", "allclose",
                 "float64", "int32", "int64", "complex128", "sum", "mean", "clip", "tile", "asarray"):
        setattr(jnp, name, getattr(np, name))
    for name in ("tril", "vstack", "tanh", "einsum", "reshape", "argmax", "max", "float32", "ones_like", "ndarray"):
        setattr(jnp, name, getattr(np, name))
    jnp.array = lambda a, dtype=None: np.array(a, dtype=dtype)
    jnp.zeros = lambda shape, dtype=None: np.zeros(shape, dtype=dtype or np.float64).view(_AtArray)
    jnp.alltrue = np.all
    linalg = types.ModuleType("jax.numpy.linalg")

    def norm(a, axis=None):
        # analytic continuation (no abs) so complex-step differentiation goes through;
        # identical to np.linalg.norm for real input.
        if np.iscomplexobj(a):
            return np.sqrt((a * a).sum(axis=axis))
        return np.linalg.norm(a, axis=axis)
    linalg.norm = norm
    linalg.slogdet = lambda a: tuple(np.linalg.slogdet(a))
    linalg.inv = np.linalg.inv
    jnp.linalg = linalg
    return jnp


class _AtArray(np.ndarray):
    """ndarray with the functional-update accessor `a.at[idx].set(v)` (src/sampler.py:36-37)."""

    @property
    def at(self):
        arr = self

        class _At:
            def __getitem__(self, idx):
                class _Ref:
                    def set(self, v):
                        out = np.array(arr, copy=True).view(_AtArray)
                        out[idx] = v
                        return out
                return _Ref()
        return _At()


def _log_softmax(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    return x - (m + np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)))


def _softmax(x, axis=-1):
    return np.exp(_log_softmax(x, axis))


# ----------------------------------------------------------------------------- jax core
class _CustomJVP:
    def __init__(self, fun):
        self.fun = fun

    def __call__(self, *a, **k):
        return self.fun(*a, **k)

    def defjvp(self, f):
        return f


def _softplus(x):
    if np.iscomplexobj(x):
        return np.where(x.real > 0, x + np.log1p(np.exp(-x)), np.log1p(np.exp(x)))
    return np.logaddexp(x, 0.0)


def _jacfwd(f):
    def jac(x):
        x = np.asarray(x)
        h = 1e-30
        cols = []
        for k in range(x.size):
            xc = x.astype(np.complex128).copy()
            xc.reshape(-1)[k] += 1j * h
            cols.append(np.imag(f(xc)) / h)
        return np.stack(cols, axis=-1)
    return jac


def _vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        B = next(np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fun(*[a if ax is None else np.take(a, b, axis=ax) for a, ax in zip(args, axes)])
                for b in range(B)]
        if isinstance(outs[0], tuple):
            return tuple(np.stack([o[i] for o in outs], axis=out_axes) for i in range(len(outs[0])))
        return np.stack(outs, axis=out_axes)
    return mapped


def _pmap(fun=None, axis_name=None, in_axes=0, static_broadcasted_argnums=(), donate_argnums=()):
    def deco(f):
        def mapped(*args):
            static = set(static_broadcasted_argnums if isinstance(static_broadcasted_argnums, (tuple, list))
                         else (static_broadcasted_argnums,))
            axes = list(in_axes)
            ndev = next(np.shape(a)[0] for i, (a, ax) in enumerate(zip(args, axes))
                        if ax is not None and i not in static and hasattr(a, "shape"))
            outs = []
            for dev in range(ndev):
                call = [a if (i in static or axes[i] is None or a is None)
                        else _tree_map(lambda t: t[dev], a) for i, a in enumerate(args)]
                outs.append(f(*call))
            return tuple(np.stack([np.asarray(o[i]) for o in outs]) for i in range(len(outs[0])))
        return mapped
    return deco if fun is None else deco(fun)


def _jit(fun=None, static_argnums=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _fori_loop(lo, hi, body, init):
    val = init
    for i in range(lo, hi):
        val = body(i, val)
    return val


def _tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: _tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    return f(tree, *rest)


# ----------------------------------------------------------------------------- jax.random
class _Random(types.ModuleType):
    """Deterministic stand-in keyed like jax (uint32[2] keys); NOT threefry.
    Every draw is appended to `log` so golden generators can record what was consumed."""

    def __init__(self):
        super().__init__("jax.random")
        self.log = []

    @staticmethod
    def _rng(key):
        k = np.asarray(key, dtype=np.uint32).reshape(-1)
        return np.random.Generator(np.random.Philox(key=int(k[0]) << 32 | int(k[1])))

    def PRNGKey(self, seed):
        return np.array([0, seed], dtype=np.uint32)

    def split(self, key, num=2):
        return self._rng(key).integers(0, 2 ** 32, size=(num, 2), dtype=np.uint32)

    def normal(self, key, shape=()):
        v = self._rng(key).standard_normal(shape)
        self.log.append(("normal", v))
        return v

    def uniform(self, key, shape=(), minval=0.0, maxval=1.0):
        v = minval + (maxval - minval) * self._rng(key).random(shape)
        self.log.append(("uniform", v))
        return v

    def categorical(self, key, logits, axis=-1):
        """jax.random.categorical = argmax(logits + Gumbel noise); the noise is logged for the golden generators."""
        g = self._rng(key).gumbel(size=np.shape(logits))
        self.log.append(("gumbel", g))
        return np.argmax(np.asarray(logits) + g, axis=axis)


# ----------------------------------------------------------------------------- haiku
def _make_haiku(jax_random):
    hk = types.ModuleType("haiku")
    state = {"mode": None, "params": None, "rng": None, "stack": [], "counters": [{}]}

    def _snake(name):
        out = ""
        for i, c in enumerate(name):
            if c.isupper() and i and (not name[i - 1].isupper()):
                out += "_"
            out += c.lower()
        return out

    class Module:
        def __init__(self, name=None):
            base = name or _snake(type(self).__name__)
            counters = state["counters"][-1]
            idx = counters.get(base, 0)
            counters[base] = idx + 1
            local = base if idx == 0 else "%s_%d" % (base, idx)
            if state["stack"]:
                parent, in_init = state["stack"][-1]
                self.module_name = parent + ("/~/" if in_init else "/") + local
            else:
                self.module_name = local

        def __init_subclass__(cls, **kw):
            super().__init_subclass__(**kw)
            if "__init__" in cls.__dict__:
                orig_init = cls.__dict__["__init__"]

                def __init__(self, *a, __orig=orig_init, **k):
                    # children created while the subclass __init__ body runs get "/~/"
                    entered = [False]
                    real_base_init = Module.__init__

                    def base_init(s, name=None):
                        real_base_init(s, name)
                        if s is self and not entered[0]:
                            state["stack"].append((s.module_name, True))
                            state["counters"].append(self.__dict__.setdefault("_hk_counters_init", {}))
                            entered[0] = True
                    Module.__init__ = base_init
                    try:
                        __orig(self, *a, **k)
                    finally:
                        Module.__init__ = real_base_init
                        if entered[0]:
                            state["stack"].pop(); state["counters"].pop()
                cls.__init__ = __init__
            if "__call__" in cls.__dict__:
                orig_call = cls.__dict__["__call__"]

                def __call__(self, *a, __orig=orig_call, **k):
                    state["stack"].append((self.module_name, False))
                    state["counters"].append({})   # method scope has its own name counters
                    try:
                        return __orig(self, *a, **k)
                    finally:
                        state["stack"].pop(); state["counters"].pop()
                cls.__call__ = __call__

    class RandomNormal:
        def __init__(self, stddev=1.0, mean=0.0):
            self.stddev, self.mean = stddev, mean

        def __call__(self, shape, rng):
            return self.mean + self.stddev * rng.standard_normal(shape)

    class Linear(Module):
        def __init__(self, output_size, with_bias=True, w_init=None, b_init=None, name=None):
            super().__init__(name=name)
            self.output_size, self.w_init = output_size, w_init

        def __call__(self, x):
            name = self.module_name
            if state["mode"] == "init" and name not in state["params"]:
                w = self.w_init((x.shape[-1], self.output_size), state["rng"])
                state["params"][name] = {"w": w, "b": np.zeros(self.output_size)}
            p = state["params"][name]
            return x @ p["w"] + p["b"]

    class Transformed:
        def __init__(self, fn):
            self.fn = fn

        def init(self, key, *args):
            state.update(mode="init", params={}, rng=jax_random._rng(key), stack=[], counters=[{}])
            self.fn(*args)
            out, state["params"], state["mode"] = state["params"], None, None
            return out

        def apply(self, params, rng, *args):
            state.update(mode="apply", params=params, stack=[], counters=[{}])
            try:
                return self.fn(*args)
            finally:
                state["params"], state["mode"] = None, None

    class VarianceScaling:
        def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal"):
            self.scale, self.mode = scale, mode

        def __call__(self, shape, rng):
            fan = shape[0] if self.mode == "fan_in" else shape[-1]
            return np.sqrt(self.scale / max(1.0, fan)) * rng.standard_normal(shape)

    class TruncatedNormal:
        def __init__(self, stddev=1.0, mean=0.0):
            self.stddev, self.mean = stddev, mean

        def __call__(self, shape, rng):
            return self.mean + self.stddev * np.clip(rng.standard_normal(shape), -2, 2)

    def get_parameter(name, shape, dtype=None, init=None):
        mod = state["stack"][-1][0]
        if state["mode"] == "init" and name not in state["params"].setdefault(mod, {}):
            state["params"][mod][name] = init(tuple(shape), state["rng"])
        return state["params"][mod][name]

    class MultiHeadAttention(Module):
        """dm-haiku's hk.MultiHeadAttention restated from its published algorithm (the library is not installed):
        per-head `query`/`key`/`value` projections, logits q.k/sqrt(key_size), masked entries -> -1e30, softmax over
        keys, heads concatenated, final `linear` projection to model_size = num_heads * key_size."""

        def __init__(self, num_heads, key_size, w_init_scale=None, value_size=None, model_size=None, name=None):
            super().__init__(name=name)
            self.num_heads, self.key_size = num_heads, key_size
            self.value_size = value_size or key_size
            self.model_size = model_size or key_size * num_heads
            self.w_init = VarianceScaling(w_init_scale if w_init_scale is not None else 1.0)

        def __call__(self, query, key, value, mask=None):
            def proj(x, size, name):
                y = Linear(self.num_heads * size, w_init=self.w_init, name=name)(x)
                return y.reshape(x.shape[:-1] + (self.num_heads, size))
            q, k, v = proj(query, self.key_size, "query"), proj(key, self.key_size, "key"), proj(value, self.value_size, "value")
            logits = np.einsum("...thd,...Thd->...htT", q, k) / np.sqrt(self.key_size)
            if mask is not None:
                logits = np.where(mask, logits, -1e30)
            w = _softmax(logits)
            att = np.einsum("...htT,...Thd->...thd", w, v)
            att = att.reshape(query.shape[:-1] + (-1,))
            return Linear(self.model_size, w_init=self.w_init)(att)

    inits = types.ModuleType("haiku.initializers")
    inits.RandomNormal = RandomNormal
    inits.VarianceScaling, inits.TruncatedNormal = VarianceScaling, TruncatedNormal
    hk.MultiHeadAttention, hk.get_parameter = MultiHeadAttention, get_parameter
    hk.Module, hk.Linear, hk.initializers, hk.transform = Module, Linear, inits, Transformed
    return hk


# ----------------------------------------------------------------------------- install
def install():
    """Put the stand-in `jax`, `jax.numpy`, `jax.scipy.special`, `jax.config`, `haiku`
    into sys.modules.  Returns the fake jax module."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "_is_cgb_shim", False):
        return sys.modules["jax"]
    jax = types.ModuleType("jax")
    jax._is_cgb_shim = True
    jnp = _make_jnp()
    jax.numpy = jnp
    jax.random = _Random()
    nn = types.ModuleType("jax.nn"); nn.softplus = _softplus; nn.log_softmax = _log_softmax; nn.softmax = _softmax; jax.nn = nn
    lax = types.ModuleType("jax.lax")
    lax.fori_loop = _fori_loop
    lax.pmean = lambda x, axis_name=None: x       # one device
    jax.lax = lax
    jscipy = types.ModuleType("jax.scipy"); jspecial = types.ModuleType("jax.scipy.special")
    jspecial.erfc = _sp.erfc; jscipy.special = jspecial; jax.scipy = jscipy
    cfgmod = types.ModuleType("jax.config")
    cfgmod.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.config = cfgmod
    jax.custom_jvp = _CustomJVP
    jax.jacfwd = _jacfwd
    jax.vmap = _vmap
    jax.pmap = _pmap
    jax.jit = _jit
    jax.tree_map = _tree_map
    jax.grad = lambda f, *a, **k: (lambda *aa, **kk: (_ for _ in ()).throw(NotImplementedError("grad is not part of the stand-in")))
    jax.device_count = lambda: 1
    hk = _make_haiku(jax.random)
    for name, mod in {"jax": jax, "jax.numpy": jnp, "jax.random": jax.random, "jax.nn": nn, "jax.lax": lax,
                      "jax.scipy": jscipy, "jax.scipy.special": jspecial, "jax.config": cfgmod,
                      "haiku": hk}.items():
        sys.modules[name] = mod
    return jax


def load_reference(root="/root/reference", names=("orbitals", "flow", "slater", "logpsi", "MCMC",
                                                    "potential", "VMC")):
    """Import the reference's src/<name>.py unmodified (without running src/__init__.py, which
    pulls in the optimiser/sampler stack) under the stand-in.  Returns {name: module}."""
    install()
    pkg = types.ModuleType("cgref_src")
    pkg.__path__ = [root + "/src"]
    sys.modules["cgref_src"] = pkg
    return {n: importlib.import_module("cgref_src." + n) for n in names}
