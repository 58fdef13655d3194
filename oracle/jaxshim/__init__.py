"""NumPy-backed stand-in for the jax / haiku API subset the reference's hot-path
source files use.  TEST INFRASTRUCTURE ONLY (used by tools/make_golden.py, in the
build container where /root/reference is mounted, to execute the reference's own
src/*.py unmodified and record golden vectors).  This is NOT JAX: arrays are
numpy float64/complex128, `jacfwd` is complex-step differentiation, `vmap`/`pmap`/
`fori_loop` are Python loops, and `jax.random` is a deterministic NumPy stand-in.
"""
from .install import install, load_reference  # noqa: F401
