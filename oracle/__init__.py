"""CPU oracle for the CoulombGas forward-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or as the CPU baseline — never on the product path.  The product
(``coulombgas_b200``) fails loudly when its CUDA library is missing; it has no
CPU fallback and never imports this package.

What it is: a float64 restatement, as written, of the reference algorithm
(fermiflow/CoulombGas, paths relative to the reference tree):

* ``tables.py``     – src/orbitals.py:22-55, src/potential.py:7-34, main.py:63-90,170-171
* ``flow.py``       – src/flow.py:5-55
* ``logpsi.py``     – src/slater.py:4-19, src/logpsi.py:7-33,174-181
* ``mcmc.py``       – src/MCMC.py:6-40, src/VMC.py:23-24
* ``potential.py``  – src/potential.py:36-77
* ``checkpoint.py`` – src/checkpoint.py:10-13 (reader for the shipped epoch_*.pkl)

Pinning (see tests/test_oracle_golden.py, tools/make_golden.py):
the reference's JAX/haiku stack is not installable in this image, so the
reference *source files* were executed unmodified from /root/reference under a
NumPy-backed stand-in for the jax/haiku API subset they use
(``oracle/jaxshim``; forward-mode Jacobian by complex-step differentiation).
The outputs are committed as ``tests/golden/*.npz`` and the oracle is checked
against every one of them, plus the Madelung constant, |G| and orbital counts,
the property tests the reference's own tests assert, and the statistical pins
from the shipped ``data.txt`` logs.  Boundaries that remain *parity unpinned*:
JAX's threefry PRNG streams (the oracle and the kernels take proposal noise and
uniforms as inputs) and jaxlib's LAPACK/cuSOLVER rounding (NumPy/torch LAPACK
is used instead; slogdet is mathematically pivot-order independent).
"""
