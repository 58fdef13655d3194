"""CPU: the NumPy restatement of Philox4x32-10 (oracle/philox.py) reproduces the Random123 known-answer vectors; the GPU
tests (tests/test_gpu_rng.py) then hold the kernels' generator against it."""
import numpy as np
from oracle import philox


def test_philox_known_answers():
    for ctr, key, out in philox.KAT:
        got = philox.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(v) for v in got) == out


def test_stream_mapping_statistics():
    u = philox.uniforms(12345, 64, 50)
    assert u.shape == (50, 64) and u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.03
    z = philox.normals(12345, 16, 10, 29)
    assert z.shape == (10, 16, 29, 2) and abs(z.mean()) < 0.05 and abs(z.std() - 1.0) < 0.05
