import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "timeout(seconds): per-test limit (pytest-timeout; a no-op marker where the plugin is absent)")


from coulombgas_b200.fixtures import load_case, load_van, CASES_2D  # noqa: E402,F401


@pytest.fixture(scope="session")
def cases():
    cache = {}

    def get(tag):
        if tag not in cache:
            cache[tag] = load_case(tag)
        return cache[tag]
    return get
