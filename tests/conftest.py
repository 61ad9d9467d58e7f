import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "ref_ops.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_ops.npz not generated yet (oracle/make_golden.py on a GPU)")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own CUDA extension rebuilt for sm_100a (oracle/_ref), or None."""
    try:
        from oracle.build_ref import load_ref
        return load_ref()
    except Exception:
        return None
