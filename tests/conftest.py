import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


_OPS = {}


def get_sw(ref_level, **kw):
    """Session cache of oracle-assembled operators (test infrastructure)."""
    from oracle.assemble_sw import assemble_sw
    key = (ref_level, tuple(sorted(kw.items())))
    if key not in _OPS:
        _OPS[key] = assemble_sw(ref_level, **kw)
    return _OPS[key]


@pytest.fixture(scope="session")
def sw3():
    return get_sw(3)


@pytest.fixture(scope="session")
def sw2():
    return get_sw(2)
