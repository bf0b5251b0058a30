import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def lib():
    """The CUDA library, loaded (not built) -- GPU tests must hit the native path or fail."""
    from gr_ais_b200 import binding as B
    return B


@pytest.fixture(scope="session")
def templates(oracle):
    import numpy as np
    return {
        120: oracle.gmsk_template_bits(np.array([1, 1, 0, 0] * 6, np.uint8)),
        140: oracle.gmsk_template_bits(np.array([1, 1, 0, 0] * 7, np.uint8)),
        1120: oracle.gmsk_template_packed(np.array([1, 1, 0, 0] * 7, np.uint8)),
    }
