"""Test configuration.

`-m "not gpu"`: oracle vs known answers / golden vectors, host logic, C-ABI symbol exports.
`-m gpu`: parity of the CUDA path (through the C-ABI) against the oracle on a real B200.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def uivr():
    """The product package (directory name carries a hyphen, hence importlib)."""
    return importlib.import_module("uivr_b200")


NEE_LOG_CAPACITY = 32   # kNeeLog of csrc/uivr_pool.cuh


@pytest.fixture(scope="session")
def oracle():
    """The checker.  Its EVENT COUNTERS model the collision log of the CUDA adjoint kernel (shadow walks with more
    tentative collisions than the log holds are walked twice there, like in the reference): results are unaffected,
    but `fused_backward_counters` must book exactly the second walks the kernel really skips."""
    from oracle import oracle as O
    O.build()
    O.set_nee_log_capacity(NEE_LOG_CAPACITY)
    return O
