"""pytest configuration: markers, import path, shared fixtures.

  -m "not gpu"  -> oracle vs golden vectors, host logic, C-ABI symbol check (CPU only)
  -m gpu        -> parity tests proper: CUDA path through the C ABI vs the oracle
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def mcb_lib():
    """libmcb200.so, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from mc_mpi_b200 import _abi, build
    try:
        build.build()
    except RuntimeError:
        if not os.path.isfile(_abi.LIB_PATH):
            raise
    return _abi.lib()


@pytest.fixture(scope="session")
def gpu(mcb_lib):
    """a CUDA device must be present for -m gpu tests: no silent CPU fallback."""
    n = mcb_lib.mcb200_device_count()
    assert n > 0, "gpu-marked test but libmcb200 sees no CUDA device"
    return 0
