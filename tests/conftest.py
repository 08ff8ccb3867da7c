"""pytest configuration: the `gpu` marker, import paths, and the back-end seams used by the tests.

`-m "not gpu"` runs here (no GPU): oracle vs. definitions / golden vectors, host logic, C-ABI
symbol checks.  `-m gpu` runs on a B200: the CUDA path through the C ABI vs. the oracle.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle back-end (test infrastructure; built on demand with gcc)."""
    from oracle import pointnet2_oracle
    return pointnet2_oracle.OracleBackend()


@pytest.fixture()
def oracle_ops(oracle):
    """Install the oracle behind pointnet2.pointnet2 for the duration of one test."""
    from ogc_b200 import backend
    prev = backend.set_backend(oracle)
    try:
        import pointnet2.pointnet2 as ops
        yield ops
    finally:
        backend.set_backend(prev)


@pytest.fixture(scope="session")
def b200():
    """The product back-end (libogc_b200.so through the C ABI). Fails loudly if not built."""
    from ogc_b200.backend import B200Backend
    return B200Backend()
