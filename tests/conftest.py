import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product's Python binding (rust-brotli-decompressor_b200/__init__.py)."""
    return importlib.import_module("rust-brotli-decompressor_b200")


@pytest.fixture(scope="session")
def corpus():
    return importlib.import_module("tools.corpus")


@pytest.fixture(scope="session")
def oracle():
    import helpers
    return helpers.Oracle()


@pytest.fixture(scope="session")
def hostsim():
    import helpers
    return helpers.HostSim()


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    """libbrotli_b200.so on a machine with a GPU; GPU tests must never silently run anything else."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    L = pkg.lib()
    assert L.BrotliB200ResidentWarps() > 0, L.BrotliB200LastError()
    # The library routes batches below ~6000 streams straight to the warp-per-stream kernel (they are faster there).  The
    # parity tests use small batches and must exercise the lane-per-stream kernel too: switch the threshold off (the
    # default routing has its own test, test_gpu_parity.py::test_batch_size_routing, and "exact_only" runs everything
    # through the other kernel).
    assert pkg.set_tuning("lane_min_streams", 0)
    return L
