"""Test tiers
  -m "not gpu": the oracle against golden vectors / the compiled reference, host logic, ABI surface (no GPU needed)
  -m gpu      : parity of the CUDA path (through the C ABI) against the oracle, the golden vectors and -- where
                oracle/_ref is present -- the unmodified reference itself.
"""
import os
import sys

import numpy as np
import pytest

# the virtual-rank tests keep up to eight streams with spinning device barriers in flight on one GPU: one hardware
# queue per stream, so that no stream is serialised behind another one's barrier kernel (must be set before CUDA starts)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # no test may hang the tier: device barriers give up after ~10 s each, everything else is bounded by this
    if config.pluginmanager.hasplugin("timeout"):
        for it in items:
            if it.get_closest_marker("timeout") is None:
                it.add_marker(pytest.mark.timeout(900))


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle_lib import Reference, reference_available

    if not reference_available():
        pytest.skip("oracle/_ref (the compiled reference) is not present on this box")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))

    return load


@pytest.fixture(scope="session")
def gpu():
    """(Context, module) on cuda:0; the CUDA library must be present -- no fallback."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ofhe_b200

    ctx = ofhe_b200.Context(0)
    yield ctx, ofhe_b200
    ctx.close()
