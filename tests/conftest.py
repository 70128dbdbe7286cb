import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pkg(name=None):
    base = "lins---lidar-inertial-slam_b200"
    return importlib.import_module(base if name is None else base + "." + name)


@pytest.fixture(scope="session")
def defs():
    return pkg("ctypes_defs")


@pytest.fixture(scope="session")
def synth():
    m = pkg("synth")
    m.build()
    return m


@pytest.fixture(scope="session")
def ob():
    from oracle import oracle_binding

    oracle_binding.build()
    return oracle_binding


@pytest.fixture(scope="session")
def capi():
    return pkg("capi")


@pytest.fixture(scope="session")
def golden_batch(defs):
    """The committed config-1 unit + three config-3 units (inputs)."""
    return defs.Batch.load(os.path.join(GOLDEN, "units_inputs.npz"))


@pytest.fixture(scope="session")
def golden_out():
    return np.load(os.path.join(GOLDEN, "units_oracle_outputs.npz"))


def have_gpu():
    try:
        import ctypes

        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cuda.cuInit(0) != 0:
            return False
        cuda.cuDeviceGetCount(ctypes.byref(n))
        return n.value > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def gpu(capi):
    """A LinsGpu context; GPU tests must FAIL (not skip) when the library cannot run on a GPU box."""
    return capi.LinsGpu()
