import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kat():
    return np.load(os.path.join(GOLDEN, "kat.npz"))


@pytest.fixture(scope="session")
def ft2_ray():
    return np.load(os.path.join(GOLDEN, "ft2_ray.npz"))


@pytest.fixture(scope="session")
def ft2_fan():
    return np.load(os.path.join(GOLDEN, "ft2_fan.npz"))


@pytest.fixture(scope="session")
def rxyten_fan():
    return np.load(os.path.join(GOLDEN, "rxyten_fan.npz"))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.port()   # builds libsidtrace_oracle.so on demand (gcc only)
    return oracle


@pytest.fixture(scope="session")
def ft2_cross():
    return np.load(os.path.join(GOLDEN, "ft2_cross.npz"))


@pytest.fixture(scope="session")
def beam_fix():
    return np.load(os.path.join(GOLDEN, "beam.npz"))


@pytest.fixture(scope="session")
def ft2_phases_fix():
    return np.load(os.path.join(GOLDEN, "ft2_phases.npz"))


@pytest.fixture(scope="session")
def long_rays_fix():
    return np.load(os.path.join(GOLDEN, "long_rays.npz"))
