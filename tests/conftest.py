import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def P():
    import prgpu_loader
    return prgpu_loader.load()


@pytest.fixture(scope="session")
def O():
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def synth(P):
    return P.synth


@pytest.fixture(scope="session")
def world_arrays(synth):
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3)
    return dh, link, xyzr, sph, box


@pytest.fixture(scope="session")
def oracle_world(O, world_arrays):
    return O.World(*world_arrays)


@pytest.fixture(scope="session")
def gpu_world(P, world_arrays):
    return P.World(*world_arrays)
