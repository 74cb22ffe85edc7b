import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
FOREST = os.path.join(GOLDEN, "forest_shared.dat")
CONFIG = os.path.join(ROOT, "resources", "keyframe_config.json")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def cv_golden():
    return np.load(os.path.join(GOLDEN, "cv_golden.npz"))


@pytest.fixture(scope="session")
def ref_golden():
    return np.load(os.path.join(GOLDEN, "ref_golden.npz"))


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build(ref=True)
    oracle.set_threads(min(8, os.cpu_count() or 1))
    return oracle
