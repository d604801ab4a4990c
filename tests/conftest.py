import json
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
def golden():
    with open(os.path.join(GOLDEN, "small_fsim_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def small_fsim():
    from gpusimilarity_b200.fsim import read_fsim
    return read_fsim(os.path.join(GOLDEN, "small.fsim"))


@pytest.fixture(scope="session")
def small_db(small_fsim):
    return small_fsim.fingerprints()


def f32bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
