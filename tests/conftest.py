import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


def bits_to_f32(bits):
    return np.array(bits, dtype=np.uint32).view(np.float32)


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree libraries exist (the driver normally runs __graft_entry__.build())."""
    import __graft_entry__ as ge
    import gunrock_b200
    if not os.path.exists(gunrock_b200.LIB_PATH) or not os.path.exists(
            os.path.join(ROOT, "oracle", "libgunrock_oracle.so")):
        ge.build()
    return True
