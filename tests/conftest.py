import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_fixture(name):
    """-> (ModelRecord, dict of golden arrays) from tests/golden/<name>.npz"""
    from pcetk_b200.formats import ModelRecord

    data = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return ModelRecord.from_npz_dict(data), data


@pytest.fixture(scope="session")
def port():
    import oracle

    oracle.build()
    return oracle.Port()
