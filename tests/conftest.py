import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree (/root/reference, or oracle/_ref staged by oracle/make_ref.sh)")


def pytest_collection_modifyitems(config, items):
    has_gpu = torch.cuda.is_available()
    from oracle import refload

    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not refload.available():
            item.add_marker(pytest.mark.skip(reason="no reference tree (/root/reference or oracle/_ref)"))


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))
