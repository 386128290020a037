import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by running the reference's own host functions (oracle/make_golden.py)."""
    with np.load(os.path.join(ROOT, "tests", "golden", "host_golden.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def drs():
    import drs_b200
    return drs_b200


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
