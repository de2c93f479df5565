import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_err(got, ref):
    """The parity metric of SURVEY.md s8(c): (max-abs error / max-abs ref, Frobenius relative error)."""
    import numpy as np

    got = np.asarray(got)
    ref = np.asarray(ref)
    return float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))), float(
        np.linalg.norm(got - ref) / np.linalg.norm(ref))


@pytest.fixture(scope="session")
def relerr():
    return rel_err
