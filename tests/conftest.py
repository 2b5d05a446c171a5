import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "mms_golden.npz")
    return np.load(path)


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    # building the checker is not using it: compile the C oracle (and the in-place
    # reference build when /root/reference exists) once per session.
    from oracle import refbind
    refbind.build(ref=os.path.isdir("/root/reference"), oracle=True)


def scaled_err(got, ref, floor=None):
    """GradientChecker-style error (reference test_gradient_check_util.hpp:172-175):
    |got-ref| / max(|got|, |ref|, floor), with floor defaulting to the tensor's own
    largest magnitude so that near-zero entries are judged on the tensor's scale."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if floor is None:
        floor = max(np.abs(ref).max() if ref.size else 0.0, 1e-30)
    den = np.maximum(np.maximum(np.abs(got), np.abs(ref)), floor)
    return float((np.abs(got - ref) / den).max()) if ref.size else 0.0
