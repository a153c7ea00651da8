import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "c1_regression.npz"))


def pytest_sessionfinish(session, exitstatus):
    """tests/test_gpu_checked_build.py runs a sub-session against the debug build of the kernels
    (libfabber_cuda_checked.so, index checks compiled in) and asks for the kernels' failure counters here."""
    path = os.environ.get("FABBER_CHECK_REPORT")
    if not path:
        return
    import ctypes as C
    import json

    from fabber_core_b200 import device

    L = device.lib()
    out = (C.c_ulonglong * 2)()
    L.fabber_cuda_check_report.argtypes = [C.POINTER(C.c_ulonglong)]
    L.fabber_cuda_check_report.restype = C.c_int
    rc = L.fabber_cuda_check_report(out)
    with open(path, "w") as f:
        json.dump({"compiled_in": rc, "failures": int(out[0]), "site": int(out[1]), "exitstatus": int(exitstatus)}, f)
