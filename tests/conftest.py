import os
import sys

import pytest

# Small-ndim RealNVP flows take the tensor-core kernel under HB_PATH_AUTO only from 2^20 rows per call on (below that the
# range guard's device-to-host read costs more than the kernel saves, csrc/hb_api.cu); the tests run it at every size.
os.environ.setdefault("HB_STC_MIN_ROWS", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "evidence_golden.npz"))
