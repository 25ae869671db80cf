import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def libqrag():
    """The in-tree C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    from quantum_rag_b200 import _lib
    _lib.build()
    return _lib.load()


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def piers():
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "piers_index.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible")
    from quantum_rag_b200 import _lib
    _lib.build()
    _lib.load()
    return torch.device("cuda", 0)
