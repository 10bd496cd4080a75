import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O  # noqa: N812  (oracle/oracle.py — test infrastructure)
    O.build()
    return O


@pytest.fixture(scope="session")
def cs():
    """The product package with its CUDA library built and loaded (no fallback)."""
    import clonalscope_b200 as pkg
    if not os.path.exists(pkg.library_path()):
        pkg.build_library()
    pkg.lib()
    return pkg


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


SAMPLER_FIXTURES = ["P5931", "P5931_updated1", "BC_ductal2", "P5847_tumor", "P5847_allcells"]
