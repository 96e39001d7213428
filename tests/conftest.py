import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (tests may import it; the product package may not)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    return importlib.import_module("sixdgs_oracle")


@pytest.fixture(scope="session")
def sx():
    """The product package (directory name starts with a digit -> importlib)."""
    return importlib.import_module("6dgs_b200")


@pytest.fixture(scope="session")
def synthetic():
    return importlib.import_module("6dgs_b200.synthetic")
