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


def pytest_sessionstart(session):
    """The .so is a build artefact (git-ignored).  If it is missing and nvcc is around, build it in-tree once so
    that the ABI tests do not depend on the order in which the driver runs build() and pytest."""
    so = os.path.join(ROOT, "6dgs_b200", "csrc", "libsixdgs.so")
    if not os.path.exists(so):
        import importlib.util
        import shutil

        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            spec = importlib.util.spec_from_file_location("sixdgs_build", os.path.join(ROOT, "6dgs_b200", "csrc", "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: ((torch.from_numpy(z[k]) if z[k].dtype.kind in "fiub" else z[k]) if z[k].ndim else z[k].item()) for k in z.files}


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
