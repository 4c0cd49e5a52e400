"""pytest configuration: registers the `gpu` marker and puts the repo's import roots on sys.path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kernel-methods-for-genomics_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


@pytest.fixture(scope="session")
def dna():
    """The 9000 challenge sequences (Xtr0,Xtr1,Xtr2,Xte0,Xte1,Xte2 in file order) as uint8 codes."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "dna9000.npz"))
    return z["codes"], z["labels"]
