"""pytest configuration: registers the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` tests run on CPU (oracle vs golden vectors, host logic, C-ABI symbol table,
world_size-2 gloo sharding); ``-m gpu`` tests are the CUDA-vs-oracle parity tests proper and
need a B200.
"""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import ref_oracle

    ref_oracle.build()
    return ref_oracle
