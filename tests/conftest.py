import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device and the built library: skip them (instead of erroring) where either is missing."""
    import torch
    lib = os.path.join(ROOT, "inverseproblemwithdiffusionmodel_b200", "libipdm_b200.so")
    why = None
    if not torch.cuda.is_available():
        why = "no CUDA device"
    elif not os.path.exists(lib):
        why = "libipdm_b200.so is not built"
    if why:
        skip = pytest.mark.skip(reason=why)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def rel_l2(a, b):
    """relative L2 error ||a-b|| / ||b|| on torch tensors or numpy arrays."""
    import torch
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    num = torch.linalg.vector_norm((a.to(b.dtype) - b).reshape(-1))
    den = torch.linalg.vector_norm(b.reshape(-1))
    return float(num / den) if float(den) > 0 else float(num)
