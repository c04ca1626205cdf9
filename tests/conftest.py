import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Compiles oracle/q4_oracle.c with gcc on first use."""
    from oracle import q4_oracle

    q4_oracle.lib()
    return q4_oracle


@pytest.fixture(scope="session")
def golden():
    """Loader for tests/golden/*.npz (outputs of the reference's own kernels on a B200, see make_golden.py)."""
    import numpy as np

    def load(name):
        path = os.path.join(GOLDEN, name + ".npz")
        if not os.path.exists(path):
            pytest.fail(f"golden fixture {path} is missing (generate with tests/golden/make_golden.py on a GPU box)")
        return np.load(path)

    return load


def iter_cases(npz):
    """Yield (key_prefix, meta list) for the cNNN_* groups of a golden file."""
    keys = sorted(k[:-5] for k in npz.files if k.endswith("_meta"))
    for k in keys:
        yield k, [str(v) for v in npz[k + "_meta"]]
