import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (*.so are git-ignored): build them once, like __graft_entry__.build()
    from mapquik_b200 import _build
    import subprocess
    if not (os.path.exists(_build.LIB) and os.path.exists(_build.HOSTLIB)):
        _build.build_all()
    if not os.path.exists(os.path.join(ROOT, "oracle", "libmq_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def have_gpu():
    return _have_gpu()


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must fail loudly, not skip: the product has no CPU fallback.
    return


ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_dna(rng, n):
    return ACGT[rng.integers(0, 4, n)]


def revcomp(a):
    comp = np.zeros(256, np.uint8)
    comp[:] = np.arange(256, dtype=np.uint8)
    for x, y in zip(b"ACGT", b"TGCA"):
        comp[x] = y
    return comp[a][::-1].copy()
