import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def rsb():
    """The product binding (radix-sort_b200/capi.py over librsb200.so)."""
    return importlib.import_module("radix-sort_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))


@pytest.fixture(scope="session")
def golden_digests():
    with open(os.path.join(ROOT, "tests", "golden", "golden_digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gpu(rsb):
    """(ctx, stream) on cuda:0 through the C ABI; fails loudly when the library or the GPU is missing."""
    ctx = rsb.Context(0)
    st = rsb.Stream(ctx)
    yield ctx, st
    st.close()
    ctx.close()
