import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def og():
    """the CPU oracle (test infrastructure only)"""
    from oracle import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """a geopolars_b200 Context on cuda:0 — created only by gpu-marked tests"""
    from geopolars_b200.engine import Context

    c = Context(0)
    yield c
    c.synchronize()


def to_og(og, a):
    """geopolars_b200.GeoArrowArray -> oracle OGArray (same buffers)"""
    return og.OGArray(int(a.type), a.xy, geom_off=a.geom_off, part_off=a.part_off, ring_off=a.ring_off, valid=a.valid)


@pytest.fixture(scope="session")
def conv(og):
    return lambda a: to_og(og, a)


def rel_close(a, b, tol=1e-9):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    denom = np.maximum(np.abs(a), np.abs(b))
    denom[denom == 0] = 1.0
    ok = (np.abs(a - b) <= tol * denom) | both_nan
    return bool(ok.all())
