import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def engine():
    from methylasso_b200 import api
    eng = api.Engine()
    yield eng
    eng.close()


def pooled_series(table):
    """FAST-path pseudodata of a one-condition table: pooled methylation and coverage per position."""
    u, inv = np.unique(table["pos"], return_inverse=True)
    cov = np.bincount(inv, table["coverage"], u.size)
    nm = np.bincount(inv, table["Nmeth"], u.size)
    return nm / cov, cov.astype(float)
