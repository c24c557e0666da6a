import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kbbq-py_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["mixed_r3", "illumina_r1", "tails_r2_second", "odd_len_unpaired", "tiny_r2"]
TABLE_KEYS = ("meanq", "rg_errs", "rg_total", "q_errs", "q_total", "pos_errs", "pos_total",
              "dinuc_errs", "dinuc_total")
DELTA_KEYS = ("rgdq", "qdq", "posdq", "dindq")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_case(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(params=CASES)
def golden_case(request):
    return load_case(request.param)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle
