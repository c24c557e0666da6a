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


# BASELINE configs 2, 3, 4 at reference-runnable size: inputs regenerated from kbbq.synth, outputs from the reference
SYNTH_CASES = ["c2_l150_r1", "c3_l150_r8", "c4_l250_r32"]


def load_case(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def load_synth_case(name):
    """-> dict like the other goldens: the reads are regenerated with kbbq.synth (and checked against the digest
    make_golden.py stored), the read-group numbers are the reference's first-seen ones."""
    import hashlib
    from kbbq import synth
    g = dict(load_case(name))
    N, L = int(g["N"]), int(g["L"])
    seq, qual, corr, rgk, second = synth.synth_reads(int(g["seed"]), 0, N, L, int(g["R_synth"]))
    digest = hashlib.sha256(seq.tobytes() + qual.tobytes() + corr.tobytes() + rgk.tobytes() + second.tobytes()).hexdigest()
    assert digest == str(g["input_sha256"]), "kbbq.synth no longer produces the bytes the golden was made from"
    g.update(seq=seq, qual=qual, corr=corr, second=second, outq=g["outq"].astype(np.int16))
    return g


@pytest.fixture(params=CASES)
def golden_case(request):
    return load_case(request.param)


@pytest.fixture(params=SYNTH_CASES)
def synth_golden_case(request):
    return load_synth_case(request.param)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle
