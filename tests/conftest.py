import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class Case:
    """A committed golden fixture (tests/golden/<name>/, written by make_golden.py)."""

    def __init__(self, name):
        import kore_oracle as ko
        from kore_b200 import chain
        d = os.path.join(GOLDEN, name)
        self.name = name
        self.meta = json.load(open(os.path.join(d, "meta.json")))
        self.A = ko.load_csr(os.path.join(d, "A.npz"))
        self.B = ko.load_csr(os.path.join(d, "B.npz")) if os.path.exists(os.path.join(d, "B.npz")) else None
        self.bf = (ko.load_csr(os.path.join(d, "B_forced.npz"))
                   if os.path.exists(os.path.join(d, "B_forced.npz")) else None)
        self.oracle = dict(np.load(os.path.join(d, "oracle.npz")))
        m = self.meta
        self.tau = complex(m["rtau"], m["itau"])
        self.perm, self.nodeptr = chain.chain_from_params(
            m["N1"], m["m"], m["lmax"], m["symm"], m["symmB0"], m["hydro"], m["magnetic"],
            m["thermal"], m["compositional"])
        self.n = self.A.shape[0]


_cache = {}


def load_case(name):
    if name not in _cache:
        _cache[name] = Case(name)
    return _cache[name]


@pytest.fixture(scope="session")
def lib():
    from kore_b200 import lib as L
    if not os.path.exists(L.LIB_PATH):
        L.build()
    return L
