import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["rbm_obc", "rbm_pbc_odd_m", "rbm_seq_custom", "ffnn_obc"]


# the tied-variable ansaetze of the reference's CPU tree (RBMTrSymm, FFNNTrSymm on the periodic chain; M = expanded width)
GOLDEN_TIED_CASES = ["rbmtrsymm_pbc", "ffnntrsymm_pbc"]


def _load_golden(name):
    import numpy as np
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    d["name"] = name
    for k in ("model", "order"):
        d[k] = str(d[k])
    for k in ("N", "M", "K", "pbc", "n_warm", "n_sr"):
        d[k] = int(d[k])
    for k in ("h", "J", "alpha", "lr", "sm_lambda"):
        d[k] = float(d[k])
    return d


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return _load_golden(request.param)


@pytest.fixture(params=GOLDEN_TIED_CASES)
def golden_tied(request):
    return _load_golden(request.param)
