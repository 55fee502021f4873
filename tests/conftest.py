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


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    import numpy as np
    d = dict(np.load(os.path.join(GOLDEN, request.param + ".npz"), allow_pickle=False))
    d["name"] = request.param
    for k in ("model", "order"):
        d[k] = str(d[k])
    for k in ("N", "M", "K", "pbc", "n_warm", "n_sr"):
        d[k] = int(d[k])
    for k in ("h", "J", "alpha", "lr", "sm_lambda"):
        d[k] = float(d[k])
    return d
