#!/usr/bin/env python
"""Times the structured S*v kernels in isolation (run under `ncu --metrics gpu__time_duration.sum -k regex:"ozaki|umma|dmma"`):
a cfg3-shaped engine, a few products.  NQS_ROWS_UMMA=0/1 picks the fp64 DMMA or the tcgen05 int8 rows kernel."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neural_network_quantum_state_b200 import Engine  # noqa: E402

H, J, ALPHA = -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
e = Engine("rbm", 128, 256, K, H, J, ALPHA, seed=5, structured_sv=True)
e.init_params_random(5)
e.warm_up(2)
v = np.random.default_rng(0).normal(size=e.P) + 0j
for _ in range(4):
    e.smatrix_dot(0.3, v)
print(e.kernel_variant("theta"), e.kernel_variant("sv"))
e.close()
