#!/usr/bin/env python
"""Cost of the reference's trng::yarn2 uniform streams (csrc/yarn2.cuh) next to the in-kernel Philox generator: CUDA-event time of
one sweep per call at a BASELINE shape, sampler-only handle.  Prints one JSON line (-> profiles/r2_yarn2.md)."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neural_network_quantum_state_b200 import Engine  # noqa: E402

H, J, ALPHA = -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0
out = {}
for name, (N, M, K) in {"cfg3": (128, 256, 16384), "cfg2": (64, 128, 4096)}.items():
    res = {}
    for kind in ("philox", "yarn2"):
        e = Engine("rbm", N, M, K, H, J, ALPHA, seed=5, sampler_only=True)
        e.init_params_random(5)
        e.set_rng(kind, 5, 100 * N * K)
        e.warm_up(5)
        times = []
        for sweeps in (1, 1, 1, 10):
            e.event_record(0)
            e.do_mcmc_steps(sweeps)
            e.event_record(1)
            times.append(e.event_elapsed_ms(0, 1) / sweeps)
        res[kind] = {"ms_per_sweep_single_call": min(times[:3]), "ms_per_sweep_10_per_call": times[3], "kernel": e.kernel_variant("sweep")}
        e.close()
    res["yarn2_overhead_ms_per_sweep"] = res["yarn2"]["ms_per_sweep_single_call"] - res["philox"]["ms_per_sweep_single_call"]
    res["uniform_bytes_per_sweep"] = 8 * N * K
    out[name] = res
print(json.dumps(out))
