"""Synthetic parameter vectors following the reference's init law, from a numpy Generator (the reference seeds from the clock).

ref: RBM ctor gpu/include/impl_neural_quantum_state.cuh:30-48  -- W = 0.1 (g + i g'), g ~ N(0, 1/(N+M)); a = 0; b = 0.1 (g + i g'), g ~ N(0, 1/M)
     FFNN ctor :766-783                                         -- W1 = g + 0.1 i g', g ~ N(0, 1/(N+M)); b1 = 0; w1o = g + 0.1 i g', g ~ N(0, 1/M)
Layout = the reference's `variables_`: RBM [W (i*M+j) | a | b], FFNN [W1 (i*M+j) | b1 | w1o].
"""
from __future__ import annotations

import math

import numpy as np


def n_variables(model: str, N: int, M: int) -> int:
    return N * M + N + M if model == "rbm" else N * M + 2 * M


def reference_init(model: str, N: int, M: int, rng: np.random.Generator) -> np.ndarray:
    sw, sm = math.sqrt(1.0 / (N + M)), math.sqrt(1.0 / M)
    if model == "rbm":
        W = 0.1 * (rng.normal(0, sw, (N, M)) + 1j * rng.normal(0, sw, (N, M)))
        a = np.zeros(N, dtype=np.complex128)
        b = 0.1 * (rng.normal(0, sm, M) + 1j * rng.normal(0, sm, M))
        return np.concatenate([W.ravel(), a, b]).astype(np.complex128)
    if model == "ffnn":
        W = rng.normal(0, sw, (N, M)) + 0.1j * rng.normal(0, sw, (N, M))
        b1 = np.zeros(M, dtype=np.complex128)
        w1o = rng.normal(0, sm, M) + 0.1j * rng.normal(0, sm, M)
        return np.concatenate([W.ravel(), b1, w1o]).astype(np.complex128)
    raise ValueError(model)
