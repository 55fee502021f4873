"""CPU restatement of trng::yarn2 + trng::uniform01_dist<double> -- TEST INFRASTRUCTURE (the checker of csrc/yarn2.cuh), never
imported by the product.

The generator is the reference's source of Metropolis uniforms (gpu/include/trng4cuda.cuh:14-65; CPU build
cpu/include/impl_mcmc_sampler.hpp:19-23,52): one engine per chain, `seed(seedNumber); jump(2ul*seedDistance*k)`, one
`uniform01_dist` draw per proposal.  It lives in TRNG4, a third-party library pinned at v4.22 by cmake/FindTRNG4.cmake:46-48
whose source is NOT under /root/reference and cannot be installed offline, so what follows restates the PUBLISHED algorithm
(H. Bauke and S. Mertens, "Random numbers for large-scale distributed Monte Carlo simulations", Phys. Rev. E 75, 066701 (2007);
TRNG documentation of `yarn2`, `mrg2`, `uniform01_dist`):

    status (r0, r1), r in [0, m), m = 2^31 - 1;  default (0, 1);  seed(s): (int64(s) mod m, 1)
    step:   (r0, r1) <- ((a0 r0 + a1 r1) mod m, r0),  a = (1498809829, 1160990996)   [parameter set "LEcuyer1"]
    output: 0 if r0 == 0 else g^r0 mod m,  g = 123567893
    jump(s): s steps at once (companion-matrix power; the library composes jump2(i) = 2^i steps over the set bits of s)
    uniform01: output * (1 / (max - min + 1)) = output / 2147483647 in [0, 1), one engine call per uniform

PARITY UNPINNED: the reference holds no golden vector of the stream and the library cannot be executed here.  What IS checked
(tests/test_yarn2_cpu.py): the algebra (jump(n) == n steps, jump additivity, output map == pow(g, r, m), g a primitive root so
the output map is a bijection, the recurrence has full period m^2 - 1 for these multipliers), and agreement of the three
restatements written separately: this file (python integers), csrc/yarn2.cuh (device, tables) and baseline/shim_yarn2 (the
library's class shape, compiled into the reference's own CUDA drivers).
"""
from __future__ import annotations

import numpy as np

M = 2147483647
GEN = 123567893
A0, A1 = 1498809829, 1160990996
MASK64 = (1 << 64) - 1


def _matmul(x, y):
    return ((x[0] * y[0] + x[1] * y[2]) % M, (x[0] * y[1] + x[1] * y[3]) % M,
            (x[2] * y[0] + x[3] * y[2]) % M, (x[2] * y[1] + x[3] * y[3]) % M)


def _matpow(s: int):
    acc, b = (1, 0, 0, 1), (A0, A1, 1, 0)
    while s:
        if s & 1:
            acc = _matmul(b, acc)
        b = _matmul(b, b)
        s >>= 1
    return acc


class Yarn2:
    """One engine.  Mirrors the calls the reference makes: seed, jump, and draws through uniform01()."""

    def __init__(self, seed: int | None = None):
        self.r0, self.r1 = 0, 1
        if seed is not None:
            self.seed(seed)

    def seed(self, s: int):
        s = int(s) & MASK64                      # `unsigned long` argument, taken over into a signed 64-bit integer by the library
        if s >= 1 << 63:
            s -= 1 << 64
        self.r0, self.r1 = s % M, 1              # t %= m; if (t < 0) t += m

    def step(self):
        self.r0, self.r1 = (A0 * self.r0 + A1 * self.r1) % M, self.r0

    def jump(self, s: int):
        s = int(s) & MASK64                      # the argument is an unsigned long long in the reference
        m = _matpow(s)
        self.r0, self.r1 = (m[0] * self.r0 + m[1] * self.r1) % M, (m[2] * self.r0 + m[3] * self.r1) % M

    def next_int(self) -> int:
        self.step()
        return 0 if self.r0 == 0 else pow(GEN, self.r0, M)

    def uniform01(self) -> float:
        return float(self.next_int()) * (1.0 / 2147483647.0)


def chain_uniforms(seed: int, seed_distance: int, n_chains: int, steps: int, chain_offset: int = 0, skip: int = 0) -> np.ndarray:
    """u[t][k]: the t-th draw (after `skip` earlier ones) of the engine of global chain chain_offset + k, as TRNGWrapper builds
    them (gpu/include/trng4cuda.cuh:47-51).  The recurrence is vectorised over chains in uint64 (a0 r0 + a1 r1 < 2^63)."""
    r0 = np.empty(n_chains, dtype=np.uint64)
    r1 = np.empty(n_chains, dtype=np.uint64)
    for k in range(n_chains):
        e = Yarn2(seed)
        e.jump((2 * int(seed_distance) * (chain_offset + k)) & MASK64)
        e.jump(skip)
        r0[k], r1[k] = e.r0, e.r1
    # output map by two tables like the library: g^r = g^(hi 2^16) g^lo
    t0 = np.empty(1 << 16, dtype=np.uint64)
    t1 = np.empty(1 << 15, dtype=np.uint64)
    x = 1
    for i in range(1 << 16):
        t0[i] = x
        x = x * GEN % M
    g16 = x                                       # g^(2^16)
    x = 1
    for i in range(1 << 15):
        t1[i] = x
        x = x * g16 % M
    u = np.empty((steps, n_chains), dtype=np.float64)
    a0, a1, m = np.uint64(A0), np.uint64(A1), np.uint64(M)
    for t in range(steps):
        r0, r1 = (a0 * r0 + a1 * r1) % m, r0
        out = (t1[r0 >> np.uint64(16)] * t0[r0 & np.uint64(0xFFFF)]) % m
        out[r0 == 0] = 0
        u[t] = out.astype(np.float64) * (1.0 / 2147483647.0)
    return u
