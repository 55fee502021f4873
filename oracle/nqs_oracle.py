"""CPU oracle for the NQS variational-Monte-Carlo hot path (TEST INFRASTRUCTURE ONLY).

This file is a plain numpy fp64 restatement of the reference algorithm
(dkkim1005/Neural_Network_Quantum_State).  It is *the checker*, never the
product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  The product path is the CUDA library
behind include/nqs_b200.h and fails loudly when that library is missing.

Parity pin: every function here was checked against the reference's own CPU
implementation (cpu/include/*.hpp compiled into oracle/_ref/libnqs_ref.so by
oracle/Makefile, plus a CPU long-range Hamiltonian shim that restates
gpu/include/impl_hamiltonians.cuh:118-259) on shared inputs; the resulting
vectors are committed under tests/golden/ together with the generating script
(tests/golden/make_golden.py).  The reference ships no tests / golden vectors
of its own.

All citations are relative to /root/reference/.  "GPU semantics" means the
behaviour of gpu/include/*.cuh, which is the semantic reference of the north
star; the CPU tree differs in a few documented places.

Index notation (as the reference): i = visible site (N), j = hidden unit (M),
k = Markov chain (K).  W is [N][M] row-major, spins are [K][N] (+1/-1).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

LN2 = 0.6931471805599453  # cpu/include/common.hpp:73 ; gpu: kln2d = std::log(2.0)

# ----------------------------------------------------------------------------------------------
# elementary functions
# ----------------------------------------------------------------------------------------------


def logcosh(z: np.ndarray) -> np.ndarray:
    """log(cosh(z)) for complex z, exactly the reference's overflow-safe formula.

    gpu/include/impl_neural_quantum_state.cuh:1238-1245, cpu/include/common.hpp:67-74:
      e = exp(-2|x|);  log( (1+e) cos y + i (1-e) sin y sgn(x) ) + |x| - ln2
    (copysign(1, x): sgn(+0) = +1, sgn(-0) = -1.)
    """
    z = np.asarray(z, dtype=np.complex128)
    x = z.real
    y = z.imag
    absx = np.abs(x)
    e = np.exp(-2.0 * absx)
    re = (1.0 + e) * np.cos(y)
    im = (1.0 - e) * np.sin(y) * np.copysign(1.0, x)
    return np.log(re + 1j * im) + (absx - LN2)


def checkerboard_order(n_sites: int) -> np.ndarray:
    """Site visiting order of one LITFIChain sweep (N proposals).

    gpu/include/impl_hamiltonians.cuh:163-180 builds a circular list
    0 -> 2 -> 4 -> ... (even) -> 1 -> 3 -> ... (odd) -> 0 and `sampling_` advances
    the pointer BEFORE using it (:209-210), starting at 0.  One sweep therefore
    visits 2,4,...,1,3,...,0 and the pointer is back at 0 afterwards.
    """
    evens = list(range(0, n_sites, 2))
    odds = list(range(1, n_sites, 2))
    ring = evens + odds  # ring[0] == 0
    return np.array(ring[1:] + ring[:1], dtype=np.int32)


def sequential_order(n_sites: int) -> np.ndarray:
    """Site order of Sampler4SpinHalf (pynqs): 1,2,...,N-1,0.

    gpu/include/impl_meas.cuh:12-21 (list) and :33-34 (advance before use).
    """
    return np.array(list(range(1, n_sites)) + [0], dtype=np.int32)


def lr_coupling_matrix(L: int, J: float, alpha: float, pbc: bool) -> np.ndarray:
    """J_ij = J * d(i,j)^-alpha, zero diagonal.  gpu/include/impl_hamiltonians.cuh:136-161.

    OBC: d = |i-j|.  PBC (L even only): d = (j-i) if (j-i) < L/2 else L-(j-i)   (:146).
    """
    if pbc and L % 2 == 1:
        raise ValueError('kL%2 == 1 (set "isPBC" to "false".)')
    Jm = np.zeros((L, L), dtype=np.float64)
    for i in range(L):
        for j in range(i + 1, L):
            d = float(j - i)
            if pbc and not ((j - i) < L // 2):
                d = float(L - (j - i))
            Jm[i, j] = J * d ** (-alpha)
            Jm[j, i] = Jm[i, j]
    return Jm


# ----------------------------------------------------------------------------------------------
# uniforms: either pre-drawn [steps][K] arrays or the engine's counter-based Philox stream
# ----------------------------------------------------------------------------------------------

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox-4x32-10 (Salmon et al., SC'11).  counter: [...,4] uint32, key: [...,2] uint32.

    Not part of the reference (which uses TRNG4 yarn2, absent offline).  This is
    the engine's own counter RNG, restated here so that accept/reject parity can also be
    checked in "internal RNG" mode.  Same round function as cuRAND/Random123.
    """
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = _PHILOX_M0 * c[0]
        p1 = _PHILOX_M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c = [(hi1 ^ c[1] ^ k0) & _MASK32, lo1, (hi0 ^ c[3] ^ k1) & _MASK32, lo0]
        k0 = (k0 + np.uint64(_PHILOX_W0)) & _MASK32
        k1 = (k1 + np.uint64(_PHILOX_W1)) & _MASK32
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_uniform(seed: int, chain_ids: np.ndarray, step: int) -> np.ndarray:
    """U[0,1) fp64 with 53 random bits for (seed, global chain id, proposal index).

    counter = (chain_lo, chain_hi, step_lo, step_hi), key = (seed_lo, seed_hi);
    u = ((w0 >> 5) * 2^26 + (w1 >> 6)) * 2^-53.  Mirrors csrc/kernels/philox.cuh.
    """
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    ctr = np.empty(chain_ids.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = (chain_ids & _MASK32).astype(np.uint32)
    ctr[..., 1] = (chain_ids >> np.uint64(32)).astype(np.uint32)
    ctr[..., 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32((step >> 32) & 0xFFFFFFFF)
    key = np.empty(chain_ids.shape + (2,), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    w = philox4x32_10(ctr, key)
    hi = (w[..., 0] >> np.uint32(5)).astype(np.float64)
    lo = (w[..., 1] >> np.uint32(6)).astype(np.float64)
    return (hi * 67108864.0 + lo) * (1.0 / 9007199254740992.0)


class UniformSource:
    """Feeds one U[0,1) per chain per proposal, like TRNGWrapper::get_uniformDist
    (gpu/include/trng4cuda.cuh:62-65).  `step` counts proposals since construction."""

    def __init__(self, n_chains: int, predrawn: Optional[np.ndarray] = None, seed: int = 0,
                 chain_offset: int = 0):
        self.n_chains = n_chains
        self.predrawn = None if predrawn is None else np.asarray(predrawn, dtype=np.float64)
        self.seed = seed
        self.chain_ids = np.arange(chain_offset, chain_offset + n_chains, dtype=np.uint64)
        self.step = 0

    def next(self) -> np.ndarray:
        if self.predrawn is not None:
            u = self.predrawn[self.step]
        else:
            u = philox_uniform(self.seed, self.chain_ids, self.step)
        self.step += 1
        return u


# ----------------------------------------------------------------------------------------------
# parameter text files: "(re,im)" tokens, reference save/load format
# ----------------------------------------------------------------------------------------------


def _fmt_complex(z: complex, prec: int) -> str:
    # std::ostream << std::complex with setprecision(prec), default floatfield == printf %.{prec}g
    return "(%.*g,%.*g)" % (prec, z.real, prec, z.imag)


def _write_rows(path: str, rows, prec: int, trailing_newline: bool):
    with open(path, "w") as f:
        for r, row in enumerate(rows):
            f.write("".join(_fmt_complex(complex(z), prec) + " " for z in row))
            if trailing_newline or r < len(rows) - 1:
                f.write("\n")


def _read_complex_tokens(path: str) -> Optional[np.ndarray]:
    if not os.path.exists(path):
        return None
    out = []
    with open(path) as f:
        for tok in f.read().split():
            tok = tok.strip()
            if tok.startswith("("):
                body = tok.strip("()")
                parts = body.split(",")
                out.append(complex(float(parts[0]), float(parts[1]) if len(parts) > 1 else 0.0))
            else:
                out.append(complex(float(tok), 0.0))
    return np.array(out, dtype=np.complex128)


# ----------------------------------------------------------------------------------------------
# ansatz ("machine")
# ----------------------------------------------------------------------------------------------


class RBM:
    """Complex RBM, GPU semantics.  gpu/include/impl_neural_quantum_state.cuh:8-299,
    kernels :1264-1465; CPU twin cpu/include/impl_neural_quantum_state.hpp:32-367.

    variables = [W (N*M, index i*M+j) | a (N) | b (M)];  lnpsi_k = sum_j logcosh(y_kj) + sum_i a_i s_ki.
    """

    kind = "rbm"

    def __init__(self, n_inputs: int, n_hiddens: int, n_chains: int, rng: Optional[np.random.Generator] = None):
        self.N, self.M, self.K = n_inputs, n_hiddens, n_chains
        self.P = n_inputs * n_hiddens + n_inputs + n_hiddens
        self.variables = np.zeros(self.P, dtype=np.complex128)
        self.spins = np.ones((n_chains, n_inputs), dtype=np.float64)
        self.y = np.zeros((n_chains, n_hiddens), dtype=np.complex128)
        self.sa = np.zeros(n_chains, dtype=np.complex128)
        self.index_ = 0  # :19
        if rng is not None:
            self.random_init(rng)

    # views
    @property
    def W(self):
        return self.variables[: self.N * self.M].reshape(self.N, self.M)

    @property
    def a(self):
        return self.variables[self.N * self.M: self.N * self.M + self.N]

    @property
    def b(self):
        return self.variables[self.N * self.M + self.N:]

    def random_init(self, rng: np.random.Generator):
        """Init law of the ctor (:30-48): W = 0.1*(g+ig'), g~N(0,1/(N+M)); a = 0; b = 0.1*(g+ig'), g~N(0,1/M).
        (The reference seeds from the clock; synthetic benches use a fixed numpy seed.)"""
        N, M = self.N, self.M
        sw, sb = math.sqrt(1.0 / (N + M)), math.sqrt(1.0 / M)
        self.W[...] = 0.1 * (rng.normal(0, sw, (N, M)) + 1j * rng.normal(0, sw, (N, M)))
        self.a[...] = 0.0
        self.b[...] = 0.1 * (rng.normal(0, sb, M) + 1j * rng.normal(0, sb, M))

    def _theta(self, spins):
        # y_kj = sum_i s_ki W_ij + b_j   (:73-79: fill, Zgeru bias, Zgemm)
        return spins @ self.W + self.b[None, :]

    def initialize(self, spins: np.ndarray) -> np.ndarray:
        """:67-91.  Returns lnpsi[K]."""
        self.spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(self.spins)
        self.sa = self.spins @ self.a
        return logcosh(self.y).sum(axis=1) + self.sa

    def forward_flip(self, idx: int) -> np.ndarray:
        """forward(int) :93-104 (k3 :1264-1278, k4 :1391-1404): lnpsi' for flipping site idx on every chain."""
        self.index_ = idx
        s = self.spins[:, idx]
        ly = logcosh(self.y - self.W[idx][None, :] * (2.0 * s)[:, None])
        return (self.sa - (2.0 * s) * self.a[idx]) + ly.sum(axis=1)

    def forward_spins(self, spins: np.ndarray, save: bool = True) -> np.ndarray:
        """forward(spins, lnpsi, save) :107-129.  NOTE the reference computes sa from the MEMBER
        spins (:119-120), not from the argument -- kept."""
        spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(spins)
        self.sa = self.spins @ self.a
        out = logcosh(self.y).sum(axis=1) + self.sa
        if save:
            self.spins = spins.copy()
        return out

    def spin_flip(self, mask: np.ndarray, idx: int = -1):
        """:172-182 (k7 :1314-1329, k8 :1451-1465, k9 :1353-1368)."""
        if idx != -1:
            self.index_ = idx
        i = self.index_
        two_delta = np.where(mask, 2.0, 0.0)
        s = self.spins[:, i]
        self.y = self.y - self.W[i][None, :] * (two_delta * s)[:, None]
        self.sa = self.sa - (two_delta * s) * self.a[i]
        self.spins[:, i] = (1.0 - two_delta) * s

    def backward(self) -> np.ndarray:
        """:146-154, k13 :1426-1449.  O[K][P] = [s_ki tanh(y_kj) (i*M+j) | s_ki | tanh(y_kj)]."""
        t = np.tanh(self.y)
        O = np.empty((self.K, self.P), dtype=np.complex128)
        NM = self.N * self.M
        O[:, :NM] = (self.spins[:, :, None] * t[:, None, :]).reshape(self.K, NM)
        O[:, NM:NM + self.N] = self.spins
        O[:, NM + self.N:] = t
        return O

    def update_variables(self, dx: np.ndarray, lr: float):
        """:156-170: variables -= lr*dx, then y and sa re-derived for the CURRENT spins."""
        self.variables = self.variables - lr * np.asarray(dx, dtype=np.complex128)
        self.y = self._theta(self.spins)
        self.sa = self.spins @ self.a

    # text files (:196-286): <prefix>Dw.dat (N lines x M tokens), Da.dat, Db.dat
    def save(self, prefix: str, prec: int = 10):
        _write_rows(prefix + "Dw.dat", [self.W[i] for i in range(self.N)], prec, True)
        _write_rows(prefix + "Da.dat", [self.a], prec, True)
        _write_rows(prefix + "Db.dat", [self.b], prec, False)

    def load(self, prefix: str):
        for suffix, view, name in (("Dw.dat", self.W, "w"), ("Da.dat", self.a, "a"), ("Db.dat", self.b, "b")):
            raw = _read_complex_tokens(prefix + suffix)
            if raw is None:
                print("# --- file-path: %s is not exist..." % (prefix + suffix))
                continue
            if raw.size == view.size:
                view[...] = raw.reshape(view.shape)
            else:
                print("# check '%s' size... " % name)


class FFNN:
    """One-hidden-layer complex FNN, GPU semantics.  gpu/include/impl_neural_quantum_state.cuh:747-899,
    kernels :1621-1690.  lnpsi_k = sum_j w1o_j logcosh(y_kj); variables = [W1 (N*M, i*M+j) | b1 (M) | w1o (M)].

    GPU emits the W-block of O transposed (index j*N+i, :1644-1663) and applies dx with the same
    transposition (:1665-1690); the CPU tree uses the natural layout.  `transposed_grad`
    selects; default True (GPU)."""

    kind = "ffnn"

    def __init__(self, n_inputs: int, n_hiddens: int, n_chains: int, rng: Optional[np.random.Generator] = None,
                 transposed_grad: bool = True):
        self.N, self.M, self.K = n_inputs, n_hiddens, n_chains
        self.P = n_inputs * n_hiddens + 2 * n_hiddens
        self.variables = np.zeros(self.P, dtype=np.complex128)
        self.spins = np.ones((n_chains, n_inputs), dtype=np.float64)
        self.y = np.zeros((n_chains, n_hiddens), dtype=np.complex128)
        self.index_ = 0
        self.transposed_grad = transposed_grad
        if rng is not None:
            self.random_init(rng)

    @property
    def W(self):
        return self.variables[: self.N * self.M].reshape(self.N, self.M)

    @property
    def b(self):
        return self.variables[self.N * self.M: self.N * self.M + self.M]

    @property
    def w1o(self):
        return self.variables[self.N * self.M + self.M:]

    def random_init(self, rng: np.random.Generator):
        """:766-783: W1 = g + 0.1 i g', g~N(0,1/(N+M)); b1 = 0; w1o = g + 0.1 i g', g~N(0,1/M)."""
        N, M = self.N, self.M
        sw, so = math.sqrt(1.0 / (N + M)), math.sqrt(1.0 / M)
        self.W[...] = rng.normal(0, sw, (N, M)) + 0.1j * rng.normal(0, sw, (N, M))
        self.b[...] = 0.0
        self.w1o[...] = rng.normal(0, so, M) + 0.1j * rng.normal(0, so, M)

    def _theta(self, spins):
        return spins @ self.W + self.b[None, :]

    def initialize(self, spins: np.ndarray) -> np.ndarray:
        """:799-819."""
        self.spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(self.spins)
        return logcosh(self.y) @ self.w1o

    def forward_flip(self, idx: int) -> np.ndarray:
        """:821-829."""
        self.index_ = idx
        s = self.spins[:, idx]
        return logcosh(self.y - self.W[idx][None, :] * (2.0 * s)[:, None]) @ self.w1o

    def forward_spins(self, spins: np.ndarray, save: bool = True) -> np.ndarray:
        """:831-847."""
        spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(spins)
        out = logcosh(self.y) @ self.w1o
        if save:
            self.spins = spins.copy()
        return out

    def spin_flip(self, mask: np.ndarray, idx: int = -1):
        """:880-888."""
        if idx != -1:
            self.index_ = idx
        i = self.index_
        two_delta = np.where(mask, 2.0, 0.0)
        s = self.spins[:, i]
        self.y = self.y - self.W[i][None, :] * (two_delta * s)[:, None]
        self.spins[:, i] = (1.0 - two_delta) * s

    def backward(self) -> np.ndarray:
        """:858-865, k15 :1622-1663.  dW_ij = tanh(y_j) s_i w1o_j; db1_j = tanh(y_j) w1o_j; dw1o_j = logcosh(y_j)."""
        t = np.tanh(self.y) * self.w1o[None, :]
        O = np.empty((self.K, self.P), dtype=np.complex128)
        NM = self.N * self.M
        if self.transposed_grad:
            O[:, :NM] = (t[:, :, None] * self.spins[:, None, :]).reshape(self.K, NM)  # index j*N+i
        else:
            O[:, :NM] = (self.spins[:, :, None] * t[:, None, :]).reshape(self.K, NM)  # index i*M+j
        O[:, NM:NM + self.M] = t
        O[:, NM + self.M:] = logcosh(self.y)
        return O

    def update_variables(self, dx: np.ndarray, lr: float):
        """:867-878 + k16 :1665-1690 (un-transposes dx's W block when transposed_grad)."""
        dx = np.asarray(dx, dtype=np.complex128)
        NM = self.N * self.M
        dW = dx[:NM].reshape(self.M, self.N).T if self.transposed_grad else dx[:NM].reshape(self.N, self.M)
        new = self.variables.copy()
        new[:NM] = (self.W - lr * dW).reshape(NM)
        new[NM:] = self.variables[NM:] - lr * dx[NM:]
        self.variables = new
        self.y = self._theta(self.spins)

    def save(self, prefix: str, prec: int = 10):
        """:901-937: Dw1.dat, Dw2.dat (w1o, newline-terminated), Db1.dat."""
        _write_rows(prefix + "Dw1.dat", [self.W[i] for i in range(self.N)], prec, True)
        _write_rows(prefix + "Dw2.dat", [self.w1o], prec, True)
        _write_rows(prefix + "Db1.dat", [self.b], prec, False)

    def load(self, prefix: str):
        for suffix, view, name in (("Dw1.dat", self.W, "w1"), ("Dw2.dat", self.w1o, "w2"), ("Db1.dat", self.b, "b1")):
            raw = _read_complex_tokens(prefix + suffix)
            if raw is None:
                print("# --- file-path: %s is not exist..." % (prefix + suffix))
                continue
            if raw.size == view.size:
                view[...] = raw.reshape(view.shape)
            else:
                print("# check '%s' size... " % name)


class RBMTrSymm:
    """Translation-symmetric complex RBM, GPU semantics.  gpu/include/impl_neural_quantum_state.cuh:301-538, kernels
    :1487-1553 (the ansatz of gpu/src/LICH-train_rbmtrsymm.cu).

    variables = [w (alpha*N, index f*N+i) | a (1) | b (alpha)], P = N*alpha + 1 + alpha.  symmetrize_variables_ (:534-538,
    kernel :1523-1553) expands them to wf[i][f*N+j] = w[f][(i+j)%N], bf[f*N+j] = b[f], af[i] = a[0]; every sampler operation
    (initialize / forward / spin_flip, :364-472) is the plain RBM's on (wf, af, bf) with M = alpha*N hidden units.
    backward (:443-450, kernel :1487-1521):  d_w[f*N+i] = sum_j tanh(y[f*N+j]) s[(N+i-j)%N],  d_a = sum_i s_i,
    d_b[f] = sum_j tanh(y[f*N+j]).  NOTE forward(spins, lnpsi, save) takes sa from its ARGUMENT here (:401-403), unlike RBM.
    """

    kind = "rbmtrsymm"

    def __init__(self, n_inputs: int, alpha: int, n_chains: int, rng: Optional[np.random.Generator] = None):
        self.N, self.alpha, self.K = n_inputs, alpha, n_chains
        self.M = alpha * n_inputs
        self.P = n_inputs * alpha + 1 + alpha
        self.variables = np.zeros(self.P, dtype=np.complex128)
        self.spins = np.ones((n_chains, n_inputs), dtype=np.float64)
        self.y = np.zeros((n_chains, self.M), dtype=np.complex128)
        self.sa = np.zeros(n_chains, dtype=np.complex128)
        self.index_ = 0
        if rng is not None:
            self.random_init(rng)

    def random_init(self, rng: np.random.Generator):
        """ctor :325-345: w = 0.1 (g + i g'), g ~ N(0, 1/((1+alpha) N)); a = 0; b = 0.1 (g + i g'), g ~ N(0, 1/(N alpha))."""
        N, al = self.N, self.alpha
        sw, sb = math.sqrt(1.0 / ((1 + al) * N)), math.sqrt(1.0 / (N * al))
        self.variables[: N * al] = 0.1 * (rng.normal(0, sw, N * al) + 1j * rng.normal(0, sw, N * al))
        self.variables[N * al] = 0.0
        self.variables[N * al + 1:] = 0.1 * (rng.normal(0, sb, al) + 1j * rng.normal(0, sb, al))

    # expanded network (recomputed from the variables on every use: the variables are the state)
    @property
    def W(self):
        N, al = self.N, self.alpha
        w = self.variables[: N * al].reshape(al, N)
        i = np.arange(N)[:, None, None]
        f = np.arange(al)[None, :, None]
        j = np.arange(N)[None, None, :]
        return w[f, (i + j) % N].reshape(N, al * N)          # wf[i][f*N+j]

    @property
    def a(self):
        return np.full(self.N, self.variables[self.N * self.alpha])

    @property
    def b(self):
        return np.repeat(self.variables[self.N * self.alpha + 1:], self.N)

    def _theta(self, spins):
        return spins @ self.W + self.b[None, :]

    def initialize(self, spins: np.ndarray) -> np.ndarray:
        self.spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(self.spins)
        self.sa = self.spins @ self.a
        return logcosh(self.y).sum(axis=1) + self.sa

    def forward_flip(self, idx: int) -> np.ndarray:
        self.index_ = idx
        s = self.spins[:, idx]
        ly = logcosh(self.y - self.W[idx][None, :] * (2.0 * s)[:, None])
        return (self.sa - (2.0 * s) * self.a[idx]) + ly.sum(axis=1)

    def forward_spins(self, spins: np.ndarray, save: bool = True) -> np.ndarray:
        spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(spins)
        self.sa = spins @ self.a                                 # the ARGUMENT's spins (:401-403)
        out = logcosh(self.y).sum(axis=1) + self.sa
        if save:
            self.spins = spins.copy()
        return out

    def spin_flip(self, mask: np.ndarray, idx: int = -1):
        if idx != -1:
            self.index_ = idx
        i = self.index_
        two_delta = np.where(mask, 2.0, 0.0)
        s = self.spins[:, i]
        self.y = self.y - self.W[i][None, :] * (two_delta * s)[:, None]
        self.sa = self.sa - (two_delta * s) * self.a[i]
        self.spins[:, i] = (1.0 - two_delta) * s

    def backward(self) -> np.ndarray:
        N, al = self.N, self.alpha
        t = np.tanh(self.y).reshape(self.K, al, N)              # [k][f][j]
        O = np.empty((self.K, self.P), dtype=np.complex128)
        i = np.arange(N)[:, None]
        j = np.arange(N)[None, :]
        sh = self.spins[:, (N + i - j) % N]                      # [k][i][j] = s[(N+i-j)%N]
        O[:, : N * al] = np.einsum("kfj,kij->kfi", t, sh).reshape(self.K, N * al)
        O[:, N * al] = self.spins.sum(axis=1)
        O[:, N * al + 1:] = t.sum(axis=2)
        return O

    def update_variables(self, dx: np.ndarray, lr: float):
        """:452-465: variables -= lr*dx, symmetrize, y and sa re-derived for the current spins."""
        self.variables = self.variables - lr * np.asarray(dx, dtype=np.complex128)
        self.y = self._theta(self.spins)
        self.sa = self.spins @ self.a

    def save(self, path: str, prec: int = 10):
        """:474-482: every variable, blank separated, in one file."""
        _write_rows(path, [self.variables], prec, False)

    def load(self, path: str):
        raw = _read_complex_tokens(path)
        if raw is None:
            print("# --- file-path: %s is not exist..." % path)
        elif raw.size == self.variables.size:
            self.variables[...] = raw
        else:
            print(" check parameter size... ")


class RBMZ2PrSymm:
    """Z2- and parity-symmetric complex RBM, GPU semantics.  gpu/include/impl_neural_quantum_state.cuh:540-745, kernels
    :1556-1618 (the ansatz of gpu/src/LICH-train_rbmz2prsymm.cu).

    variables = [w (N*alpha, index i*alpha+f) | b (alpha)], P = N*alpha + alpha, NO visible bias.  symmetrize_variables_
    (:740-745, kernel :1588-1618) expands them to 4 hidden units per filter: wf[i][4f+0] = w[i][f], wf[i][4f+1] = -w[i][f],
    wf[i][4f+2] = w[N-1-i][f], wf[i][4f+3] = -w[N-1-i][f], bf[4f+j] = b[f]; every sampler operation (:589-637, 670-678) is the
    plain RBM's on (wf, 0, bf) with M = 4 alpha.  backward (:639-646, kernel :1556-1585):
    d_w[i*alpha+f] = (tanh y[4f] - tanh y[4f+1]) s_i + (tanh y[4f+2] - tanh y[4f+3]) s_{N-1-i},  d_b[f] = sum of the 4 tanh."""

    kind = "rbmz2prsymm"

    def __init__(self, n_inputs: int, alpha: int, n_chains: int, rng: Optional[np.random.Generator] = None):
        self.N, self.alpha, self.K = n_inputs, alpha, n_chains
        self.M = 4 * alpha
        self.P = n_inputs * alpha + alpha
        self.variables = np.zeros(self.P, dtype=np.complex128)
        self.spins = np.ones((n_chains, n_inputs), dtype=np.float64)
        self.y = np.zeros((n_chains, self.M), dtype=np.complex128)
        self.sa = np.zeros(n_chains, dtype=np.complex128)
        self.index_ = 0
        if rng is not None:
            self.random_init(rng)

    def random_init(self, rng: np.random.Generator):
        """ctor :562-579: w = 0.1 (g + i g'), g ~ N(0, 1/(4 alpha + N)); b = 0.1 (g + i g'), g ~ N(0, 1/(4 alpha))."""
        N, al = self.N, self.alpha
        sw, sb = math.sqrt(1.0 / (4 * al + N)), math.sqrt(1.0 / (4 * al))
        self.variables[: N * al] = 0.1 * (rng.normal(0, sw, N * al) + 1j * rng.normal(0, sw, N * al))
        self.variables[N * al:] = 0.1 * (rng.normal(0, sb, al) + 1j * rng.normal(0, sb, al))

    @property
    def W(self):
        N, al = self.N, self.alpha
        w = self.variables[: N * al].reshape(N, al)
        wr = w[::-1]                                            # w[N-1-i][f]
        return np.stack([w, -w, wr, -wr], axis=2).reshape(N, 4 * al)   # wf[i][4f+j]

    @property
    def a(self):
        return np.zeros(self.N, dtype=np.complex128)

    @property
    def b(self):
        return np.repeat(self.variables[self.N * self.alpha:], 4)

    def _theta(self, spins):
        return spins @ self.W + self.b[None, :]

    def initialize(self, spins: np.ndarray) -> np.ndarray:
        self.spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(self.spins)
        return logcosh(self.y).sum(axis=1)

    def forward_flip(self, idx: int) -> np.ndarray:
        self.index_ = idx
        s = self.spins[:, idx]
        return logcosh(self.y - self.W[idx][None, :] * (2.0 * s)[:, None]).sum(axis=1)

    def forward_spins(self, spins: np.ndarray, save: bool = True) -> np.ndarray:
        spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(spins)
        out = logcosh(self.y).sum(axis=1)
        if save:
            self.spins = spins.copy()
        return out

    def spin_flip(self, mask: np.ndarray, idx: int = -1):
        if idx != -1:
            self.index_ = idx
        i = self.index_
        two_delta = np.where(mask, 2.0, 0.0)
        s = self.spins[:, i]
        self.y = self.y - self.W[i][None, :] * (two_delta * s)[:, None]
        self.spins[:, i] = (1.0 - two_delta) * s

    def backward(self) -> np.ndarray:
        N, al = self.N, self.alpha
        t = np.tanh(self.y).reshape(self.K, al, 4)              # [k][f][j]
        d0, d1 = t[:, :, 0] - t[:, :, 1], t[:, :, 2] - t[:, :, 3]
        O = np.empty((self.K, self.P), dtype=np.complex128)
        O[:, : N * al] = (d0[:, None, :] * self.spins[:, :, None] + d1[:, None, :] * self.spins[:, ::-1, None]).reshape(self.K, N * al)
        O[:, N * al:] = t.sum(axis=2)
        return O

    def update_variables(self, dx: np.ndarray, lr: float):
        """:648-661: variables -= lr*dx, symmetrize, y re-derived for the current spins."""
        self.variables = self.variables - lr * np.asarray(dx, dtype=np.complex128)
        self.y = self._theta(self.spins)

    def save(self, path: str, prec: int = 10):
        """:680-689: every variable, blank separated, in one file."""
        _write_rows(path, [self.variables], prec, False)

    def load(self, path: str):
        raw = _read_complex_tokens(path)
        if raw is None:
            print("# --- file-path: %s is not exist..." % path)
        elif raw.size == self.variables.size:
            self.variables[...] = raw
        else:
            print(" check parameter size... ")


class FFNNTrSymm:
    """Translation-symmetric one-hidden-layer complex FNN, GPU semantics.  gpu/include/impl_neural_quantum_state.cuh:1019-1223,
    kernels :1693-1750 (the ansatz of gpu/src/LICH-train_ffnntrsymm.cu).

    variables = [wi1 (alpha*N, index f*N+i) | b1 (alpha) | w1o (alpha)], P = N*alpha + 2 alpha.  symmetrize_variables_ (:1217-1223,
    kernel :1693-1717): W1[i][f*N+j] = wi1[f][(i+j)%N], b1f[f*N+j] = b1[f], w1of[f*N+j] = w1o[f]; sampler operations
    (:1081-1131, 1157-1165) are the plain FFNN's on the expansion with M = alpha*N.  backward (:1133-1141, kernel :1720-1750):
    d_wi1[f*N+i] = sum_j w1of[fN+j] tanh(y[fN+j]) s[(N+i-j)%N],  d_b1[f] = sum_j w1of tanh(y),  d_w1o[f] = sum_j logcosh(y[fN+j])."""

    kind = "ffnntrsymm"

    def __init__(self, n_inputs: int, alpha: int, n_chains: int, rng: Optional[np.random.Generator] = None):
        self.N, self.alpha, self.K = n_inputs, alpha, n_chains
        self.M = alpha * n_inputs
        self.P = n_inputs * alpha + 2 * alpha
        self.variables = np.zeros(self.P, dtype=np.complex128)
        self.spins = np.ones((n_chains, n_inputs), dtype=np.float64)
        self.y = np.zeros((n_chains, self.M), dtype=np.complex128)
        self.index_ = 0
        if rng is not None:
            self.random_init(rng)

    def random_init(self, rng: np.random.Generator):
        """ctor :1041-1061: wi1 = g + 0.1 i g', g ~ N(0, 1/((1+alpha) N)); b1 = 0; w1o = g + 0.1 i g', g ~ N(0, 1/(alpha N))."""
        N, al = self.N, self.alpha
        sw, so = math.sqrt(1.0 / ((1 + al) * N)), math.sqrt(1.0 / (al * N))
        self.variables[: N * al] = rng.normal(0, sw, N * al) + 0.1j * rng.normal(0, sw, N * al)
        self.variables[N * al: N * al + al] = 0.0
        self.variables[N * al + al:] = rng.normal(0, so, al) + 0.1j * rng.normal(0, so, al)

    @property
    def W(self):
        N, al = self.N, self.alpha
        w = self.variables[: N * al].reshape(al, N)
        i = np.arange(N)[:, None, None]
        f = np.arange(al)[None, :, None]
        j = np.arange(N)[None, None, :]
        return w[f, (i + j) % N].reshape(N, al * N)

    @property
    def b(self):
        return np.repeat(self.variables[self.N * self.alpha: self.N * self.alpha + self.alpha], self.N)

    @property
    def w1o(self):
        return np.repeat(self.variables[self.N * self.alpha + self.alpha:], self.N)

    def _theta(self, spins):
        return spins @ self.W + self.b[None, :]

    def initialize(self, spins: np.ndarray) -> np.ndarray:
        self.spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(self.spins)
        return logcosh(self.y) @ self.w1o

    def forward_flip(self, idx: int) -> np.ndarray:
        self.index_ = idx
        s = self.spins[:, idx]
        return logcosh(self.y - self.W[idx][None, :] * (2.0 * s)[:, None]) @ self.w1o

    def forward_spins(self, spins: np.ndarray, save: bool = True) -> np.ndarray:
        spins = np.array(spins, dtype=np.float64).reshape(self.K, self.N)
        self.y = self._theta(spins)
        out = logcosh(self.y) @ self.w1o
        if save:
            self.spins = spins.copy()
        return out

    def spin_flip(self, mask: np.ndarray, idx: int = -1):
        if idx != -1:
            self.index_ = idx
        i = self.index_
        two_delta = np.where(mask, 2.0, 0.0)
        s = self.spins[:, i]
        self.y = self.y - self.W[i][None, :] * (two_delta * s)[:, None]
        self.spins[:, i] = (1.0 - two_delta) * s

    def backward(self) -> np.ndarray:
        N, al = self.N, self.alpha
        t = (np.tanh(self.y) * self.w1o[None, :]).reshape(self.K, al, N)      # [k][f][j]
        O = np.empty((self.K, self.P), dtype=np.complex128)
        i = np.arange(N)[:, None]
        j = np.arange(N)[None, :]
        sh = self.spins[:, (N + i - j) % N]                                  # [k][i][j] = s[(N+i-j)%N]
        O[:, : N * al] = np.einsum("kfj,kij->kfi", t, sh).reshape(self.K, N * al)
        O[:, N * al: N * al + al] = t.sum(axis=2)
        O[:, N * al + al:] = logcosh(self.y).reshape(self.K, al, N).sum(axis=2)
        return O

    def update_variables(self, dx: np.ndarray, lr: float):
        """:1143-1155: variables -= lr*dx (no transposition: the tied block is f*N+i on both sides), symmetrize, y re-derived."""
        self.variables = self.variables - lr * np.asarray(dx, dtype=np.complex128)
        self.y = self._theta(self.spins)

    def save(self, path: str, prec: int = 10):
        """:1167-1176: every variable, blank separated, in one file."""
        _write_rows(path, [self.variables], prec, False)

    def load(self, path: str):
        raw = _read_complex_tokens(path)
        if raw is None:
            print("# --- file-path: %s is not exist..." % path)
        elif raw.size == self.variables.size:
            self.variables[...] = raw
        else:
            print(" check parameter size... ")


def make_ansatz(kind: str, N: int, M: int, K: int, rng=None):
    if kind == "rbm":
        return RBM(N, M, K, rng)
    if kind == "ffnn":
        return FFNN(N, M, K, rng)
    if kind == "rbmtrsymm":          # M = the expanded width alpha*N (as in nqs_config.n_hiddens)
        assert M % N == 0
        return RBMTrSymm(N, M // N, K, rng)
    if kind == "rbmz2prsymm":        # M = 4*alpha
        assert M % 4 == 0
        return RBMZ2PrSymm(N, M // 4, K, rng)
    if kind == "ffnntrsymm":         # M = alpha*N
        assert M % N == 0
        return FFNNTrSymm(N, M // N, K, rng)
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------
# sampler + Hamiltonian
# ----------------------------------------------------------------------------------------------


class LITFIChainSampler:
    """BaseParallelSampler + LITFIChain, GPU semantics.

    gpu/include/impl_mcmc_sampler.cuh:6-102 and gpu/include/impl_hamiltonians.cuh:118-259.
    `order` defaults to the checkerboard ring; pass sequential_order(N) and a random initial
    state for the pynqs Sampler4SpinHalf flavour (gpu/include/impl_meas.cuh:5-41).
    """

    def __init__(self, machine, h: float, J: float, alpha: float, pbc: bool, uniforms: UniformSource,
                 order: Optional[np.ndarray] = None):
        self.machine = machine
        self.N, self.K = machine.N, machine.K
        self.h, self.J = float(h), float(J)
        self.Jm = lr_coupling_matrix(self.N, J, alpha, pbc)
        self.order = checkerboard_order(self.N) if order is None else np.asarray(order, dtype=np.int32)
        self.pos = 0  # position inside `order` of the NEXT site to visit
        self.uniforms = uniforms
        self.lnpsi0 = np.zeros(self.K, dtype=np.complex128)
        self.lnpsi1 = np.zeros(self.K, dtype=np.complex128)
        self.accept_log = []  # per proposal: bool[K]  (only kept when record=True)
        self.ratio_log = []
        self.record = False

    def initial_spins(self) -> np.ndarray:
        """initialize_ :192-204: Neel (+,-,+,...) if J > 0 else all up."""
        s = np.ones((self.K, self.N), dtype=np.float64)
        if self.J > 0:
            s[:, 1::2] = -1.0
        return s

    def warm_up(self, n_sweeps: int = 100, spins: Optional[np.ndarray] = None):
        """impl_mcmc_sampler.cuh:18-25, including the quirk: after initialize_ the sampler
        calls accept_next_state_ with an all-true mask, i.e. machine.spin_flip(all, index_) which flips
        site index_ (0 after construction) on every chain while lnpsi0 keeps the un-flipped value."""
        self.lnpsi0 = self.machine.initialize(self.initial_spins() if spins is None else spins)
        self.machine.spin_flip(np.ones(self.K, dtype=bool))
        self.do_mcmc_steps(n_sweeps)

    def do_mcmc_steps(self, n_sweeps: int = 1):
        """impl_mcmc_sampler.cuh:28-39; accept kernel :75-102:
        ratio = exp(2*min(0, Re lnpsi1 - Re lnpsi0)); acc = u < ratio; lnpsi0 += acc*(lnpsi1-lnpsi0)."""
        for _ in range(n_sweeps * self.N):
            idx = int(self.order[self.pos])
            self.pos = (self.pos + 1) % len(self.order)
            self.lnpsi1 = self.machine.forward_flip(idx)
            u = self.uniforms.next()
            d = self.lnpsi1.real - self.lnpsi0.real
            ratio = np.exp(2.0 * np.where(d < 0, 1.0, 0.0) * d)
            acc = u < ratio
            self.lnpsi0 = self.lnpsi0 + np.where(acc, 1.0, 0.0) * (self.lnpsi1 - self.lnpsi0)
            if self.record:
                self.accept_log.append(acc.copy())
                self.ratio_log.append(ratio.copy())
            self.machine.spin_flip(acc)

    def get_htilda(self) -> np.ndarray:
        """get_htilda_ :220-241 (k10 :871-887, k11 :857-869):
        h_k = ( 1/2 sum_ij s_i J_ij s_j + h sum_i exp(lnpsi(s^(i)) - lnpsi0_k) ) / L   (per site!)."""
        s = self.machine.spins
        SJ = s @ self.Jm.T
        ht = (0.5 * (SJ * s).sum(axis=1)).astype(np.complex128)
        for i in range(self.N):
            self.lnpsi1 = self.machine.forward_flip(i)
            ht = ht + self.h * np.exp(self.lnpsi1 - self.lnpsi0)
        return ht * (1.0 / self.N)

    def get_lnpsiGradients(self) -> np.ndarray:
        return self.machine.backward()

    def evolve(self, dx: np.ndarray, lr: float):
        self.machine.update_variables(dx, lr)


# ----------------------------------------------------------------------------------------------
# stochastic reconfiguration with matrix-free PCG
# ----------------------------------------------------------------------------------------------


class SMatrix:
    """SMatrixForCG, gpu/include/functor_for_CG.cuh:91-195.  S = <O^H O> - <O>^H <O> + lambda*diag."""

    def __init__(self, O: np.ndarray, lam: float, reduce: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                 n_total: Optional[int] = None):
        # `reduce` sums an array over ranks (identity on one rank); n_total = global chain count.
        self.O = O
        self.lam = float(lam)
        self.reduce = reduce or (lambda x: x)
        self.Ktot = O.shape[0] if n_total is None else n_total
        self.n_dot = 0
        self.aO = self.reduce(O.sum(axis=0)) / self.Ktot  # :99
        # :141-160: diag_i = (1/K) sum_k |O_ki|^2 - |<O>_i|^2
        self.diag = self.reduce((O.real ** 2 + O.imag ** 2).sum(axis=0)) / self.Ktot - np.abs(self.aO) ** 2

    def dot(self, v: np.ndarray) -> np.ndarray:
        """:107-127 after un-doing the conj tricks:
        (S v)_i = (1/K) sum_k conj(O_ki) (sum_j O_kj v_j) - conj(<O>_i) sum_j <O>_j v_j + lambda diag_i v_i."""
        self.n_dot += 1
        z = self.O @ v
        b = self.reduce(self.O.conj().T @ z) / self.Ktot - self.aO.conj() * (self.aO @ v)
        return b + self.lam * self.diag * v

    def precond(self, r: np.ndarray) -> np.ndarray:
        """:179-195: x = r / ((1+lambda) diag)."""
        return r / ((1.0 + self.lam) * self.diag)


def pcg_solve(S: SMatrix, rhs: np.ndarray, x: np.ndarray, tol: float = 1e-5, max_iter: int = 1000,
              fixed_iters: Optional[int] = None):
    """ConjugateGradient::solve, gpu/include/conjugate_gradient.cuh:29-74 (Eigen-style PCG, warm start in x).
    Returns (x, n_iterations).  hermition_inner_product(a,b) = sum a_i conj(b_i)  (thrust_util.cuh:86-92).
    `fixed_iters` (not in the reference) runs exactly that many iterations ignoring the tolerance, for
    iteration-count-independent parity checks."""
    x = np.array(x, dtype=np.complex128)
    r = rhs - S.dot(x)
    rhs_norm2 = float((np.abs(rhs) ** 2).sum())
    if rhs_norm2 == 0.0:
        return np.zeros_like(x), 0
    thr = max(tol * tol * rhs_norm2, np.finfo(np.float64).tiny)
    res2 = float((np.abs(r) ** 2).sum())
    if fixed_iters is None and res2 < thr:
        return x, 0
    p = S.precond(r)
    abs_new = float((p * r.conj()).sum().real)
    it = 0
    n_max = max_iter if fixed_iters is None else fixed_iters
    while it < n_max:
        t = S.dot(p)
        alpha = abs_new / float((t * p.conj()).sum().real)
        x = x + alpha * p
        r = r - alpha * t
        res2 = float((np.abs(r) ** 2).sum())
        it += 1
        if fixed_iters is None and res2 < thr:
            break
        z = S.precond(r)
        abs_old = abs_new
        abs_new = float((z * r.conj()).sum().real)
        beta = abs_new / abs_old
        p = z + beta * p
    return x, it


@dataclass
class SRStats:
    iteration: int = 0
    e_mean: complex = 0j       # conj(conjHavg): <h> per site
    rsd: float = 0.0
    lam: float = 0.0
    cg_iters: int = 0
    finite: bool = True
    F: Optional[np.ndarray] = None
    dx: Optional[np.ndarray] = None


class StochasticReconfigurationCG:
    """gpu/include/optimizer.cuh:112-181, gpu/include/impl_optimizer.cuh:45-96.
    GPU settings: tol = 1e-5, maxIter = 1000, dx warm-started across iterations (zero at construction),
    lambda_p = max(100 * 0.9^p, 1e-2), p = 1, 2, ...   (:72-78)."""

    lambda0, kb, lamb_min = 100.0, 0.9, 1e-2

    def __init__(self, n_chains: int, n_variables: int, tol: float = 1e-5, max_iter: int = 1000,
                 reduce: Optional[Callable[[np.ndarray], np.ndarray]] = None, n_total: Optional[int] = None):
        self.K, self.P = n_chains, n_variables
        self.Ktot = n_chains if n_total is None else n_total
        self.dx = np.zeros(n_variables, dtype=np.complex128)
        self.bp = 1.0
        self.tol, self.max_iter = tol, max_iter
        self.reduce = reduce or (lambda x: x)

    def schedule(self) -> float:
        self.bp *= self.kb
        lam = self.lambda0 * self.bp
        return lam if lam > self.lamb_min else self.lamb_min

    def gradient(self, ht: np.ndarray, O: np.ndarray):
        """optimizer.cuh:131-146 + SR__FStep2__ (impl_optimizer.cuh:82-96):
        F_i = (1/K) sum_k conj(O_ki) h_k - conj(<O>_i) <h>."""
        hsum = self.reduce(np.array([ht.sum()]))[0]
        havg = hsum / self.Ktot
        aO = self.reduce(O.sum(axis=0)) / self.Ktot
        F = self.reduce(O.conj().T @ ht) / self.Ktot - aO.conj() * havg
        return havg, aO, F

    def step(self, sampler: LITFIChainSampler, n_mc_steps: int, lr: float, fixed_cg_iters: Optional[int] = None,
             lam: Optional[float] = None) -> SRStats:
        """One iteration of propagate's loop body (optimizer.cuh:127-165)."""
        sampler.do_mcmc_steps(n_mc_steps)
        ht = sampler.get_htilda()
        O = sampler.get_lnpsiGradients()
        havg, aO, F = self.gradient(ht, O)
        st = SRStats()
        st.e_mean = havg
        if not np.isfinite(havg.real):
            st.finite = False
            return st
        st.lam = self.schedule() if lam is None else lam
        S = SMatrix(O, st.lam, self.reduce, self.Ktot)
        self.dx, st.cg_iters = pcg_solve(S, F, self.dx, self.tol, self.max_iter, fixed_cg_iters)
        sampler.evolve(self.dx, lr)
        h2 = self.reduce(np.array([(np.abs(ht) ** 2).sum()]))[0].real
        st.rsd = math.sqrt((h2 / self.Ktot - abs(havg) ** 2) / abs(havg) ** 2)
        st.F, st.dx = F, self.dx.copy()
        return st


# ----------------------------------------------------------------------------------------------
# exact diagonalisation known answer -- small N only
# ----------------------------------------------------------------------------------------------


def exact_ground_energy_per_site(N: int, J: float, h: float, alpha: float, pbc: bool = False) -> float:
    """E0/N of H = sum_{i<j} J_ij sz_i sz_j + h sum_i sx_i by dense/sparse diagonalisation (N <= 16)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    Jm = lr_coupling_matrix(N, J, alpha, pbc)
    dim = 1 << N
    states = np.arange(dim, dtype=np.int64)
    sz = np.empty((N, dim), dtype=np.float64)
    for i in range(N):
        sz[i] = 1.0 - 2.0 * ((states >> i) & 1)
    diag = np.zeros(dim)
    for i in range(N):
        for j in range(i + 1, N):
            diag += Jm[i, j] * sz[i] * sz[j]
    rows, cols = [], []
    for i in range(N):
        rows.append(states)
        cols.append(states ^ (1 << i))
    H = sp.coo_matrix((np.full(N * dim, h), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim)).tocsr()
    H = H + sp.diags(diag)
    if dim <= 4096:
        w = np.linalg.eigvalsh(H.toarray())
        return float(w[0]) / N
    w = spl.eigsh(H, k=1, which="SA", return_eigenvectors=False)
    return float(w[0]) / N
