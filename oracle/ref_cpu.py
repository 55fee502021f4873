"""ctypes front-end of oracle/_ref/libnqs_ref.so -- the reference's own CPU implementation (TEST INFRASTRUCTURE).

See oracle/ref_harness.cpp for what is the unmodified reference (RBM/FFNN, BaseParallelSampler, SMatrixForCG,
ConjugateGradient from /root/reference/cpu/include) and what is shim (long-range Hamiltonian, SR loop body with the
GPU solver settings, TRNG feed).  Only tests/, tests/golden/make_golden.py and bench.py's reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libnqs_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(
                "%s missing: run `make -C oracle` in the development container (needs /root/reference)" % _LIB_PATH)
        L = C.CDLL(_LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int] * 4 + [C.c_double] * 3 + [C.c_int] * 2
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_destroy.restype = None
        L.ref_n_variables.argtypes = [C.c_void_p]
        vp, ip, dbl = C.c_void_p, C.c_int, C.c_double
        for name, args in {
            "ref_set_uniforms": [vp, vp, C.c_long],
            "ref_set_params": [vp, vp], "ref_get_params": [vp, vp],
            "ref_load": [vp, C.c_char_p], "ref_save": [vp, C.c_char_p, ip],
            "ref_set_initial_spins": [vp, vp],
            "ref_warm_up": [vp, ip], "ref_do_mcmc_steps": [vp, ip],
            "ref_get_lnpsi": [vp, vp], "ref_get_spins": [vp, vp], "ref_get_y": [vp, vp],
            "ref_forward_flip": [vp, ip, vp], "ref_get_htilda": [vp, vp], "ref_get_gradients": [vp, vp],
            "ref_smatrix_set": [vp, vp, dbl, vp, vp], "ref_smatrix_dot": [vp, vp, vp], "ref_smatrix_diag": [vp, vp],
            "ref_evolve": [vp, vp, dbl],
            "ref_sr_step": [vp, ip, dbl, ip, dbl, dbl, vp, vp, vp],
        }.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_set_threads.restype = None
        L.ref_max_threads.restype = C.c_int
        _lib = L
    return _lib


def set_blas_threads(n: int) -> bool:
    """Thread count of the OpenBLAS the reference harness is linked against (openblas_set_num_threads on the copy that is
    already mapped into this process).  Returns False if that symbol cannot be found (the OPENBLAS_NUM_THREADS value stays)."""
    lib()
    try:
        with open("/proc/self/maps") as f:
            paths = {ln.split()[-1] for ln in f if "libopenblas" in ln}
        for path in paths:
            fn = getattr(C.CDLL(path), "openblas_set_num_threads", None)
            if fn is not None:
                fn.argtypes = [C.c_int]
                fn.restype = None
                fn(int(n))
                return True
    except Exception:
        pass
    return False


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class RefSampler:
    """Reference CPU RBM/FFNN + sampler + long-range TFI shim + SR-CG, driven with pre-drawn uniforms.

    model: "rbm" | "ffnn" | "rbmtrsymm" | "ffnntrsymm";  order: "checkerboard" (LITFIChain) | "sequential" (Sampler4SpinHalf).
    NOTE for "ffnn": the CPU tree emits the W-block of O in natural (i*M+j) layout; the GPU tree transposes.
    """

    def __init__(self, model: str, N: int, M: int, K: int, h: float, J: float, alpha: float, pbc: bool = False,
                 order: str = "checkerboard"):
        self.L = lib()
        self.model, self.N, self.M, self.K = model, N, M, K
        # the tied ansaetze (cpu/include/neural_quantum_state.hpp:68-102, 184-217) are constructed with the number of filters;
        # M here is the expanded width alpha*N (the shape of theta), as in nqs_config.n_hiddens
        width = M // N if model in ("rbmtrsymm", "ffnntrsymm") else M
        self.h = self.L.ref_create({"rbm": 0, "ffnn": 1, "rbmtrsymm": 2, "ffnntrsymm": 4}[model], N, width, K, h, J, alpha, int(pbc),
                                   {"checkerboard": 0, "sequential": 1}[order])
        if not self.h:
            raise RuntimeError("ref_create failed")
        self.P = self.L.ref_n_variables(self.h)
        self._uniforms = None

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("reference harness call failed")

    def set_uniforms(self, u: np.ndarray):
        """u[steps][K]; consumed from the sampler's CURRENT draw count (the feed is indexed by total draws)."""
        self._uniforms = np.ascontiguousarray(u, dtype=np.float64)
        self._chk(self.L.ref_set_uniforms(self.h, _p(self._uniforms), self._uniforms.shape[0]))

    def set_params(self, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        assert v.size == self.P
        self._chk(self.L.ref_set_params(self.h, _p(v)))

    def get_params(self) -> np.ndarray:
        v = np.empty(self.P, dtype=np.complex128)
        self._chk(self.L.ref_get_params(self.h, _p(v)))
        return v

    def load(self, prefix: str):
        self._chk(self.L.ref_load(self.h, prefix.encode()))

    def save(self, prefix: str, prec: int = 10):
        self._chk(self.L.ref_save(self.h, prefix.encode(), prec))

    def set_initial_spins(self, s: np.ndarray):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.size == self.K * self.N
        self._chk(self.L.ref_set_initial_spins(self.h, _p(s)))

    def warm_up(self, n: int):
        self._chk(self.L.ref_warm_up(self.h, n))

    def do_mcmc_steps(self, n: int):
        self._chk(self.L.ref_do_mcmc_steps(self.h, n))

    def _get(self, fn, shape, dtype=np.complex128):
        out = np.empty(shape, dtype=dtype)
        self._chk(fn(self.h, _p(out)))
        return out

    def get_lnpsi(self):
        return self._get(self.L.ref_get_lnpsi, self.K)

    def get_spins(self):
        return self._get(self.L.ref_get_spins, (self.K, self.N), np.float64)

    def get_y(self):
        return self._get(self.L.ref_get_y, (self.K, self.M))

    def forward_flip(self, idx: int):
        out = np.empty(self.K, dtype=np.complex128)
        self._chk(self.L.ref_forward_flip(self.h, idx, _p(out)))
        return out

    def get_htilda(self):
        return self._get(self.L.ref_get_htilda, self.K)

    def get_gradients(self):
        return self._get(self.L.ref_get_gradients, (self.K, self.P))

    def smatrix_set(self, O: np.ndarray, lam: float):
        O = np.ascontiguousarray(O, dtype=np.complex128)
        aO = np.empty(self.P, dtype=np.complex128)
        diag = np.empty(self.P, dtype=np.float64)
        self._chk(self.L.ref_smatrix_set(self.h, _p(O), lam, _p(aO), _p(diag)))
        return aO, diag

    def smatrix_dot(self, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.zeros(self.P, dtype=np.complex128)
        self._chk(self.L.ref_smatrix_dot(self.h, _p(v), _p(out)))
        return out

    def smatrix_diag(self):
        """diag of the S matrix most recently set (by smatrix_set or inside sr_step)."""
        return self._get(self.L.ref_smatrix_diag, self.P, np.float64)

    def evolve(self, dx: np.ndarray, lr: float):
        dx = np.ascontiguousarray(dx, dtype=np.complex128)
        self._chk(self.L.ref_evolve(self.h, _p(dx), lr))

    def sr_step(self, nms: int, lr: float, max_iter: int = 1000, tol: float = 1e-5, lam: Optional[float] = None):
        st = np.zeros(6, dtype=np.float64)
        F = np.zeros(self.P, dtype=np.complex128)
        dx = np.zeros(self.P, dtype=np.complex128)
        self._chk(self.L.ref_sr_step(self.h, nms, lr, max_iter, tol, -1.0 if lam is None else lam, _p(st), _p(F), _p(dx)))
        return {"e_mean": complex(st[0], st[1]), "rsd": st[2], "lam": st[3], "cg_iters": int(st[4]),
                "finite": bool(st[5]), "F": F, "dx": dx}
