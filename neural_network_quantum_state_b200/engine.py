"""Python host side of the C ABI (include/nqs_b200.h): one `Engine` == one nqs_handle == the chains of one GPU.

Method names follow the reference's sampler / optimizer interface (gpu/include/mcmc_sampler.cuh:21-29,
gpu/include/optimizer.cuh:117-121) so parity tests read like reference usage: warm_up, do_mcmc_steps, get_lnpsi,
get_htilda, get_lnpsiGradients, evolve, plus sr_step (= one iteration of StochasticReconfigurationCG::propagate).
All arrays are numpy on the host; device memory is owned by the library.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib as L


class NQSError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("%s: %s" % (L.STATUS_NAMES[status] if 0 <= status < len(L.STATUS_NAMES) else status, message))
        self.status = status


@dataclass
class SRResult:
    e_mean: complex
    rsd: float
    lam: float
    cg_iters: int
    finite: bool
    cg_res2: float
    cg_rhs2: float


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, model: str, n_inputs: int, n_hiddens: int, n_chains: int, h: float, J: float, alpha: float,
                 pbc: bool = False, order: str = "checkerboard", seed: int = 0, device: int = 0,
                 n_chains_total: int = 0, chain_offset: int = 0, max_predrawn_steps: int = 0,
                 sampler_only: bool = False, accept_log: bool = False, force_generic: bool = False,
                 two_pass_sv: bool = False, structured_sv: bool = False, no_dmma: bool = False):
        self.lib = L.load()
        self.model = model
        self.N, self.M, self.K = int(n_inputs), int(n_hiddens), int(n_chains)
        cfg = L.Config()
        cfg.abi_version = L.ABI_VERSION
        # tied ansaetze take the EXPANDED width: rbmtrsymm / ffnntrsymm n_hiddens = alpha*N, rbmz2prsymm n_hiddens = 4*alpha
        cfg.model = {"rbm": L.MODEL_RBM, "ffnn": L.MODEL_FFNN, "rbmtrsymm": L.MODEL_RBMTRSYMM,
                     "rbmz2prsymm": L.MODEL_RBMZ2PRSYMM, "ffnntrsymm": L.MODEL_FFNNTRSYMM}[model]
        cfg.n_inputs, cfg.n_hiddens, cfg.n_chains = self.N, self.M, self.K
        cfg.n_chains_total, cfg.chain_offset = int(n_chains_total), int(chain_offset)
        cfg.h, cfg.J, cfg.alpha, cfg.pbc = float(h), float(J), float(alpha), int(bool(pbc))
        cfg.order = {"checkerboard": L.ORDER_CHECKERBOARD, "sequential": L.ORDER_SEQUENTIAL}[order]
        cfg.seed, cfg.device = int(seed), int(device)
        cfg.flags = (L.FLAG_NO_SR if sampler_only else 0) | (L.FLAG_ACCEPT_LOG if accept_log else 0) | \
                    (L.FLAG_FORCE_GENERIC if force_generic else 0) | (L.FLAG_TWO_PASS_SV if two_pass_sv else 0) | \
                    (L.FLAG_STRUCTURED_SV if structured_sv else 0) | (L.FLAG_NO_DMMA if no_dmma else 0)
        cfg.max_predrawn_steps = int(max_predrawn_steps)
        self._h = C.c_void_p()
        # the environment opt-in of the structured S*v is for plain RBM / FNN handles: hidden from the library while a tied-variable
        # handle is made (the library reads the C environment: putenv / unsetenv through os.environ reach it)
        hide = os.environ.pop("NQS_STRUCTURED_SV", None) if model not in ("rbm", "ffnn") else None
        try:
            rc = self.lib.nqs_create(C.byref(cfg), C.byref(self._h))
        finally:
            if hide is not None:
                os.environ["NQS_STRUCTURED_SV"] = hide
        if rc != L.OK:
            raise NQSError(rc, (self.lib.nqs_last_error(None) or b"").decode())
        p = C.c_int64()
        self._chk(self.lib.nqs_n_variables(self._h, C.byref(p)))
        self.P = int(p.value)
        self._last_steps = 0

    # ---- plumbing
    def _chk(self, rc: int):
        if rc != L.OK:
            raise NQSError(rc, (self.lib.nqs_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.nqs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._chk(self.lib.nqs_sync(self._h))

    # ---- parameters
    def set_params(self, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        self._chk(self.lib.nqs_set_params(self._h, _ptr(v), v.size))

    def get_params(self) -> np.ndarray:
        v = np.empty(self.P, dtype=np.complex128)
        self._chk(self.lib.nqs_get_params(self._h, _ptr(v), v.size))
        return v

    def init_params_random(self, seed: int):
        self._chk(self.lib.nqs_init_params_random(self._h, int(seed)))

    def set_rng(self, kind: str, seed: int, seed_distance: int = 0):
        """kind "yarn2": the reference's trng::yarn2 stream, seed(seedNumber) + jump(2*seedDistance*chain) per chain
        (gpu/include/trng4cuda.cuh:41-54); "philox": the engine's counter generator.  Restarts the proposal counter."""
        kinds = {"philox": L.RNG_PHILOX, "yarn2": L.RNG_YARN2}
        if kind not in kinds:
            raise ValueError("rng kind must be 'philox' or 'yarn2'")
        mask = (1 << 64) - 1
        self._chk(self.lib.nqs_set_rng(self._h, kinds[kind], int(seed) & mask, int(seed_distance) & mask))

    def load(self, prefix: str):
        self._chk(self.lib.nqs_load_params(self._h, prefix.encode()))

    def save(self, prefix: str, precision: int = 10):
        self._chk(self.lib.nqs_save_params(self._h, prefix.encode(), int(precision)))

    # ---- sampler (reference BaseParallelSampler interface)
    def _spins_arg(self, spins):
        if spins is None:
            return None
        s = np.ascontiguousarray(np.asarray(spins).reshape(self.K, self.N), dtype=np.int8)
        return s

    def initialize(self, spins=None):
        s = self._spins_arg(spins)
        self._chk(self.lib.nqs_initialize(self._h, _ptr(s)))

    def warm_up(self, n_sweeps: int = 100, spins=None):
        s = self._spins_arg(spins)
        self._last_steps = int(n_sweeps) * self.N
        self._chk(self.lib.nqs_warm_up(self._h, int(n_sweeps), _ptr(s)))

    def do_mcmc_steps(self, n_sweeps: int = 1):
        self._last_steps = int(n_sweeps) * self.N
        self._chk(self.lib.nqs_do_mcmc_steps(self._h, int(n_sweeps)))

    def set_uniforms(self, u: Optional[np.ndarray]):
        if u is None:
            self._chk(self.lib.nqs_set_uniforms(self._h, None, 0))
            return
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.ndim == 2 and u.shape[1] == self.K
        self._u_keepalive = u
        self._chk(self.lib.nqs_set_uniforms(self._h, _ptr(u), u.shape[0]))

    def get_spinStates(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """`out`: a caller-owned C-contiguous int8 [K][N] buffer to fill instead of a fresh array (page-locked memory makes the
        device-to-host copy a single DMA instead of a staged one)."""
        s = np.empty((self.K, self.N), dtype=np.int8) if out is None else out
        assert s.dtype == np.int8 and s.shape == (self.K, self.N) and s.flags.c_contiguous
        self._chk(self.lib.nqs_get_spins(self._h, _ptr(s)))
        return s

    def get_lnpsi(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        v = np.empty(self.K, dtype=np.complex128) if out is None else out
        assert v.dtype == np.complex128 and v.shape == (self.K,) and v.flags.c_contiguous
        self._chk(self.lib.nqs_get_lnpsi(self._h, _ptr(v)))
        return v

    def get_theta(self) -> np.ndarray:
        v = np.empty((self.K, self.M), dtype=np.complex128)
        self._chk(self.lib.nqs_get_theta(self._h, _ptr(v)))
        return v

    def get_accept_log(self) -> np.ndarray:
        a = np.empty((self._last_steps, self.K), dtype=np.uint8)
        self._chk(self.lib.nqs_get_accept_log(self._h, _ptr(a), self._last_steps))
        return a.astype(bool)

    def forward_flip(self, site: int) -> np.ndarray:
        v = np.empty(self.K, dtype=np.complex128)
        self._chk(self.lib.nqs_forward_flip(self._h, int(site), _ptr(v)))
        return v

    def get_lnpsi_for_fixed_spins(self, spins) -> np.ndarray:
        s = self._spins_arg(spins)
        v = np.empty(self.K, dtype=np.complex128)
        self._chk(self.lib.nqs_lnpsi_fixed_spins(self._h, _ptr(s), _ptr(v)))
        return v

    # ---- measurement / optimisation
    def get_htilda(self, copy: bool = True) -> Optional[np.ndarray]:
        v = np.empty(self.K, dtype=np.complex128) if copy else None
        self._chk(self.lib.nqs_local_energy(self._h, _ptr(v)))
        return v

    def get_lnpsiGradients(self, copy: bool = True) -> Optional[np.ndarray]:
        O = np.empty((self.K, self.P), dtype=np.complex128) if copy else None
        self._chk(self.lib.nqs_log_derivs(self._h, _ptr(O)))
        return O

    def smatrix_dot(self, lam: float, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        Sv = np.empty(self.P, dtype=np.complex128)
        aO = np.empty(self.P, dtype=np.complex128)
        diag = np.empty(self.P, dtype=np.float64)
        self._chk(self.lib.nqs_smatrix_dot(self._h, float(lam), _ptr(v), _ptr(Sv), _ptr(aO), _ptr(diag)))
        return Sv, aO, diag

    def sr_step(self, n_mc_steps: int = 1, lr: float = 1e-2, tol: float = 1e-5, max_iter: int = 1000,
                fixed_iters: int = 0, lam: Optional[float] = None, apply_update: bool = True) -> SRResult:
        opt = L.SROptions()
        self._chk(self.lib.nqs_sr_options_default(C.byref(opt)))
        opt.lr, opt.tol, opt.max_iter, opt.fixed_iters = float(lr), float(tol), int(max_iter), int(fixed_iters)
        opt.lam = -1.0 if lam is None else float(lam)
        opt.n_mc_steps, opt.apply_update = int(n_mc_steps), int(bool(apply_update))
        st = L.SRStats()
        self._last_steps = int(n_mc_steps) * self.N
        self._chk(self.lib.nqs_sr_step(self._h, C.byref(opt), C.byref(st)))
        return SRResult(complex(st.e_re, st.e_im), st.rsd, st.lam, int(st.cg_iters), bool(st.finite), st.cg_res2, st.cg_rhs2)

    def sr_reset(self):
        """New optimisation run: lambda schedule back to its start, CG warm start zeroed (ref: a fresh StochasticReconfigurationCG)."""
        self._chk(self.lib.nqs_sr_reset(self._h))

    def get_sr_vectors(self):
        F = np.empty(self.P, dtype=np.complex128)
        dx = np.empty(self.P, dtype=np.complex128)
        self._chk(self.lib.nqs_get_sr_vectors(self._h, _ptr(F), _ptr(dx)))
        return F, dx

    def evolve(self, dx: np.ndarray, lr: float):
        dx = np.ascontiguousarray(dx, dtype=np.complex128)
        assert dx.size == self.P
        self._chk(self.lib.nqs_evolve(self._h, _ptr(dx), float(lr)))

    # ---- full-state checkpoint (sidecar next to the reference's parameter files)
    def save_state(self, path: str):
        """Variables, chain state, RNG counter, lambda-schedule state and CG warm start of this handle (one rank's shard)."""
        self._chk(self.lib.nqs_checkpoint_save(self._h, path.encode()))

    def load_state(self, path: str):
        """Continue bit-identically from save_state() of a handle with the same shape."""
        self._chk(self.lib.nqs_checkpoint_load(self._h, path.encode()))

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = L.load()
        buf = C.create_string_buffer(L.UNIQUE_ID_BYTES)
        rc = lib.nqs_comm_get_unique_id(buf)
        if rc != L.OK:
            raise NQSError(rc, (lib.nqs_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, L.UNIQUE_ID_BYTES)
        self._chk(self.lib.nqs_comm_init(self._h, int(n_ranks), int(rank), buf))

    def comm_p2p_export(self) -> bytes:
        buf = C.create_string_buffer(L.IPC_HANDLE_BYTES)
        self._chk(self.lib.nqs_comm_p2p_export(self._h, buf))
        return buf.raw

    def comm_p2p_import(self, handles) -> bool:
        """handles: the exported handles of all ranks in rank order.  Returns False (and keeps NCCL) if peers cannot be mapped."""
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob))
        rc = self.lib.nqs_comm_p2p_import(self._h, buf)
        if rc == L.ERR_UNSUPPORTED:
            return False
        self._chk(rc)
        return True

    def comm_p2p_disable(self):
        self._chk(self.lib.nqs_comm_p2p_disable(self._h))

    # ---- introspection
    def set_timing(self, on: bool = True):
        self._chk(self.lib.nqs_set_timing(self._h, int(on)))

    def get_timing(self) -> dict:
        t = L.Timing()
        self._chk(self.lib.nqs_get_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in L.Timing._fields_}

    def event_record(self, slot: int):
        self._chk(self.lib.nqs_event_record(self._h, int(slot)))

    def event_elapsed_ms(self, slot_begin: int, slot_end: int) -> float:
        ms = C.c_float()
        self._chk(self.lib.nqs_event_elapsed_ms(self._h, int(slot_begin), int(slot_end), C.byref(ms)))
        return float(ms.value)

    def kernel_variant(self, stage: str) -> str:
        return (self.lib.nqs_kernel_variant(self._h, stage.encode()) or b"").decode()
