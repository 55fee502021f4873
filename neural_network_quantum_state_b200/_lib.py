"""ctypes binding of libnqs_b200.so.  There is no fallback: if the CUDA library is missing the import of the engine
fails loudly (build it with `python -m neural_network_quantum_state_b200.build`)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 1
UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_IO, ERR_STATE, ERR_NCCL, ERR_NONFINITE, ERR_UNSUPPORTED = range(9)
STATUS_NAMES = ["NQS_OK", "NQS_ERR_INVALID", "NQS_ERR_CUDA", "NQS_ERR_NOMEM", "NQS_ERR_IO", "NQS_ERR_STATE",
                "NQS_ERR_NCCL", "NQS_ERR_NONFINITE", "NQS_ERR_UNSUPPORTED"]
MODEL_RBM, MODEL_FFNN, MODEL_RBMTRSYMM, MODEL_RBMZ2PRSYMM, MODEL_FFNNTRSYMM = 0, 1, 2, 3, 4
ORDER_CHECKERBOARD, ORDER_SEQUENTIAL = 0, 1
RNG_PHILOX, RNG_YARN2 = 0, 1
FLAG_NO_SR, FLAG_ACCEPT_LOG, FLAG_FORCE_GENERIC, FLAG_TWO_PASS_SV = 1, 2, 4, 8
FLAG_SETUP_FROM_O, FLAG_STRUCTURED_SV, FLAG_NO_DMMA = 16, 32, 64


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("model", C.c_int32), ("n_inputs", C.c_int32), ("n_hiddens", C.c_int32),
                ("n_chains", C.c_int64), ("n_chains_total", C.c_int64), ("chain_offset", C.c_int64),
                ("h", C.c_double), ("J", C.c_double), ("alpha", C.c_double), ("pbc", C.c_int32), ("order", C.c_int32),
                ("seed", C.c_uint64), ("device", C.c_int32), ("flags", C.c_int32), ("max_predrawn_steps", C.c_int64)]


class SRStats(C.Structure):
    _fields_ = [("e_re", C.c_double), ("e_im", C.c_double), ("rsd", C.c_double), ("lam", C.c_double),
                ("cg_iters", C.c_int32), ("finite", C.c_int32), ("cg_res2", C.c_double), ("cg_rhs2", C.c_double)]


class SROptions(C.Structure):
    _fields_ = [("lr", C.c_double), ("tol", C.c_double), ("max_iter", C.c_int32), ("fixed_iters", C.c_int32),
                ("lam", C.c_double), ("n_mc_steps", C.c_int32), ("apply_update", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("sweep_ms", C.c_float), ("eloc_ms", C.c_float), ("oderiv_ms", C.c_float), ("setup_ms", C.c_float),
                ("cg_ms", C.c_float), ("update_ms", C.c_float), ("rows_ms", C.c_float), ("cols_ms", C.c_float),
                ("rows_count", C.c_int32), ("cols_count", C.c_int32), ("kernel_launches", C.c_int64)]


# every symbol include/nqs_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _dbl, _cp = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_char_p
SYMBOLS = {
    "nqs_create": (_i32, [C.POINTER(Config), C.POINTER(_vp)]),
    "nqs_destroy": (None, [_vp]),
    "nqs_last_error": (_cp, [_vp]),
    "nqs_abi_version": (_i32, []),
    "nqs_sync": (_i32, [_vp]),
    "nqs_set_hamiltonian": (_i32, [_vp, _dbl, _dbl, _dbl, _i32, _i32]),
    "nqs_set_seed": (_i32, [_vp, C.c_uint64]),
    "nqs_set_rng": (_i32, [_vp, _i32, C.c_uint64, C.c_uint64]),
    "nqs_enable_sr": (_i32, [_vp]),
    "nqs_sr_reset": (_i32, [_vp]),
    "nqs_n_variables": (_i32, [_vp, C.POINTER(_i64)]),
    "nqs_set_params": (_i32, [_vp, _vp, _i64]),
    "nqs_get_params": (_i32, [_vp, _vp, _i64]),
    "nqs_init_params_random": (_i32, [_vp, C.c_uint64]),
    "nqs_load_params": (_i32, [_vp, _cp]),
    "nqs_save_params": (_i32, [_vp, _cp, _i32]),
    "nqs_initialize": (_i32, [_vp, _vp]),
    "nqs_warm_up": (_i32, [_vp, _i32, _vp]),
    "nqs_do_mcmc_steps": (_i32, [_vp, _i32]),
    "nqs_set_uniforms": (_i32, [_vp, _vp, _i64]),
    "nqs_get_spins": (_i32, [_vp, _vp]),
    "nqs_get_lnpsi": (_i32, [_vp, _vp]),
    "nqs_get_theta": (_i32, [_vp, _vp]),
    "nqs_get_accept_log": (_i32, [_vp, _vp, _i64]),
    "nqs_forward_flip": (_i32, [_vp, _i32, _vp]),
    "nqs_lnpsi_fixed_spins": (_i32, [_vp, _vp, _vp]),
    "nqs_local_energy": (_i32, [_vp, _vp]),
    "nqs_log_derivs": (_i32, [_vp, _vp]),
    "nqs_smatrix_dot": (_i32, [_vp, _dbl, _vp, _vp, _vp, _vp]),
    "nqs_sr_step": (_i32, [_vp, C.POINTER(SROptions), C.POINTER(SRStats)]),
    "nqs_sr_options_default": (_i32, [C.POINTER(SROptions)]),
    "nqs_get_sr_vectors": (_i32, [_vp, _vp, _vp]),
    "nqs_evolve": (_i32, [_vp, _vp, _dbl]),
    "nqs_comm_get_unique_id": (_i32, [_vp]),
    "nqs_comm_init": (_i32, [_vp, _i32, _i32, _vp]),
    "nqs_comm_p2p_export": (_i32, [_vp, _vp]),
    "nqs_comm_p2p_import": (_i32, [_vp, _vp]),
    "nqs_comm_p2p_disable": (_i32, [_vp]),
    "nqs_get_timing": (_i32, [_vp, C.POINTER(Timing)]),
    "nqs_set_timing": (_i32, [_vp, _i32]),
    "nqs_event_record": (_i32, [_vp, _i32]),
    "nqs_event_elapsed_ms": (_i32, [_vp, _i32, _i32, C.POINTER(C.c_float)]),
    "nqs_kernel_variant": (_cp, [_vp, _cp]),
    "nqs_checkpoint_save": (_i32, [_vp, _cp]),
    "nqs_checkpoint_load": (_i32, [_vp, _cp]),
}

_lib = None


def load(build_if_missing: bool = False):
    """dlopen libnqs_b200.so and bind every ABI symbol.  Raises if the library is absent or lacks a symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        if build_if_missing:
            _build.build()
        else:
            raise ImportError(
                "%s not found: the CUDA engine is not built (run `python -m neural_network_quantum_state_b200.build`). "
                "There is no CPU fallback." % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.nqs_abi_version() != ABI_VERSION:
        raise ImportError("libnqs_b200.so ABI version %d != binding %d" % (lib.nqs_abi_version(), ABI_VERSION))
    _lib = lib
    return lib
