"""Runs the reference's own CUDA driver (gpu/src/LICH-train_rbm.cu compiled for sm_100 by baseline/Makefile) as a subprocess.

BENCHMARK / TEST INFRASTRUCTURE: only bench.py's `gpu_reference` block and tests/test_gpu_reference_cuda.py use this; nothing
of the product imports it.  The binary is the unmodified reference source; the one substitution is the TRNG4 shim
(baseline/shim_cuda/trng: Philox keyed by (seed, chain, draw) exactly like libnqs_b200's internal generator), so for the same
`-seed` both programs see the same uniforms and can be compared iteration by iteration.
"""
from __future__ import annotations

import os
import re
import subprocess
import tempfile
import threading
import time
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(_HERE, "_ref", "LICH-train_rbm-gpu-ref")
BINARY_TRSYMM = os.path.join(_HERE, "_ref", "LICH-train_rbmtrsymm-gpu-ref")
THETA_STR = "0.785398"          # what the driver receives for -theta (J = sin, h = -cos of exactly this double)


def available(driver: str = "rbm") -> bool:
    b = BINARY if driver == "rbm" else BINARY_TRSYMM
    return os.path.exists(b) and os.access(b, os.X_OK)


def prefix_for(path: str, L: int, nh: int, alpha_str: str = "2", theta_str: str = THETA_STR, ver: int = 0, driver: str = "rbm") -> str:
    """gpu/src/LICH-train_rbm.cu:94: path + "RBMLICH-L" + L + "NH" + nh + "A" + alpha + "T" + theta + "V" + ver;
    gpu/src/LICH-train_rbmtrsymm.cu:89: "RBMTrSymmLICH-L" + L + "NF" + nf + ... (that one is the whole file name)."""
    if driver == "rbm":
        return os.path.join(path, "RBMLICH-L%dNH%dA%sT%sV%d" % (L, nh, alpha_str, theta_str, ver))
    return os.path.join(path, "RBMTrSymmLICH-L%dNF%dA%sT%sV%d" % (L, nh, alpha_str, theta_str, ver))


def run(L: int, nh: int, ns: int, niter: int, nwarm: int, seed: int, path: str, lr: float = 1e-2, nms: int = 1,
        device: int = 0, timeout: float = 600.0, driver: str = "rbm") -> dict:
    """One run of the reference driver; parameter files under `path` (prefix_for) are loaded if present and rewritten at the end.
    driver "rbmtrsymm": nh is the number of filters (-nf); that driver runs the chain with periodic boundaries."""
    # the reference's parser takes -option=value (cpu/include/argparse.hpp:19-116)
    opts = {"L": L, ("nh" if driver == "rbm" else "nf"): nh, "ns": ns, "niter": niter, "alpha": "2", "theta": THETA_STR, "ver": 0, "nwarm": nwarm, "nms": nms,
            "dev": device, "lr": repr(lr), "rsd": "1e-30", "seed": seed, "path": path}
    cmd = [BINARY if driver == "rbm" else BINARY_TRSYMM] + ["-%s=%s" % (k, v) for k, v in opts.items()]
    # The driver prints one row per SR iteration and flushes it (`<< std::endl << std::flush`, gpu/include/optimizer.cuh:156-159), so
    # the arrival time of each row on the pipe is the end of that iteration: per-iteration wall times without touching the program.
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, bufsize=1)
    killer = threading.Timer(timeout, proc.kill)
    killer.start()
    energies, rsd, row_times, lines = [], [], [], []
    elapsed = None
    try:
        for ln in proc.stdout:
            now = time.perf_counter()
            lines.append(ln)
            m = re.match(r"^\s*(\d+)\s+(\S+)\s+(\S+)\s*$", ln)
            if m and not ln.lstrip().startswith("#"):
                energies.append(float(m.group(2)))
                rsd.append(float(m.group(3)))
                row_times.append(now)
            m = re.match(r"^# elapsed time:\s*(\S+)\(sec\)", ln)
            if m:
                elapsed = float(m.group(1))
        rc = proc.wait()
    finally:
        killer.cancel()
    out = "".join(lines)
    if rc != 0:
        raise RuntimeError("reference CUDA driver failed (%d): %s" % (rc, out[-2000:]))
    return {"energies": energies, "rsd": rsd, "elapsed_s": elapsed, "row_times_s": row_times, "stdout_tail": out[-400:]}


def load_vars(path: str) -> np.ndarray:
    """every "(re,im)" token of one file (the RBMTrSymm variables file)."""
    txt = open(path).read()
    return np.array([complex(float(a), float(b)) for a, b in re.findall(r"\(([^,]+),([^)]+)\)", txt)])


def load_params(prefix: str, N: int, M: int) -> np.ndarray:
    """[W (i*M+j) | a | b] from the reference's text files (Dw / Da / Db, "(re,im)" tokens)."""
    def read(name):
        txt = open(prefix + name).read()
        return np.array([complex(float(a), float(b)) for a, b in re.findall(r"\(([^,]+),([^)]+)\)", txt)])
    w, a, b = read("Dw.dat"), read("Da.dat"), read("Db.dat")
    assert w.size == N * M and a.size == N and b.size == M
    return np.concatenate([w, a, b])
