"""In-tree build of libnqs_b200.so (hand-written CUDA for sm_100a behind the C ABI of include/nqs_b200.h).

The .so is written next to this file so that it travels to the GPU box with the repository snapshot.  nvcc
cross-compiles without a GPU, so this also is the "does it build" check of __graft_entry__.build().
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libnqs_b200.so")
STAMP = os.path.join(PKG_DIR, ".libnqs_b200.stamp")
SOURCES = ["engine.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libnqs_b200.so)")


def _source_digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/nqs_b200.h"]
    for name in files:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


HOST = os.path.join(PKG_DIR, "host")
BIN_DIR = os.path.join(PKG_DIR, "bin")
# reference-compatible command-line programs (C++ host over the C ABI; ref gpu/src/LICH-train_rbm.cu)
HOST_PROGRAMS = {
    "LICH-train_rbm-gpu": ("LICH-train_rbm.cpp", []),
    "LICH-train_ffnn-gpu": ("LICH-train_rbm.cpp", ["-DNQS_DRIVER_FFNN"]),
    "LICH-train_rbmtrsymm-gpu": ("LICH-train_rbm.cpp", ["-DNQS_DRIVER_RBMTRSYMM"]),
    "LICH-train_rbmz2prsymm-gpu": ("LICH-train_rbm.cpp", ["-DNQS_DRIVER_RBMZ2PRSYMM"]),
    "LICH-train_ffnntrsymm-gpu": ("LICH-train_rbm.cpp", ["-DNQS_DRIVER_FFNNTRSYMM"]),
}


def _gxx() -> str:
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found (needed to build the host programs)")


def build_host(force: bool = False) -> list:
    """Compile the C++ host programs against libnqs_b200.so (rpath $ORIGIN/.. so they run from the tree on the GPU box)."""
    os.makedirs(BIN_DIR, exist_ok=True)
    out = []
    for name, (src, defs) in HOST_PROGRAMS.items():
        exe = os.path.join(BIN_DIR, name)
        srcs = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [LIB_PATH]
        if not force and os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(f) for f in srcs):
            out.append(exe)
            continue
        cmd = [_gxx(), "-O2", "-std=c++17", "-Wall"] + defs + ["-o", exe, os.path.join(HOST, src), "-L" + PKG_DIR, "-lnqs_b200",
               "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath-link," + PKG_DIR]
        env = dict(os.environ)
        env.pop("CXX", None)
        proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if proc.returncode != 0:
            raise RuntimeError("g++ failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-4000:]))
        out.append(exe)
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libnqs_b200.so if the sources changed.  Returns the library path."""
    if not force and not needs_build():
        build_host()
        return LIB_PATH
    extra = os.environ.get("NQS_NVCC_EXTRA", "").split()     # e.g. -DNQS_RU_TRACE_BUILD (in-kernel phase clocks of the UMMA kernels)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-Xlinker", "-soname=libnqs_b200.so"]
    env = dict(os.environ)
    env.pop("CXX", None)  # the image exports a wrapper g++ that nvcc must not pick up
    env.pop("CC", None)
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose:
        sys.stderr.write(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-4000:]))
    with open(STAMP, "w") as f:
        f.write(_source_digest())
    build_host(force=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
