"""The drop-in boundary without a GPU: libnqs_b200.so loads, exports every entry point include/nqs_b200.h declares (and the
ctypes host binds exactly that set), reports its ABI version, and refuses to create a handle on a box without a CUDA device
(there is no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nqs_b200.h")


def declared_entry_points():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nqs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_entry_point():
    from neural_network_quantum_state_b200 import _lib as L
    lib = L.load()
    names = declared_entry_points()
    assert len(names) >= 40, names
    for n in names:
        assert hasattr(lib, n), "libnqs_b200.so does not export %s (declared in include/nqs_b200.h)" % n
    raw = C.CDLL(L.LIB_PATH) if hasattr(L, "LIB_PATH") else lib
    for n in names:
        getattr(raw, n)                       # dlsym: raises AttributeError when the symbol is missing


def test_ctypes_host_binds_exactly_the_declared_set():
    from neural_network_quantum_state_b200 import _lib as L
    src = open(os.path.join(ROOT, "neural_network_quantum_state_b200", "_lib.py")).read()
    bound = sorted(set(re.findall(r'"(nqs_[a-z0-9_]+)"\s*:', src)))
    assert bound == declared_entry_points()
    assert L.ABI_VERSION == int(re.search(r"#define\s+NQS_B200_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))


def test_abi_version_and_flag_values_match_the_header():
    from neural_network_quantum_state_b200 import _lib as L
    lib = L.load()
    assert lib.nqs_abi_version() == L.ABI_VERSION
    hdr = open(HEADER).read()
    for name, val in (("NO_SR", L.FLAG_NO_SR), ("ACCEPT_LOG", L.FLAG_ACCEPT_LOG), ("FORCE_GENERIC", L.FLAG_FORCE_GENERIC),
                      ("TWO_PASS_SV", L.FLAG_TWO_PASS_SV), ("SETUP_FROM_O", L.FLAG_SETUP_FROM_O),
                      ("STRUCTURED_SV", L.FLAG_STRUCTURED_SV), ("NO_DMMA", L.FLAG_NO_DMMA)):
        m = re.search(r"#define\s+NQS_FLAG_%s\s+(\d+)" % name, hdr)
        assert m and int(m.group(1)) == val, name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the refusal path cannot be shown")
    from neural_network_quantum_state_b200 import Engine, NQSError
    from neural_network_quantum_state_b200 import _lib as L
    with pytest.raises(NQSError) as ei:
        Engine("rbm", 8, 8, 16, -0.7, 0.7, 2.0)
    assert ei.value.status == L.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)
