"""Shared helpers for the parity tests."""
import numpy as np

# fp64 parity bar of the north star / SURVEY 8c: rel 1e-10 (abs 1e-12 near zero)
RTOL, ATOL = 1e-10, 1e-12


def assert_close(a, b, rtol=RTOL, atol=ATOL, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-300)
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert err <= atol + rtol * scale, "%s: max abs err %.3e (scale %.3e, rel %.3e)" % (what, err, scale, err / scale)


def ffnn_cpu_to_gpu_layout(v, N, M):
    """The reference CPU FFNN lays the W block of O / F / dx as i*M+j, the GPU FFNN as j*N+i.
    Golden vectors come from the CPU tree; the engine and oracle default to the GPU layout."""
    v = np.asarray(v)
    lead = v.shape[:-1]
    w = v[..., : N * M].reshape(lead + (N, M))
    w = np.swapaxes(w, -1, -2).reshape(lead + (N * M,))
    return np.concatenate([w, v[..., N * M:]], axis=-1)
