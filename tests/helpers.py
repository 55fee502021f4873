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


def audit_accepts(acc, acc_ref, U, ratio_log, what="accept masks"):
    """Exact accept/reject parity with SURVEY section 7's audit rule: the decision is `u < ratio` in fp64, and the only legitimate
    way two correct implementations can differ is a rounding tie, |u - ratio| < 1e-12 * ratio (the order of the sum over hidden
    units decides it).  Any other mismatch is a logic error and fails with the numbers needed to diagnose it.  A chain that hit
    a tie follows a different Markov path afterwards, so it is dropped from the comparison; the mask of chains that stayed
    comparable is returned (all True when the masks are equal, which is what every committed case produces)."""
    acc = np.asarray(acc, dtype=bool)
    acc_ref = np.asarray(acc_ref, dtype=bool)
    assert acc.shape == acc_ref.shape, "%s: shape %s vs %s" % (what, acc.shape, acc_ref.shape)
    keep = np.ones(acc.shape[1], dtype=bool)
    if np.array_equal(acc, acc_ref):
        return keep
    U = np.asarray(U)
    for k in np.unique(np.argwhere(acc != acc_ref)[:, 1]):
        t = int(np.argmax(acc[:, k] != acc_ref[:, k]))       # first proposal at which chain k differs
        u, r = float(U[t, k]), float(ratio_log[t][k])
        tie = abs(u - r) < 1e-12 * r
        assert tie, ("%s: chain %d differs first at proposal %d: engine %s, oracle %s, u = %.17g, oracle ratio = %.17g, "
                     "|u - ratio| / ratio = %.3e -- not a rounding tie (audit bound 1e-12): logic error"
                     % (what, k, t, bool(acc[t, k]), bool(acc_ref[t, k]), u, r, abs(u - r) / r))
        keep[k] = False
    assert keep.sum() >= 0.99 * keep.size, "%s: %d chains hit a rounding tie -- implausible" % (what, int((~keep).sum()))
    return keep
