"""CPU checks of the oracle's restatement of the translation-symmetric RBM (ref gpu/include/impl_neural_quantum_state.cuh:301-538,
kernels :1487-1553): it must be the plain RBM on the expanded weights, and its gradient the chain rule through the expansion.
(The reference ships no golden vectors for it; on the GPU box tests/test_gpu_reference_cuda.py pins engine and oracle to the
reference's own CUDA driver, gpu/src/LICH-train_rbmtrsymm.cu, compiled for sm_100.)"""
import numpy as np

from oracle import nqs_oracle as o


def _pair(N, al, K, seed):
    rng = np.random.default_rng(seed)
    t = o.RBMTrSymm(N, al, K, rng)
    t.variables[N * al] = 0.2 - 0.1j                       # a non-zero visible bias
    r = o.RBM(N, al * N, K)
    r.variables = np.concatenate([t.W.ravel(), t.a, t.b])
    spins = (2 * rng.integers(0, 2, size=(K, N)) - 1).astype(np.float64)
    return t, r, spins


def test_expansion_indices():
    N, al = 5, 2
    t = o.RBMTrSymm(N, al, 1, np.random.default_rng(0))
    w = t.variables[: N * al].reshape(al, N)
    W = t.W
    for i in range(N):
        for f in range(al):
            for j in range(N):
                assert W[i, f * N + j] == w[f, (i + j) % N]           # :1537-1540
    assert np.all(t.b.reshape(al, N) == t.variables[N * al + 1:][:, None])
    assert np.all(t.a == t.variables[N * al])


def test_sampler_side_equals_plain_rbm_on_expanded_weights():
    t, r, spins = _pair(6, 3, 7, 1)
    np.testing.assert_allclose(t.initialize(spins), r.initialize(spins), rtol=1e-14)
    for idx in (0, 3, 5):
        np.testing.assert_allclose(t.forward_flip(idx), r.forward_flip(idx), rtol=1e-13)
    mask = np.array([True, False, True, True, False, False, True])
    t.spin_flip(mask, 2)
    r.spin_flip(mask, 2)
    np.testing.assert_allclose(t.y, r.y, rtol=1e-14)
    np.testing.assert_allclose(t.sa, r.sa, rtol=1e-14)
    assert np.array_equal(t.spins, r.spins)


def test_gradient_is_chain_rule_through_the_expansion():
    N, al, K = 6, 2, 5
    t, r, spins = _pair(N, al, K, 2)
    t.initialize(spins)
    r.initialize(spins)
    Of = r.backward()                                        # [K][N*M + N + M], full RBM
    M = al * N
    want = np.zeros((K, t.P), dtype=np.complex128)
    for i in range(N):
        for f in range(al):
            for j in range(N):
                want[:, f * N + (i + j) % N] += Of[:, i * M + f * N + j]
    want[:, N * al] = Of[:, N * M:N * M + N].sum(axis=1)
    for f in range(al):
        want[:, N * al + 1 + f] = Of[:, N * M + N + f * N:N * M + N + (f + 1) * N].sum(axis=1)
    np.testing.assert_allclose(t.backward(), want, rtol=1e-12, atol=1e-14)


def test_gradient_matches_finite_differences():
    N, al, K = 4, 2, 3
    t, _, spins = _pair(N, al, K, 3)
    O = None
    base = t.initialize(spins)
    O = t.backward()
    eps = 1e-6
    for p in (0, 3, N * al - 1, N * al, N * al + 1, N * al + al):
        v0 = t.variables[p]
        t.variables[p] = v0 + eps
        up = t.initialize(spins)
        t.variables[p] = v0 - eps
        dn = t.initialize(spins)
        t.variables[p] = v0
        np.testing.assert_allclose((up - dn) / (2 * eps), O[:, p], rtol=1e-6, atol=1e-8)
    assert base.shape == (K,)


def test_variables_file_round_trip(tmp_path):
    t, _, _ = _pair(5, 2, 2, 4)
    path = str(tmp_path / "RBMTrSymmLICH-L5NF2A2T0.785398V0")
    t.save(path, 17)
    u = o.RBMTrSymm(5, 2, 2)
    u.load(path)
    assert np.array_equal(u.variables, t.variables)
    assert "\n" not in open(path).read()                     # one blank-separated line, no trailing newline (:474-482)
