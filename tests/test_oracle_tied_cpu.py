"""CPU checks of the oracle's restatement of the two other tied-variable ansaetze the reference drives on the long-range chain:
RBMZ2PrSymm (ref gpu/include/impl_neural_quantum_state.cuh:540-745, kernels :1556-1618; gpu/src/LICH-train_rbmz2prsymm.cu) and
FFNNTrSymm (:1019-1223, kernels :1693-1750; gpu/src/LICH-train_ffnntrsymm.cu).  Each must be the plain network on the expanded
weights, and its gradient the chain rule through the expansion.  (The reference ships no golden vectors for them; on the GPU box
tests/test_gpu_tied.py pins engine and oracle to the reference's own CUDA drivers compiled for sm_100.)"""
import numpy as np
import pytest

from oracle import nqs_oracle as o


def _z2(N, al, K, seed):
    rng = np.random.default_rng(seed)
    t = o.RBMZ2PrSymm(N, al, K, rng)
    t.variables *= 5.0
    r = o.RBM(N, 4 * al, K)
    r.variables = np.concatenate([t.W.ravel(), t.a, t.b])
    spins = (2 * rng.integers(0, 2, size=(K, N)) - 1).astype(np.float64)
    return t, r, spins


def _ftr(N, al, K, seed):
    rng = np.random.default_rng(seed)
    t = o.FFNNTrSymm(N, al, K, rng)
    t.variables[N * al: N * al + al] = 0.1 * (rng.normal(size=al) + 1j * rng.normal(size=al))    # non-zero hidden biases
    r = o.FFNN(N, al * N, K, transposed_grad=False)
    r.variables = np.concatenate([t.W.ravel(), t.b, t.w1o])
    spins = (2 * rng.integers(0, 2, size=(K, N)) - 1).astype(np.float64)
    return t, r, spins


def test_z2pr_expansion_indices():
    N, al = 5, 3
    t = o.RBMZ2PrSymm(N, al, 1, np.random.default_rng(0))
    w = t.variables[: N * al].reshape(N, al)
    W = t.W
    for i in range(N):
        for f in range(al):                                                   # :1600-1607
            assert W[i, 4 * f + 0] == w[i, f] and W[i, 4 * f + 1] == -w[i, f]
            assert W[i, 4 * f + 2] == w[N - 1 - i, f] and W[i, 4 * f + 3] == -w[N - 1 - i, f]
    assert np.all(t.b.reshape(al, 4) == t.variables[N * al:][:, None])
    assert t.P == N * al + al and t.M == 4 * al


def test_ffnntr_expansion_indices():
    N, al = 5, 2
    t = o.FFNNTrSymm(N, al, 1, np.random.default_rng(0))
    w = t.variables[: N * al].reshape(al, N)
    W = t.W
    for i in range(N):
        for f in range(al):
            for j in range(N):
                assert W[i, f * N + j] == w[f, (i + j) % N]                   # :1706
    assert np.all(t.b.reshape(al, N) == t.variables[N * al: N * al + al][:, None])
    assert np.all(t.w1o.reshape(al, N) == t.variables[N * al + al:][:, None])
    assert t.P == N * al + 2 * al


@pytest.mark.parametrize("make", [_z2, _ftr])
def test_sampler_side_equals_plain_network_on_expanded_weights(make):
    t, r, spins = make(6, 3, 7, 1)
    np.testing.assert_allclose(t.initialize(spins), r.initialize(spins), rtol=1e-13)
    for idx in (0, 3, 5):
        np.testing.assert_allclose(t.forward_flip(idx), r.forward_flip(idx), rtol=1e-12, atol=1e-14)
    mask = np.array([True, False, True, True, False, False, True])
    t.spin_flip(mask, 2)
    r.spin_flip(mask, 2)
    np.testing.assert_allclose(t.y, r.y, rtol=1e-13)
    assert np.array_equal(t.spins, r.spins)
    np.testing.assert_allclose(t.forward_spins(-spins, save=False), r.forward_spins(-spins, save=False), rtol=1e-13)


def test_z2pr_amplitude_is_even_under_global_flip_and_reflection():
    t, _, spins = _z2(6, 2, 5, 9)
    a = t.initialize(spins)
    np.testing.assert_allclose(t.initialize(-spins), a, rtol=1e-13)            # Z2
    np.testing.assert_allclose(t.initialize(spins[:, ::-1]), a, rtol=1e-13)    # parity


def test_ffnntr_amplitude_is_translation_invariant():
    t, _, spins = _ftr(6, 2, 5, 9)
    a = t.initialize(spins)
    np.testing.assert_allclose(t.initialize(np.roll(spins, 2, axis=1)), a, rtol=1e-12)


def test_z2pr_gradient_is_chain_rule_through_the_expansion():
    N, al, K = 6, 2, 5
    t, r, spins = _z2(N, al, K, 2)
    t.initialize(spins)
    r.initialize(spins)
    Of = r.backward()                                                          # [K][N*M + N + M]
    M = 4 * al
    want = np.zeros((K, t.P), dtype=np.complex128)
    for i in range(N):
        for f in range(al):
            want[:, i * al + f] += Of[:, i * M + 4 * f] - Of[:, i * M + 4 * f + 1]
            want[:, (N - 1 - i) * al + f] += Of[:, i * M + 4 * f + 2] - Of[:, i * M + 4 * f + 3]
    for f in range(al):
        want[:, N * al + f] = Of[:, N * M + N + 4 * f: N * M + N + 4 * f + 4].sum(axis=1)
    np.testing.assert_allclose(t.backward(), want, rtol=1e-12, atol=1e-14)


def test_ffnntr_gradient_is_chain_rule_through_the_expansion():
    N, al, K = 6, 2, 5
    t, r, spins = _ftr(N, al, K, 2)
    t.initialize(spins)
    r.initialize(spins)
    Of = r.backward()                                                          # natural layout [W1 (i*M+c) | b1 | w1o]
    M = al * N
    want = np.zeros((K, t.P), dtype=np.complex128)
    for i in range(N):
        for f in range(al):
            for j in range(N):
                want[:, f * N + (i + j) % N] += Of[:, i * M + f * N + j]
    for f in range(al):
        want[:, N * al + f] = Of[:, N * M + f * N: N * M + (f + 1) * N].sum(axis=1)
        want[:, N * al + al + f] = Of[:, N * M + M + f * N: N * M + M + (f + 1) * N].sum(axis=1)
    np.testing.assert_allclose(t.backward(), want, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("make", [_z2, _ftr])
def test_gradient_matches_finite_differences(make):
    N, al, K = 4, 2, 3
    t, _, spins = make(N, al, K, 3)
    t.initialize(spins)
    O = t.backward()
    eps = 1e-6
    for p in range(t.P):
        v0 = t.variables[p]
        t.variables[p] = v0 + eps
        up = t.initialize(spins)
        t.variables[p] = v0 - eps
        dn = t.initialize(spins)
        t.variables[p] = v0
        np.testing.assert_allclose((up - dn) / (2 * eps), O[:, p], rtol=2e-6, atol=1e-8)


@pytest.mark.parametrize("cls,name", [(o.RBMZ2PrSymm, "RBMZ2PrSymmLICH-L5NF2A2T0.785398V0"), (o.FFNNTrSymm, "FFNNTrSymmLICH-L5NF2A2T0.785398V0")])
def test_variables_file_round_trip(tmp_path, cls, name):
    t = cls(5, 2, 2, np.random.default_rng(4))
    path = str(tmp_path / name)
    t.save(path, 17)
    u = cls(5, 2, 2)
    u.load(path)
    assert np.array_equal(u.variables, t.variables)
    assert "\n" not in open(path).read()                                       # one blank-separated line (:680-689, :1167-1176)


def test_sr_step_runs_through_the_oracle_sampler():
    """the tied ansaetze plug into the same sampler / optimiser restatement as the plain ones"""
    for kind, M in (("rbmz2prsymm", 8), ("ffnntrsymm", 12)):
        N, K = 6, 40
        rng = np.random.default_rng(5)
        m = o.make_ansatz(kind, N, M, K, rng)
        s = o.LITFIChainSampler(m, -0.7, 0.7, 2.0, kind == "ffnntrsymm", o.UniformSource(K, predrawn=rng.random((6 * N, K))))
        s.warm_up(3)
        st = o.StochasticReconfigurationCG(K, m.P).step(s, 1, 0.02)
        assert np.isfinite(st.e_mean.real) and st.cg_iters >= 1
