"""The closed forms the transcendental-free RBM kernels (csrc/fast_kernels.cuh) are built on, checked in numpy against the
reference's literal arithmetic (oracle.logcosh: gpu/include/impl_neural_quantum_state.cuh:1238-1245).  These are the maths of
the kernels, not the kernels (those are checked on the GPU against the oracle, accept/reject decisions exactly)."""
import numpy as np

from oracle import nqs_oracle as o


def test_double_angle_flip_factor_and_state_update():
    rng = np.random.default_rng(0)
    n = 4000
    th = rng.normal(0, 1.5, n) + 1j * rng.normal(0, 2.0, n)          # theta_j = x + i y
    W = rng.normal(0, 0.4, n) + 1j * rng.normal(0, 0.6, n)
    sig = rng.choice([-1.0, 1.0], n)
    x, y = th.real, th.imag
    C2, S2, c2, s2 = np.cosh(2 * x), np.sinh(2 * x), np.cos(2 * y), np.sin(2 * y)      # the registers of the sweep kernel
    c4w, s4w, c4b, s4b = np.cosh(4 * W.real), np.sinh(4 * W.real), np.cos(4 * W.imag), np.sin(4 * W.imag)   # ftab_a / ftab_b
    # flipped factor: 2 |cosh(theta - 2 sigma W)|^2 = A + sigma B
    A = C2 * c4w + c2 * c4b
    B = s2 * s4b - S2 * s4w
    thp = th - 2 * sig * W
    want = 2 * np.exp(2 * o.logcosh(thp).real)
    np.testing.assert_allclose(A + sig * B, want, rtol=1e-11)
    # accepted flip: the four registers of theta' from those of theta (10 fp64 instructions per unit in the kernel)
    tys, tbs = sig * s4w, sig * s4b
    np.testing.assert_allclose(S2 * c4w - C2 * tys, np.sinh(2 * thp.real), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(C2 * c4w - S2 * tys, np.cosh(2 * thp.real), rtol=1e-10)
    np.testing.assert_allclose(c2 * c4b + s2 * tbs, np.cos(2 * thp.imag), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(s2 * c4b - c2 * tbs, np.sin(2 * thp.imag), rtol=1e-10, atol=1e-12)


def test_local_energy_flip_ratio_product_form():
    """psi(s^(i))/psi(s) = prod_j [cosh 2W_ij - s_i tanh(theta_j) sinh 2W_ij] exp(-2 s_i a_i)  (rbm_eloc_sites_kernel)."""
    rng = np.random.default_rng(1)
    N, M, K = 6, 9, 5
    m = o.RBM(N, M, K, rng)
    m.W[...] *= 6.0
    m.a[...] = 0.2 * (rng.normal(size=N) + 1j * rng.normal(size=N))
    spins = rng.choice([-1.0, 1.0], (K, N))
    ln0 = m.initialize(spins)
    T = np.tanh(m.y)
    for i in range(N):
        s = spins[:, i]
        f = np.cosh(2 * m.W[i])[None, :] - s[:, None] * T * np.sinh(2 * m.W[i])[None, :]
        ratio = f.prod(axis=1) * np.exp(-2 * s * m.a[i])
        np.testing.assert_allclose(ratio, np.exp(m.forward_flip(i) - ln0), rtol=1e-11)


def test_accept_test_on_products_equals_the_reference_rule():
    """u < P' A / R0 with P' = prod_j |cosh theta'_j|^2, R0 = prod_j |cosh theta_j|^2, A = exp(-4 s Re a)
    is the reference's u < exp(2 min(0, Re lnpsi' - Re lnpsi0)) (impl_mcmc_sampler.cuh:75-99) for u in [0, 1)."""
    rng = np.random.default_rng(2)
    N, M, K = 5, 40, 64
    m = o.RBM(N, M, K, rng)
    m.W[...] *= 5.0
    m.a[...] = 0.3 * (rng.normal(size=N) + 1j * rng.normal(size=N))
    spins = rng.choice([-1.0, 1.0], (K, N))
    ln0 = m.initialize(spins)
    u = rng.random(K)
    for i in range(N):
        ln1 = m.forward_flip(i)
        ref = u < np.exp(2 * np.minimum(0.0, ln1.real - ln0.real))
        s = spins[:, i]
        Pp = np.exp(2 * o.logcosh(m.y - 2 * s[:, None] * m.W[i][None, :]).real).prod(axis=1)
        R0 = np.exp(2 * o.logcosh(m.y).real).prod(axis=1)
        mine = u * R0 < Pp * np.exp(-4 * s * m.a[i].real)
        assert np.array_equal(ref, mine)
