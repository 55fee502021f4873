"""Translation-symmetric RBM (ref RBMTrSymm, gpu/include/impl_neural_quantum_state.cuh:301-538, driven by
gpu/src/LICH-train_rbmtrsymm.cu on the PERIODIC chain): the CUDA engine against the numpy oracle."""
import math

import numpy as np
import pytest

from helpers import assert_close, audit_accepts
from oracle import nqs_oracle as o

pytestmark = pytest.mark.gpu
H, J, ALPHA = -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0


def _vars(N, al, rng, scale=4.0):
    t = o.RBMTrSymm(N, al, 1, rng)
    v = t.variables * scale
    v[N * al] = 0.15 - 0.05j
    return v


@pytest.mark.parametrize("N,al,K", [(16, 2, 130), (24, 1, 64), (32, 3, 77), (64, 2, 96), (128, 2, 40)])
@pytest.mark.parametrize("force_generic", [False, True])
def test_sampler_energy_gradients_match_oracle(N, al, K, force_generic):
    from neural_network_quantum_state_b200 import Engine
    rng = np.random.default_rng(100 * N + al)
    v = _vars(N, al, rng)
    n_warm, n_more = 4, 2
    U = rng.random(((n_warm + n_more) * N, K))
    m = o.RBMTrSymm(N, al, K)
    m.variables = v.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, True, o.UniformSource(K, predrawn=U))
    s.record = True
    e = Engine("rbmtrsymm", N, al * N, K, H, J, ALPHA, pbc=True, max_predrawn_steps=U.shape[0], accept_log=True,
               force_generic=force_generic)
    assert e.P == N * al + 1 + al
    e.set_params(v)
    e.set_uniforms(U)
    s.warm_up(n_warm)
    e.warm_up(n_warm)
    keep = audit_accepts(e.get_accept_log(), np.array(s.accept_log), U[:n_warm * N], s.ratio_log)
    assert keep.all()
    assert np.array_equal(e.get_spinStates(), m.spins.astype(np.int8))
    assert_close(e.get_theta(), m.y, what="theta")
    assert_close(e.get_lnpsi(), s.lnpsi0, what="lnpsi0")
    s.accept_log, s.ratio_log = [], []
    s.do_mcmc_steps(n_more)
    e.do_mcmc_steps(n_more)
    assert audit_accepts(e.get_accept_log(), np.array(s.accept_log), U[n_warm * N:], s.ratio_log).all()
    assert_close(e.get_htilda(), s.get_htilda(), what="htilda")
    O = s.get_lnpsiGradients()
    assert_close(e.get_lnpsiGradients(), O, what="O")
    vv = rng.normal(size=m.P) + 1j * rng.normal(size=m.P)
    S = o.SMatrix(O, 0.41)
    Sv, aO, diag = e.smatrix_dot(0.41, vv)
    assert_close(aO, S.aO, what="<O>")
    assert_close(diag, S.diag, rtol=1e-9, what="diag S")
    assert_close(Sv, S.dot(vv), rtol=1e-9, what="S v")
    spins = (2 * rng.integers(0, 2, size=(K, N)) - 1)
    mm = o.RBMTrSymm(N, al, K)
    mm.variables = v.copy()
    assert_close(e.get_lnpsi_for_fixed_spins(spins), mm.forward_spins(spins, save=False), what="forward(spins)")
    e.close()


@pytest.mark.parametrize("N,al,K", [(16, 2, 400), (32, 2, 300)])
def test_sr_trajectory_matches_oracle(N, al, K):
    from neural_network_quantum_state_b200 import Engine
    rng = np.random.default_rng(9)
    v = _vars(N, al, rng, scale=1.0)
    U = rng.random((14 * N, K))
    spins0 = (2 * rng.integers(0, 2, size=(K, N)) - 1).astype(np.float64)     # random start: no zero-variance columns
    m = o.RBMTrSymm(N, al, K)
    m.variables = v.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, True, o.UniformSource(K, predrawn=U))
    e = Engine("rbmtrsymm", N, al * N, K, H, J, ALPHA, pbc=True, max_predrawn_steps=U.shape[0])
    e.set_params(v)
    e.set_uniforms(U)
    s.warm_up(8, spins0)
    e.warm_up(8, spins0.astype(np.int8))
    sr = o.StochasticReconfigurationCG(K, m.P)
    for it in range(4):
        st_o = sr.step(s, 1, 0.03)
        st = e.sr_step(n_mc_steps=1, lr=0.03)
        assert st.cg_iters == st_o.cg_iters
        assert_close(st.e_mean, st_o.e_mean, what="<H>")
        F, dx = e.get_sr_vectors()
        assert_close(F, st_o.F, what="F")
        assert_close(dx, st_o.dx, rtol=1e-6, what="dx")
    assert_close(e.get_params(), m.variables, rtol=1e-8, what="variables")
    e.close()


def test_variables_file_and_pynqs(tmp_path):
    from neural_network_quantum_state_b200 import Engine
    from neural_network_quantum_state_b200.pynqs import sampler as pysampler
    N, al, K = 12, 2, 32
    rng = np.random.default_rng(4)
    v = _vars(N, al, rng)
    m = o.RBMTrSymm(N, al, K)
    m.variables = v.copy()
    path = str(tmp_path / "vars")
    m.save(path, 17)
    e = Engine("rbmtrsymm", N, al * N, 4, H, J, ALPHA, pbc=True, sampler_only=True)
    e.load(path)
    assert np.array_equal(e.get_params(), v)
    out = str(tmp_path / "out")
    e.save(out, 17)
    assert open(out).read() == open(path).read()
    e.close()
    r = pysampler.RBM(floatType="float64", symmType="tr")
    r.init(nInputs=N, nHiddens=al, nChains=K, seedNumber=3, seedDistance=1000, path_to_load=path, init_mcmc_steps=5)
    r.do_mcmc_steps(2)
    sp = r.get_spinStates()
    assert sp.shape == (K, N) and set(np.unique(sp)) <= {-1.0, 1.0}
    mm = o.RBMTrSymm(N, al, K)
    mm.variables = v.copy()
    assert_close(r.get_lnpsi(), mm.forward_spins(sp), what="pynqs get_lnpsi")
    assert_close(r.get_lnpsi_for_fixed_spins(sp), mm.forward_spins(sp), what="pynqs get_lnpsi_for_fixed_spins")
