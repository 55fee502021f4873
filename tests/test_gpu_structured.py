"""fp64 tensor-core (DMMA) kernels of csrc/sv_struct.cuh through the C ABI:
  * theta = S W + b (+ fused lnpsi epilogue) against the scalar-FMA kernel and the oracle;
  * structured S*v (NQS_FLAG_STRUCTURED_SV: two GEMMs on the factors of O, no O matrix) against the golden vectors of the
    reference CPU build, the numpy oracle and the explicit-O kernels of the same engine.
Bars as everywhere: fp64 quantities rel 1e-10; dx after a converged CG rel 1e-6 with the same iteration count.
"""
import numpy as np
import pytest

from helpers import assert_close
from oracle import nqs_oracle as o
from test_gpu_parity import ALPHA, CASES, H, J, _engine, engine_from_golden, synth, to_gpu

pytestmark = pytest.mark.gpu

# shapes that reach every tile variant of spin_cols_dmma_kernel (N <= 16, 32, 64, 128, 256), ragged M and K
STRUCT_CASES = CASES + [
    ("rbm", 8, 5, 37, False),
    ("rbm", 64, 128, 300, False),
    ("rbm", 100, 72, 129, True),
    ("rbm", 128, 256, 520, False),   # cfg3 network
    ("rbm", 200, 40, 65, False),
    ("ffnn", 128, 96, 257, False),
]


def test_structured_golden_smatrix_and_trajectory(golden):
    g = golden
    e = engine_from_golden(g, structured_sv=True)
    assert e.kernel_variant("sv").startswith("structured_dmma"), e.kernel_variant("sv")
    e.warm_up(g["n_warm"], g.get("init_spins"))
    assert np.array_equal(e.get_spinStates(), g["warm_spins"])
    assert_close(e.get_theta(), g["warm_y"], what="theta")
    assert_close(e.get_htilda(), g["htilda"], what="htilda")
    Sv, aO, diag = e.smatrix_dot(g["sm_lambda"], to_gpu(g, g["sm_v"]))
    assert_close(aO, to_gpu(g, g["sm_aO"]), what="<O>")
    assert_close(diag, to_gpu(g, g["sm_diag"]), what="diag S")
    assert_close(Sv, to_gpu(g, g["sm_Sv"]), what="S v (structured)")
    # O is still available on request (allocated lazily) and unchanged
    assert_close(e.get_lnpsiGradients(), to_gpu(g, g["O"]), what="O")
    for it in range(g["n_sr"]):
        st = e.sr_step(n_mc_steps=1, lr=g["lr"])
        assert st.finite
        assert_close(st.e_mean, g["sr_E"][it], what="<H> it %d" % it)
        assert st.cg_iters == int(g["sr_cg_iters"][it])
        F, dx = e.get_sr_vectors()
        assert_close(F, to_gpu(g, g["sr_F"][it]), what="F it %d" % it)
        assert_close(dx, to_gpu(g, g["sr_dx"][it]), rtol=1e-6, what="dx it %d" % it)
    assert_close(e.get_params(), g["final_params"], rtol=1e-8, what="params")
    assert np.array_equal(e.get_spinStates(), g["final_spins"])
    e.close()


@pytest.mark.parametrize("model,N,M,K,pbc", STRUCT_CASES)
def test_structured_sv_matches_explicit_and_oracle(model, N, M, K, pbc):
    rng = np.random.default_rng(N * 77 + M)
    params = synth(model, N, M, rng)
    P = params.size
    v = rng.normal(size=P) + 1j * rng.normal(size=P)
    res = []
    for structured in (False, True):
        e = _engine(model, N, M, K, H, J, ALPHA, pbc=pbc, seed=4242, structured_sv=structured)
        e.set_params(params)
        e.warm_up(3)
        e.get_htilda()
        if not structured:
            O = e.get_lnpsiGradients()
        res.append(e.smatrix_dot(0.37, v))
        e.close()
    for a, b, what in zip(res[1], res[0], ("S v", "<O>", "diag")):
        assert_close(a, b, what="%s structured vs explicit" % what)
    # oracle-independent statement of the product from the explicit O of the other engine
    aO = O.mean(axis=0)
    diag = (np.abs(O) ** 2).mean(axis=0) - np.abs(aO) ** 2
    Sv = O.conj().T @ (O @ v) / K - aO.conj() * (aO @ v) + 0.37 * diag * v
    assert_close(res[1][0], Sv, what="S v structured vs numpy")


@pytest.mark.parametrize("model,N,M,K,pbc", [CASES[1], CASES[3], CASES[6], STRUCT_CASES[-3]])
def test_structured_sr_matches_oracle_fixed_cg_iterations(model, N, M, K, pbc):
    rng = np.random.default_rng(11)
    params = synth(model, N, M, rng)
    U = rng.random((8 * N, K))
    m = o.make_ansatz(model, N, M, K)
    m.variables = params.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, pbc, o.UniformSource(K, predrawn=U))
    e = _engine(model, N, M, K, H, J, ALPHA, pbc=pbc, max_predrawn_steps=U.shape[0], structured_sv=True)
    e.set_params(params)
    e.set_uniforms(U)
    s.warm_up(5)
    e.warm_up(5)
    sr = o.StochasticReconfigurationCG(K, m.P)
    for it in range(2):
        st_o = sr.step(s, 1, 0.03, fixed_cg_iters=6, lam=0.5)
        st = e.sr_step(n_mc_steps=1, lr=0.03, fixed_iters=6, lam=0.5)
        assert_close(st.e_mean, st_o.e_mean, what="<H>")
        F, dx = e.get_sr_vectors()
        assert_close(F, st_o.F, what="F")
        assert_close(dx, st_o.dx, rtol=1e-9, what="dx (6 CG its)")
    assert_close(e.get_params(), m.variables, rtol=1e-9, what="params")
    e.close()


@pytest.mark.parametrize("model,N,M,K,pbc", STRUCT_CASES)
def test_dmma_theta_and_lnpsi_match_fma_kernel_and_oracle(model, N, M, K, pbc):
    rng = np.random.default_rng(N + 31 * M)
    params = synth(model, N, M, rng)
    spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(K, N))
    outs = []
    for no_dmma in (False, True):
        e = _engine(model, N, M, K, H, J, ALPHA, seed=1, sampler_only=True, no_dmma=no_dmma)
        e.set_params(params)
        e.initialize(spins)
        assert e.kernel_variant("theta") == ("tiled_fma" if no_dmma else "dmma_rows")
        other = -spins
        outs.append((e.get_theta(), e.get_lnpsi(), e.get_lnpsi_for_fixed_spins(other)))
        e.close()
    assert_close(outs[0][0], outs[1][0], what="theta dmma vs fma")
    assert_close(outs[0][1], outs[1][1], what="lnpsi dmma vs fma")
    assert_close(outs[0][2], outs[1][2], what="lnpsi(fixed spins) dmma vs fma")
    m = o.make_ansatz(model, N, M, K)
    m.variables = params.copy()
    lnpsi = m.initialize(spins.astype(np.float64))
    assert_close(outs[0][0], m.y, what="theta vs oracle")
    assert_close(outs[0][1], lnpsi, what="lnpsi vs oracle")


def test_structured_rejects_long_chains_and_conflicting_flags():
    from neural_network_quantum_state_b200 import NQSError
    with pytest.raises(NQSError):
        _engine("rbm", 300, 8, 16, H, J, ALPHA, structured_sv=True)
    with pytest.raises(NQSError):
        _engine("rbm", 16, 8, 16, H, J, ALPHA, structured_sv=True, two_pass_sv=True)
