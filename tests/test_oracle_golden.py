"""Pins oracle/nqs_oracle.py against vectors produced by the reference's own CPU code (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import nqs_oracle as o
from helpers import assert_close, ffnn_cpu_to_gpu_layout


def build(g):
    m = o.make_ansatz(g["model"], g["N"], g["M"], g["K"])
    m.variables = g["params"].copy()
    us = o.UniformSource(g["K"], predrawn=g["uniforms"])
    order = o.sequential_order(g["N"]) if g["order"] == "sequential" else None
    s = o.LITFIChainSampler(m, g["h"], g["J"], g["alpha"], bool(g["pbc"]), us, order)
    return m, s


def to_gpu(g, v):
    return ffnn_cpu_to_gpu_layout(v, g["N"], g["M"]) if g["model"] == "ffnn" else v


def test_sweep_energy_gradients(golden):
    g = golden
    m, s = build(g)
    s.warm_up(g["n_warm"], g["init_spins"].astype(np.float64) if "init_spins" in g else None)
    assert np.array_equal(m.spins.astype(np.int8), g["warm_spins"])       # accept/reject decisions: exact
    assert_close(m.y, g["warm_y"], what="theta")
    assert_close(s.lnpsi0, g["warm_lnpsi"], what="lnpsi0")
    for i in range(g["N"]):
        assert_close(m.forward_flip(i), g["flip_lnpsi"][i], what="flip lnpsi %d" % i)
    assert_close(s.get_htilda(), g["htilda"], what="htilda")
    O = s.get_lnpsiGradients()
    assert_close(O, to_gpu(g, g["O"]), what="O")
    S = o.SMatrix(O, g["sm_lambda"])
    assert_close(S.aO, to_gpu(g, g["sm_aO"]), what="<O>")
    assert_close(S.diag, to_gpu(g, g["sm_diag"]), what="diag S")
    assert_close(S.dot(to_gpu(g, g["sm_v"])), to_gpu(g, g["sm_Sv"]), what="S v")


def test_sr_trajectory(golden):
    g = golden
    m, s = build(g)
    s.warm_up(g["n_warm"], g["init_spins"].astype(np.float64) if "init_spins" in g else None)
    sr = o.StochasticReconfigurationCG(g["K"], m.P)
    for it in range(g["n_sr"]):
        st = sr.step(s, 1, g["lr"])
        assert_close(st.e_mean, g["sr_E"][it], what="<H> it %d" % it)
        assert abs(st.rsd - g["sr_rsd"][it]) < 1e-10
        assert st.lam == pytest.approx(g["sr_lambda"][it], rel=1e-14)
        assert st.cg_iters == int(g["sr_cg_iters"][it])
        assert_close(st.F, to_gpu(g, g["sr_F"][it]), what="F it %d" % it)
        # dx is the output of a converged-to-1e-5 PCG: agreement is limited by conditioning, not by the tolerance
        assert_close(st.dx, to_gpu(g, g["sr_dx"][it]), rtol=1e-8, what="dx it %d" % it)
    assert_close(m.variables, g["final_params"], what="params")
    assert np.array_equal(m.spins.astype(np.int8), g["final_spins"])
    assert_close(m.y, g["final_y"], what="final theta")
    assert_close(s.lnpsi0, g["final_lnpsi"], what="final lnpsi0")


def test_param_files_roundtrip(golden, tmp_path, capsys):
    """Files written by the reference's save() (precision 10) are read back; files written by the oracle are
    byte-identical to the reference's."""
    g = golden
    m = o.make_ansatz(g["model"], g["N"], g["M"], 1)
    prefix = os.path.join(os.path.dirname(__file__), "golden", "files", g["name"] + "_")
    m.load(prefix)
    assert_close(m.variables, g["params"], rtol=2e-10, atol=1e-12, what="loaded params")
    m.variables = g["params"].copy()
    out = str(tmp_path / "x_")
    m.save(out, 10)
    sufs = ("Dw.dat", "Da.dat", "Db.dat") if g["model"] == "rbm" else ("Dw1.dat", "Dw2.dat", "Db1.dat")
    for suf in sufs:
        assert open(out + suf).read() == open(prefix + suf).read(), suf


def test_exact_diagonalisation_known_answers():
    import math
    J, h = math.sin(math.pi / 4), -math.cos(math.pi / 4)
    assert o.exact_ground_energy_per_site(8, J, h, 2.0) == pytest.approx(-0.837475552251, abs=1e-10)
    assert o.exact_ground_energy_per_site(12, J, h, 2.0) == pytest.approx(-0.842986271888, abs=1e-10)


def test_structured_factorisation_of_S_dot_v(golden):
    """The algebra the tensor-core kernels of csrc/sv_struct.cuh rely on, checked on the CPU against the reference's golden S v:
    every row of O is an outer product plus two short blocks, so O v and O^H z (and <O>, diag S) follow from the [K][N] spins
    and the [K][M] hidden-unit factors without O.  RBM: O_k = [s_ki T_kj (i*M+j) | s_ki | T_kj], T = tanh(theta);
    FFNN (GPU layout): O_k = [s_ki T'_kj (j*N+i) | T'_kj | L_kj], T' = tanh(theta) w1o, L = logcosh(theta)."""
    g = golden
    m, s = build(g)
    s.warm_up(g["n_warm"], g["init_spins"].astype(np.float64) if "init_spins" in g else None)
    N, M, K = g["N"], g["M"], g["K"]
    S, th = m.spins, m.y
    v = to_gpu(g, g["sm_v"])
    lam = g["sm_lambda"]
    if g["model"] == "rbm":
        T = np.tanh(th)
        V, va, vb = v[:N * M].reshape(N, M), v[N * M:N * M + N], v[N * M + N:]
        z = np.einsum("kj,kj->k", T, S @ V + vb[None, :]) + S @ va
        C = T.conj() * z[:, None]
        OHz = np.concatenate([(S.T @ C).ravel(), S.T @ z, C.sum(axis=0)])
        aO = np.concatenate([(S.T @ T).ravel(), S.sum(axis=0).astype(np.complex128), T.sum(axis=0)]) / K
        t2 = (np.abs(T) ** 2).sum(axis=0) / K
        m2 = np.concatenate([np.tile(t2, N), np.ones(N), t2])
    else:
        T = np.tanh(th) * m.w1o[None, :]
        L = o.logcosh(th)
        V, vb, vl = v[:N * M].reshape(M, N).T, v[N * M:N * M + M], v[N * M + M:]      # W block stored j*N+i
        z = np.einsum("kj,kj->k", T, S @ V + vb[None, :]) + L @ vl
        C = T.conj() * z[:, None]
        OHz = np.concatenate([(S.T @ C).T.ravel(), C.sum(axis=0), (L.conj() * z[:, None]).sum(axis=0)])
        aO = np.concatenate([(S.T @ T).T.ravel(), T.sum(axis=0), L.sum(axis=0)]) / K
        t2, l2 = (np.abs(T) ** 2).sum(axis=0) / K, (np.abs(L) ** 2).sum(axis=0) / K
        m2 = np.concatenate([np.repeat(t2, N), t2, l2])
    diag = m2 - np.abs(aO) ** 2
    Sv = OHz / K - aO.conj() * (aO @ v) + lam * diag * v
    assert_close(aO, to_gpu(g, g["sm_aO"]), what="<O> from the factors")
    assert_close(diag, to_gpu(g, g["sm_diag"]), atol=1e-11, what="diag S from the factors")
    assert_close(Sv, to_gpu(g, g["sm_Sv"]), what="S v from the factors")


# ---- the tied-variable ansaetze of the reference's CPU tree (cpu/include/neural_quantum_state.hpp:68-102, 184-217; vectors by
# `python tests/golden/make_golden.py --tied`): pins oracle.RBMTrSymm / oracle.FFNNTrSymm to the reference's own code --------------
def test_tied_sweep_energy_gradients(golden_tied):
    test_sweep_energy_gradients(golden_tied)


def test_tied_sr_trajectory(golden_tied):
    test_sr_trajectory(golden_tied)


def test_tied_variables_file_is_byte_identical(golden_tied, tmp_path):
    g = golden_tied
    m = o.make_ansatz(g["model"], g["N"], g["M"], g["K"])
    m.variables = g["params"].copy()
    out = str(tmp_path / "vars")
    m.save(out, 10)
    want = os.path.join(os.path.dirname(__file__), "golden", "files", g["name"] + "_")
    assert open(out).read() == open(want).read()
    m2 = o.make_ansatz(g["model"], g["N"], g["M"], g["K"])
    m2.load(want)
    assert_close(m2.variables, g["params"], rtol=2e-10, what="10-digit file")
