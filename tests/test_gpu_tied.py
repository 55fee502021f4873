"""The two other tied-variable ansaetze the reference drives on the long-range chain -- RBMZ2PrSymm (ref
gpu/include/impl_neural_quantum_state.cuh:540-745, gpu/src/LICH-train_rbmz2prsymm.cu, OPEN chain) and FFNNTrSymm (:1019-1223,
gpu/src/LICH-train_ffnntrsymm.cu, PERIODIC chain): the CUDA engine against the numpy oracle, and the command-line programs
against the reference's own CUDA drivers compiled for sm_100 (baseline/Makefile, *-ref-yarn2)."""
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import assert_close, audit_accepts
from oracle import nqs_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_cuda  # noqa: E402

pytestmark = pytest.mark.gpu
H, J, ALPHA = -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0
KINDS = {"rbmz2prsymm": (o.RBMZ2PrSymm, lambda N, al: 4 * al, False), "ffnntrsymm": (o.FFNNTrSymm, lambda N, al: al * N, True)}


def _vars(kind, N, al, rng, scale):
    t = KINDS[kind][0](N, al, 1, rng)
    v = t.variables * scale
    if kind == "ffnntrsymm":
        v[N * al: N * al + al] = 0.1 * (rng.normal(size=al) + 1j * rng.normal(size=al))      # non-zero hidden biases
    return v


@pytest.mark.parametrize("kind,N,al,K,scale", [("rbmz2prsymm", 16, 2, 130, 5.0), ("rbmz2prsymm", 33, 3, 77, 5.0), ("rbmz2prsymm", 64, 16, 64, 4.0),
                                                ("rbmz2prsymm", 128, 5, 40, 4.0),
                                                ("ffnntrsymm", 16, 2, 130, 1.0), ("ffnntrsymm", 24, 1, 64, 1.5), ("ffnntrsymm", 32, 3, 77, 1.0),
                                                ("ffnntrsymm", 64, 2, 48, 1.0)])
@pytest.mark.parametrize("force_generic", [False, True])
def test_sampler_energy_gradients_match_oracle(kind, N, al, K, scale, force_generic):
    from neural_network_quantum_state_b200 import Engine
    cls, width, pbc = KINDS[kind]
    rng = np.random.default_rng(100 * N + al)
    v = _vars(kind, N, al, rng, scale)
    n_warm, n_more = 4, 2
    U = rng.random(((n_warm + n_more) * N, K))
    m = cls(N, al, K)
    m.variables = v.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, pbc, o.UniformSource(K, predrawn=U))
    s.record = True
    e = Engine(kind, N, width(N, al), K, H, J, ALPHA, pbc=pbc, max_predrawn_steps=U.shape[0], accept_log=True, force_generic=force_generic)
    assert e.P == m.P
    e.set_params(v)
    assert np.array_equal(e.get_params(), v)
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
    acc = e.get_accept_log()
    assert audit_accepts(acc, np.array(s.accept_log), U[n_warm * N:], s.ratio_log).all()
    assert 0.01 < acc.mean() < 0.99
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
    mm = cls(N, al, K)
    mm.variables = v.copy()
    assert_close(e.get_lnpsi_for_fixed_spins(spins), mm.forward_spins(spins, save=False), what="forward(spins)")
    e.close()


@pytest.mark.parametrize("kind,N,al,K", [("rbmz2prsymm", 16, 2, 400), ("rbmz2prsymm", 32, 4, 300), ("ffnntrsymm", 16, 2, 400), ("ffnntrsymm", 32, 2, 300)])
def test_sr_trajectory_matches_oracle(kind, N, al, K):
    from neural_network_quantum_state_b200 import Engine
    cls, width, pbc = KINDS[kind]
    rng = np.random.default_rng(9)
    v = _vars(kind, N, al, rng, 1.0)
    U = rng.random((14 * N, K))
    spins0 = (2 * rng.integers(0, 2, size=(K, N)) - 1).astype(np.float64)     # random start: no zero-variance columns
    m = cls(N, al, K)
    m.variables = v.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, pbc, o.UniformSource(K, predrawn=U))
    e = Engine(kind, N, width(N, al), K, H, J, ALPHA, pbc=pbc, max_predrawn_steps=U.shape[0])
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


@pytest.mark.parametrize("kind,symm", [("rbmz2prsymm", "z2pr"), ("ffnntrsymm", None)])
def test_variables_file_init_law_and_pynqs(tmp_path, kind, symm):
    from neural_network_quantum_state_b200 import Engine
    from neural_network_quantum_state_b200.pynqs import sampler as pysampler, _pynqs_gpu
    cls, width, pbc = KINDS[kind]
    N, al, K = 12, 2, 32
    rng = np.random.default_rng(4)
    v = _vars(kind, N, al, rng, 3.0)
    m = cls(N, al, K)
    m.variables = v.copy()
    path = str(tmp_path / "vars")
    m.save(path, 17)
    e = Engine(kind, N, width(N, al), 4, H, J, ALPHA, pbc=pbc, sampler_only=True)
    e.load(path)
    assert np.array_equal(e.get_params(), v)
    out = str(tmp_path / "out")
    e.save(out, 17)
    assert open(out).read() == open(path).read()
    # the constructor's law (ref :562-579 / :1041-1061): scales of the blocks
    big = Engine(kind, 64, width(64, 8), 4, H, J, ALPHA, pbc=pbc, sampler_only=True)
    big.init_params_random(3)
    p = big.get_params()
    w = p[: 64 * 8]
    if kind == "rbmz2prsymm":
        assert abs(w.real.std() / (0.1 * math.sqrt(1.0 / (32 + 64))) - 1) < 0.15 and abs(w.imag.std() / w.real.std() - 1) < 0.15
    else:
        assert abs(w.real.std() / math.sqrt(1.0 / (9 * 64)) - 1) < 0.15 and abs(w.imag.std() / (0.1 * w.real.std()) - 1) < 0.15
        assert np.all(p[64 * 8: 64 * 8 + 8] == 0)
    big.close()
    e.close()
    if symm is not None:
        r = pysampler.RBM(floatType="float64", symmType=symm)
        r.init(nInputs=N, nHiddens=al, nChains=K, seedNumber=3, seedDistance=1000, path_to_load=path, init_mcmc_steps=5)
        r.do_mcmc_steps(2)
        sp = r.get_spinStates()
        ln, ln_fixed = r.get_lnpsi(), r.get_lnpsi_for_fixed_spins(sp)
    else:                                                    # pynqs.sampler wraps the RBM family only (ref sampler.py:27-40)
        q = _pynqs_gpu.dFFNNTrSymmSampler({"nInputs": N, "nHiddens": al, "nChains": K, "seedNumber": 3, "seedDistance": 1000})
        q.load(path)
        q.warm_up(5)
        q.do_mcmc_steps(2)
        sp = q.get_spinStates().reshape(K, N)
        ln, ln_fixed = q.get_lnpsi(), q.get_lnpsi_for_fixed_spins(sp)
    assert sp.shape == (K, N) and set(np.unique(sp)) <= {-1.0, 1.0}
    mm = cls(N, al, K)
    mm.variables = v.copy()
    assert_close(ln, mm.forward_spins(sp), what="pynqs get_lnpsi")
    assert_close(ln_fixed, mm.forward_spins(sp), what="pynqs get_lnpsi_for_fixed_spins")


@pytest.mark.parametrize("driver,kind,nf", [("rbmz2prsymm", "rbmz2prsymm", 3), ("ffnntrsymm", "ffnntrsymm", 2)])
def test_cli_driver_prints_the_reference_drivers_table(tmp_path, driver, kind, nf):
    """Same -option=value list, same variables file and the same -seed to the reference's own program (gpu/src/LICH-train_<driver>.cu
    for sm_100, TRNG4 -> the restated yarn2 of baseline/shim_yarn2) and to bin/LICH-train_<driver>-gpu: same iteration table, same
    variables written at the end."""
    ref_bin = os.path.join(os.path.dirname(ref_cuda.BINARY), "LICH-train_%s-gpu-ref-yarn2" % driver)
    if not os.path.exists(ref_bin):
        pytest.skip("baseline/_ref/LICH-train_%s-gpu-ref-yarn2 not built" % driver)
    from neural_network_quantum_state_b200 import Engine, build
    cls, width, pbc = KINDS[kind]
    L, ns, niter, nwarm = 16, 512, 5, 40
    tag = {"rbmz2prsymm": "RBMZ2PrSymmLICH", "ffnntrsymm": "FFNNTrSymmLICH"}[driver]
    dirs = [tmp_path / "ref", tmp_path / "ours"]
    e = Engine(kind, L, width(L, nf), 4, 0.0, 0.0, 0.0, sampler_only=True)
    e.init_params_random(8)
    names = []
    for d in dirs:
        d.mkdir()
        names.append(os.path.join(str(d), "%s-L%dNF%dA2T%sV0" % (tag, L, nf, ref_cuda.THETA_STR)))
        e.save(names[-1], 17)
    e.close()
    args = ["-L=%d" % L, "-nf=%d" % nf, "-ns=%d" % ns, "-niter=%d" % niter, "-alpha=2", "-theta=%s" % ref_cuda.THETA_STR,
            "-ver=0", "-nwarm=%d" % nwarm, "-dev=0", "-lr=0.02", "-rsd=1e-30", "-seed=99"]
    our_bin = os.path.join(build.BIN_DIR, "LICH-train_%s-gpu" % driver)
    env = {k: val for k, val in os.environ.items() if k != "NQS_RNG"}
    outs = []
    for b, d in zip((ref_bin, our_bin), dirs):
        r = subprocess.run([b] + args + ["-path=%s" % d], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        rows = [ln.split() for ln in r.stdout.splitlines() if re.match(r"^\s*\d+\s+\S+\s+\S+\s*$", ln)]
        outs.append([(int(a), float(b_), float(c)) for a, b_, c in rows])
        assert "# of loop\t<H>" in r.stdout and "# elapsed time:" in r.stdout
    assert len(outs[0]) == niter and len(outs[1]) == niter
    for (n0, e0, r0), (n1, e1, r1) in zip(*outs):
        assert n0 == n1
        assert e1 == pytest.approx(e0, rel=3e-6, abs=2e-7)
        assert r1 == pytest.approx(r0, rel=3e-5, abs=2e-7)
    want, have = ref_cuda.load_vars(names[0]), ref_cuda.load_vars(names[1])
    assert want.size == have.size == L * nf + (1 if driver == "rbmz2prsymm" else 2) * nf
    assert np.abs(have - want).max() <= 3e-9 * max(1.0, np.abs(want).max())


# ---- against the reference's own CPU code (cpu/include/neural_quantum_state.hpp:68-102, 184-217 compiled in place; vectors by
# `python tests/golden/make_golden.py --tied`): RBMTrSymm and FFNNTrSymm on the periodic chain ------------------------------------
@pytest.mark.parametrize("force_generic", [False, True])
def test_golden_tied_sampler_energy_gradients(golden_tied, force_generic):
    import test_gpu_parity as tp
    tp.test_golden_sampler_energy_gradients(golden_tied, force_generic)


@pytest.mark.parametrize("force_generic", [False, True])
def test_golden_tied_sr_trajectory(golden_tied, force_generic):
    import test_gpu_parity as tp
    tp.test_golden_sr_trajectory(golden_tied, force_generic)


def test_golden_tied_variables_file(golden_tied, tmp_path):
    from neural_network_quantum_state_b200 import Engine
    g = golden_tied
    e = Engine(g["model"], g["N"], g["M"], 4, g["h"], g["J"], g["alpha"], pbc=True, sampler_only=True)
    want = os.path.join(os.path.dirname(__file__), "golden", "files", g["name"] + "_")
    e.load(want)
    assert_close(e.get_params(), g["params"], rtol=2e-10, what="loaded variables")
    e.set_params(g["params"])
    out = str(tmp_path / "vars")
    e.save(out, 10)
    assert open(out).read() == open(want).read()
    e.close()


def test_structured_sv_environment_switch_does_not_reach_tied_handles(tmp_path, monkeypatch):
    """NQS_STRUCTURED_SV=1 (the opt-in of callers that cannot pass flags) is for plain RBM / FNN handles; a tied-variable handle made
    under it -- Python host or command-line program -- must run its explicit-O path, not fail."""
    from neural_network_quantum_state_b200 import Engine, build
    monkeypatch.setenv("NQS_STRUCTURED_SV", "1")
    e = Engine("rbmz2prsymm", 12, 8, 128, H, J, ALPHA, seed=3)
    e.init_params_random(4)
    e.warm_up(5)
    st = e.sr_step(n_mc_steps=1, lr=0.02)
    assert st.finite and st.cg_iters >= 1
    e.close()
    assert os.environ["NQS_STRUCTURED_SV"] == "1"
    args = ["-L=12", "-nf=2", "-ns=256", "-niter=3", "-alpha=2", "-theta=0.785398", "-ver=0", "-nwarm=10", "-dev=0", "-lr=0.02",
            "-rsd=1e-30", "-seed=5", "-path=%s" % tmp_path]
    exe = os.path.join(build.BIN_DIR, "LICH-train_ffnntrsymm-gpu")
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len([ln for ln in r.stdout.splitlines() if re.match(r"^\s*\d+\s+\S+\s+\S+\s*$", ln)]) == 3
