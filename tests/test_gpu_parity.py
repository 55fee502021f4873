"""GPU parity tests proper: the CUDA engine (through the C ABI, via ctypes) against
  (1) the golden vectors produced by the reference's own CPU code (tests/golden/*.npz), and
  (2) the numpy oracle on seeded inputs at sizes the oracle finishes in seconds.
Bars (north star / SURVEY 8c): fp64 quantities rel 1e-10 (abs 1e-12 near zero); accept/reject decisions EXACT given the same
uniforms; dx after a converged-to-1e-5 CG rel 1e-6 (iteration count must match here).
"""
import math
import os

import numpy as np
import pytest

from helpers import assert_close, audit_accepts, ffnn_cpu_to_gpu_layout
from oracle import nqs_oracle as o

pytestmark = pytest.mark.gpu


def _engine(*a, **kw):
    from neural_network_quantum_state_b200 import Engine
    return Engine(*a, **kw)


def to_gpu(g, v):
    return ffnn_cpu_to_gpu_layout(v, g["N"], g["M"]) if g["model"] == "ffnn" else v


def engine_from_golden(g, **kw):
    e = _engine(g["model"], g["N"], g["M"], g["K"], g["h"], g["J"], g["alpha"], pbc=bool(g["pbc"]), order=g["order"],
                max_predrawn_steps=g["uniforms"].shape[0], accept_log=True, **kw)
    e.set_params(g["params"])
    e.set_uniforms(g["uniforms"])
    return e


@pytest.mark.parametrize("force_generic", [False, True])
def test_golden_sampler_energy_gradients(golden, force_generic):
    g = golden
    e = engine_from_golden(g, force_generic=force_generic)
    e.warm_up(g["n_warm"], g.get("init_spins"))
    assert np.array_equal(e.get_spinStates(), g["warm_spins"]), "accept/reject decisions differ from the reference"
    assert_close(e.get_theta(), g["warm_y"], what="theta")
    assert_close(e.get_lnpsi(), g["warm_lnpsi"], what="lnpsi0")
    for i in range(g["N"]):
        assert_close(e.forward_flip(i), g["flip_lnpsi"][i], what="forward(flip %d)" % i)
    assert_close(e.get_htilda(), g["htilda"], what="htilda")
    O = e.get_lnpsiGradients()
    assert_close(O, to_gpu(g, g["O"]), what="O")
    Sv, aO, diag = e.smatrix_dot(g["sm_lambda"], to_gpu(g, g["sm_v"]))
    assert_close(aO, to_gpu(g, g["sm_aO"]), what="<O>")
    assert_close(diag, to_gpu(g, g["sm_diag"]), what="diag S")
    assert_close(Sv, to_gpu(g, g["sm_Sv"]), what="S v")
    e.close()


@pytest.mark.parametrize("force_generic", [False, True])
def test_golden_sr_trajectory(golden, force_generic):
    g = golden
    e = engine_from_golden(g, force_generic=force_generic)
    e.warm_up(g["n_warm"], g.get("init_spins"))
    for it in range(g["n_sr"]):
        st = e.sr_step(n_mc_steps=1, lr=g["lr"])
        assert st.finite
        assert_close(st.e_mean, g["sr_E"][it], what="<H> it %d" % it)
        assert abs(st.rsd - g["sr_rsd"][it]) < 1e-10
        assert st.lam == pytest.approx(g["sr_lambda"][it], rel=1e-14)
        assert st.cg_iters == int(g["sr_cg_iters"][it])
        F, dx = e.get_sr_vectors()
        assert_close(F, to_gpu(g, g["sr_F"][it]), what="F it %d" % it)
        assert_close(dx, to_gpu(g, g["sr_dx"][it]), rtol=1e-6, what="dx it %d" % it)
    assert_close(e.get_params(), g["final_params"], rtol=1e-8, what="params")
    assert np.array_equal(e.get_spinStates(), g["final_spins"])
    assert_close(e.get_theta(), g["final_y"], rtol=1e-8, what="final theta")
    assert_close(e.get_lnpsi(), g["final_lnpsi"], rtol=1e-8, what="final lnpsi0")
    e.close()


def test_golden_param_files(golden, tmp_path):
    g = golden
    e = _engine(g["model"], g["N"], g["M"], 4, g["h"], g["J"], g["alpha"], sampler_only=True)
    prefix = os.path.join(os.path.dirname(__file__), "golden", "files", g["name"] + "_")
    e.load(prefix)
    assert_close(e.get_params(), g["params"], rtol=2e-10, what="loaded params")
    e.set_params(g["params"])
    out = str(tmp_path / "y_")
    e.save(out, 10)
    sufs = ("Dw.dat", "Da.dat", "Db.dat") if g["model"] == "rbm" else ("Dw1.dat", "Dw2.dat", "Db1.dat")
    for suf in sufs:
        assert open(out + suf).read() == open(prefix + suf).read(), suf
    e.close()


# ---------------------------------------------------------------------------------------------------------------------
# engine vs numpy oracle on seeded inputs (ragged sizes: M not a multiple of 32, odd N, K not a multiple of the CTA size)
# ---------------------------------------------------------------------------------------------------------------------
H, J, ALPHA = -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0
RBM_SWEEP_VARIANTS = ("rbm_f32filter", "rbm_regs")   # fp32-filtered exact sweep (M <= 512) / all-fp64 register-resident sweep
FFNN_SWEEP_VARIANT = "ffnn_resident"     # kernel_variant("sweep") prefix the FNN sampler must report when not forced generic (M <= 512)


def synth(model, N, M, rng, scale=1.0):
    sw = math.sqrt(1.0 / (N + M))
    if model == "rbm":
        W = scale * 0.8 * (rng.normal(0, sw, (N, M)) + 1j * rng.normal(0, sw, (N, M)))
        a = 0.3 * (rng.normal(size=N) + 1j * rng.normal(size=N))
        b = 0.5 * math.sqrt(1.0 / M) * (rng.normal(size=M) + 1j * rng.normal(size=M))
        return np.concatenate([W.ravel(), a, b])
    W = scale * (rng.normal(0, sw, (N, M)) + 0.1j * rng.normal(0, sw, (N, M)))
    b1 = 0.3 * (rng.normal(size=M) + 1j * rng.normal(size=M))
    w1o = rng.normal(0, math.sqrt(1.0 / M), M) + 0.1j * rng.normal(0, math.sqrt(1.0 / M), M)
    return np.concatenate([W.ravel(), b1, w1o])


CASES = [
    ("rbm", 16, 16, 512, False),     # cfg1 shape
    ("rbm", 24, 40, 200, True),      # M not multiple of 32, PBC
    ("rbm", 33, 64, 130, False),     # odd N
    ("rbm", 32, 128, 96, False),
    ("rbm", 20, 256, 64, False),     # M = 256 (cfg3 hidden width)
    ("ffnn", 16, 48, 200, False),
    ("ffnn", 21, 70, 77, False),
    ("rbm", 10, 600, 40, False),     # 512 < M <= 1024: two warps per chain in the sweep (cfg5 width class)
    ("rbm", 12, 1024, 21, False),    # cfg5 hidden width, odd number of chains (a CTA with idle warp pairs)
    ("rbm", 6, 1100, 10, False),     # M > 1024: generic sweep, product-form local energy
    ("rbm", 12, 384, 40, False),     # 256 < M <= 512: rbm_sweep_fast_kernel<16,1,1> (one chain per warp, 16 slots per lane)
    ("rbm", 14, 512, 33, False),     # same kernel at its widest, odd chain count
]
# BASELINE.json shapes (N, M of cfg2 / cfg3 / cfg4 / cfg5) at a chain count the numpy oracle finishes in seconds: these are the
# launch shapes bench.py runs -- the 4-slot TMA ring over 128 sites, the 4-warp rbm_eloc_sites_kernel, spin_rows_dmma MT 7/8,
# the two-warps-per-chain sweep at N = 256, and the FNN sampler at cfg4 width.  (model, N, M, K, n_warm, n_more)
BASELINE_SHAPES = [
    ("rbm", 64, 128, 96, 3, 2),      # cfg2
    ("rbm", 128, 256, 64, 3, 1),     # cfg3 (headline)
    ("rbm", 128, 256, 130, 2, 1),    # cfg3, ragged chain count (partial CTAs of the sweep, E_loc and GEMM tiles)
    ("ffnn", 128, 512, 24, 2, 1),    # cfg4
    ("rbm", 256, 1024, 12, 2, 1),    # cfg5
]


def _sweep_vs_oracle(model, N, M, K, pbc, force_generic, n_warm, n_more, check_O=True):
    rng = np.random.default_rng(N * 1000 + M)
    params = synth(model, N, M, rng)
    U = rng.random(((n_warm + n_more) * N, K))
    m = o.make_ansatz(model, N, M, K)
    m.variables = params.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, pbc, o.UniformSource(K, predrawn=U))
    s.record = True
    e = _engine(model, N, M, K, H, J, ALPHA, pbc=pbc, max_predrawn_steps=U.shape[0], accept_log=True,
                force_generic=force_generic)
    e.set_params(params)
    e.set_uniforms(U)
    s.warm_up(n_warm)
    e.warm_up(n_warm)
    expect = "generic" if (force_generic or M > 1024) else (RBM_SWEEP_VARIANTS if model == "rbm" else FFNN_SWEEP_VARIANT)
    if model == "rbm" and M > 1024 and not force_generic:
        e.get_htilda()
        assert e.kernel_variant("eloc").startswith("rbm_sites"), e.kernel_variant("eloc")
    assert e.kernel_variant("sweep").startswith(expect), e.kernel_variant("sweep")
    # exact accept/reject parity, audited: a mismatch must be a rounding tie |u - ratio| < 1e-12 ratio
    keep = audit_accepts(e.get_accept_log(), np.array(s.accept_log), U[:n_warm * N], s.ratio_log)
    assert keep.all(), "a chain hit a rounding tie: pick another seed for this case"
    assert np.array_equal(e.get_spinStates(), m.spins.astype(np.int8))
    assert_close(e.get_theta(), m.y, what="theta")
    assert_close(e.get_lnpsi(), s.lnpsi0, what="lnpsi0")
    s.accept_log, s.ratio_log = [], []
    s.do_mcmc_steps(n_more)
    e.do_mcmc_steps(n_more)
    keep = audit_accepts(e.get_accept_log(), np.array(s.accept_log), U[n_warm * N:], s.ratio_log)
    assert keep.all()
    assert_close(e.get_lnpsi(), s.lnpsi0, what="lnpsi0 after more sweeps")
    assert_close(e.get_htilda(), s.get_htilda(), what="htilda")
    if check_O:
        assert_close(e.get_lnpsiGradients(), s.get_lnpsiGradients(), what="O")
    return e, s, m


@pytest.mark.parametrize("force_generic", [False, True])
@pytest.mark.parametrize("model,N,M,K,pbc", CASES)
def test_sweep_matches_oracle_step_by_step(model, N, M, K, pbc, force_generic):
    e, _, _ = _sweep_vs_oracle(model, N, M, K, pbc, force_generic, 5, 3)
    e.close()


# chains per warp of the register-resident sweep: small shards take one chain per warp (a single wave), big ones the throughput
# choice; NQS_SWEEP_C pins it so that every instantiation meets the oracle whatever the shard size of the test
@pytest.mark.parametrize("N,M,K,C", [(16, 16, 70, 4), (16, 16, 70, 2), (16, 16, 70, 1), (24, 40, 50, 4), (24, 40, 50, 2), (24, 40, 50, 1),
                                     (20, 128, 45, 4), (20, 128, 45, 2), (20, 128, 45, 1), (20, 256, 33, 2), (20, 256, 33, 1),
                                     (128, 256, 40, 2), (128, 256, 40, 1)])
def test_sweep_chains_per_warp_variants(monkeypatch, N, M, K, C):
    monkeypatch.setenv("NQS_SWEEP_F32", "0")           # the all-fp64 register-resident kernels (fast_kernels.cuh)
    monkeypatch.setenv("NQS_SWEEP_C", str(C))
    e, _, _ = _sweep_vs_oracle("rbm", N, M, K, False, False, 3, 1, check_O=False)
    assert e.kernel_variant("sweep").endswith("_c%d" % C), e.kernel_variant("sweep")
    e.close()


@pytest.mark.parametrize("delta", ["1", "1e7"])
@pytest.mark.parametrize("N,M,K", [(16, 16, 70), (24, 40, 50), (20, 128, 45), (20, 256, 33), (12, 384, 40), (14, 512, 33), (128, 256, 40)])
def test_f32_filtered_sweep_is_exact(monkeypatch, N, M, K, delta):
    """sweep_f32.cuh: the fp32 product only DECIDES when u R0 lies outside its rigorous error interval, otherwise the fp64 product
    does -- accept masks must equal the oracle's whatever the share of each path: "1" = the real bound (almost everything fp32),
    "1e7" = an interval wider than any ratio (every proposal takes the fp64 path)."""
    monkeypatch.setenv("NQS_SWEEP_F32", "1")
    monkeypatch.setenv("NQS_SWEEP_F32_DELTA", delta)
    e, _, _ = _sweep_vs_oracle("rbm", N, M, K, False, False, 3, 1, check_O=False)
    if N < 100:      # (with these weights the overflow bound of 128 sites exceeds the fp32 kernel's range guard: the fp64 kernel runs)
        assert e.kernel_variant("sweep").startswith("rbm_f32filter"), e.kernel_variant("sweep")
    e.close()


def test_f32_filter_agrees_bitwise_with_the_fp64_sweep(monkeypatch):
    """Same seeds through the fp32-filtered kernel and the all-fp64 register kernel: spins and theta bit-identical after many sweeps
    (the decisions are the same; theta is replayed exactly in both)."""
    from neural_network_quantum_state_b200 import Engine
    model, N, M, K = "rbm", 64, 128, 700
    params = synth(model, N, M, np.random.default_rng(9), scale=1.5)   # (larger weights trip the fp32 kernel's range guard)
    outs = []
    for f32 in ("1", "0"):
        monkeypatch.setenv("NQS_SWEEP_F32", f32)
        e = Engine(model, N, M, K, H, J, ALPHA, seed=99)
        e.set_params(params)
        e.warm_up(20)
        e.do_mcmc_steps(5)
        outs.append((e.get_spinStates(), e.get_theta(), e.get_lnpsi(), e.kernel_variant("sweep")))
        e.close()
    assert outs[0][3].startswith("rbm_f32filter") and outs[1][3].startswith("rbm_regs")
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert_close(outs[0][2], outs[1][2], what="lnpsi0")


@pytest.mark.parametrize("model,N,M,K,n_warm,n_more", BASELINE_SHAPES)
def test_baseline_shapes_match_oracle(model, N, M, K, n_warm, n_more):
    """Sampler, local energy, O, SR sums and S*v at the (N, M) of BASELINE.json's configurations, against the numpy oracle."""
    e, s, m = _sweep_vs_oracle(model, N, M, K, False, False, n_warm, n_more)
    # the launch shapes the benchmark runs, not a fallback
    if model == "rbm":
        assert e.kernel_variant("eloc").startswith("rbm_sites"), e.kernel_variant("eloc")
    assert e.kernel_variant("theta") == "dmma_rows", e.kernel_variant("theta")
    O = s.get_lnpsiGradients()
    rng = np.random.default_rng(5)
    v = rng.normal(size=m.P) + 1j * rng.normal(size=m.P)
    S = o.SMatrix(O, 0.37)
    Sv, aO, diag = e.smatrix_dot(0.37, v)
    assert_close(aO, S.aO, what="<O>")
    assert_close(diag, S.diag, rtol=1e-9, what="diag S")          # difference of two O(1) means
    assert_close(Sv, S.dot(v), rtol=1e-9, what="S v")
    e.close()


@pytest.mark.parametrize("model,N,M,K,pbc", CASES[:3] + CASES[5:6])
def test_internal_philox_stream_matches_oracle(model, N, M, K, pbc):
    rng = np.random.default_rng(7)
    params = synth(model, N, M, rng)
    seed, koff = 0x1234567890ABCDEF, 1000
    m = o.make_ansatz(model, N, M, K)
    m.variables = params.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, pbc, o.UniformSource(K, seed=seed, chain_offset=koff))
    e = _engine(model, N, M, K, H, J, ALPHA, pbc=pbc, seed=seed, chain_offset=koff, n_chains_total=K + koff)
    e.set_params(params)
    s.warm_up(4)
    e.warm_up(4)
    s.do_mcmc_steps(2)
    e.do_mcmc_steps(2)
    assert np.array_equal(e.get_spinStates(), m.spins.astype(np.int8))
    assert_close(e.get_lnpsi(), s.lnpsi0, what="lnpsi0")
    e.close()


@pytest.mark.parametrize("model,N,M,K,pbc", [CASES[1], CASES[3], CASES[6]])
def test_sr_matches_oracle_fixed_cg_iterations(model, N, M, K, pbc):
    """Fixed CG iteration count on both sides -> dx agrees to 1e-10-ish (no dependence on the stopping rule)."""
    rng = np.random.default_rng(11)
    params = synth(model, N, M, rng)
    U = rng.random((12 * N, K))
    m = o.make_ansatz(model, N, M, K)
    m.variables = params.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, pbc, o.UniformSource(K, predrawn=U))
    e = _engine(model, N, M, K, H, J, ALPHA, pbc=pbc, max_predrawn_steps=U.shape[0])
    e.set_params(params)
    e.set_uniforms(U)
    s.warm_up(8)
    e.warm_up(8)
    sr = o.StochasticReconfigurationCG(K, m.P)
    for it in range(3):
        st_o = sr.step(s, 1, 0.03, fixed_cg_iters=7, lam=0.5)
        st = e.sr_step(n_mc_steps=1, lr=0.03, fixed_iters=7, lam=0.5)
        assert st.cg_iters == 7
        assert_close(st.e_mean, st_o.e_mean, what="<H>")
        F, dx = e.get_sr_vectors()
        assert_close(F, st_o.F, what="F")
        assert_close(dx, st_o.dx, rtol=1e-9, what="dx (7 CG its)")
    assert_close(e.get_params(), m.variables, rtol=1e-9, what="params")
    e.close()


def test_large_theta_falls_back_to_generic_kernels():
    """The product-form kernels are only used while they cannot overflow; a network with huge |Re theta| must take the
    generic path and still match the oracle."""
    model, N, M, K = "rbm", 12, 32, 40
    rng = np.random.default_rng(21)
    params = synth(model, N, M, rng)
    params[:N * M] = params[:N * M].real * 600.0 + 1j * params[:N * M].imag
    U = rng.random((4 * N, K))
    m = o.make_ansatz(model, N, M, K)
    m.variables = params.copy()
    s = o.LITFIChainSampler(m, H, J, ALPHA, False, o.UniformSource(K, predrawn=U))
    e = _engine(model, N, M, K, H, J, ALPHA, max_predrawn_steps=U.shape[0])
    e.set_params(params)
    e.set_uniforms(U)
    s.warm_up(3)
    e.warm_up(3)
    assert e.kernel_variant("sweep") == "generic"
    assert np.array_equal(e.get_spinStates(), m.spins.astype(np.int8))
    assert_close(e.get_lnpsi(), s.lnpsi0, what="lnpsi0")
    assert_close(e.get_htilda(), s.get_htilda(), what="htilda")
    e.close()


def test_fast_and_generic_kernels_agree_bitwise_on_theta():
    """The specialised sweep replays accepted flips on the exact theta in the reference order: theta and spins must be
    bit-identical to the generic kernel's; lnpsi0 / htilda agree to rounding."""
    model, N, M, K = "rbm", 32, 256, 70
    rng = np.random.default_rng(9)
    params = synth(model, N, M, rng)
    outs = []
    for fg in (False, True):
        e = _engine(model, N, M, K, H, J, ALPHA, seed=99, force_generic=fg)
        e.set_params(params)
        e.warm_up(7)
        e.do_mcmc_steps(3)
        outs.append((e.get_spinStates(), e.get_theta(), e.get_lnpsi(), e.get_htilda(), e.kernel_variant("sweep")))
        e.close()
    assert outs[0][4].startswith(RBM_SWEEP_VARIANTS) and outs[1][4] == "generic"
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert_close(outs[0][2], outs[1][2], what="lnpsi0 fast vs generic")
    assert_close(outs[0][3], outs[1][3], what="htilda fast vs generic")


def test_identities_flip_and_rank1_update():
    """oracle-independent: forward(i) == full forward of the flipped configuration; rank-1 updated theta == recomputed."""
    model, N, M, K = "rbm", 18, 50, 64
    rng = np.random.default_rng(3)
    params = synth(model, N, M, rng)
    params[N * M:N * M + N] = 0.0  # a = 0 so that the forward(spins) visible-bias quirk does not enter
    e = _engine(model, N, M, K, H, J, ALPHA, seed=5)
    e.set_params(params)
    e.warm_up(6)
    spins = e.get_spinStates()
    for i in (0, 7, N - 1):
        flipped = spins.copy()
        flipped[:, i] *= -1
        assert_close(e.forward_flip(i), e.get_lnpsi_for_fixed_spins(flipped), what="flip identity %d" % i)
    th = e.get_theta()
    e.evolve(np.zeros(e.P, dtype=np.complex128), 0.0)  # re-derives theta from the spins
    assert_close(e.get_theta(), th, what="rank-1 theta vs recomputed")
    e.close()


def test_error_behaviour():
    from neural_network_quantum_state_b200 import NQSError
    from neural_network_quantum_state_b200 import _lib as L
    e = _engine("rbm", 8, 8, 16, H, J, ALPHA, max_predrawn_steps=8)
    with pytest.raises(NQSError) as ei:
        e.do_mcmc_steps(1)
    assert ei.value.status == L.ERR_STATE
    e.set_uniforms(np.random.default_rng(0).random((8, 16)))
    e.warm_up(1)
    with pytest.raises(NQSError) as ei:
        e.do_mcmc_steps(1)          # feed exhausted
    assert ei.value.status == L.ERR_STATE
    with pytest.raises(NQSError):
        e.set_params(np.zeros(3, dtype=np.complex128))
    with pytest.raises(NQSError):
        _engine("rbm", 7, 8, 16, H, J, ALPHA, pbc=True)   # odd L with PBC (ref: invalid_argument)
    e.close()


def test_zero_chains_edge_and_single_chain():
    from neural_network_quantum_state_b200 import NQSError
    with pytest.raises(NQSError):
        _engine("rbm", 8, 8, 0, H, J, ALPHA)
    e = _engine("rbm", 5, 3, 1, H, J, ALPHA, seed=1, sampler_only=True)
    e.init_params_random(3)
    e.warm_up(3)
    assert e.get_spinStates().shape == (1, 5)
    e.close()


@pytest.mark.parametrize("model,N,M,K", [("rbm", 128, 256, 130), ("rbm", 24, 40, 200), ("ffnn", 16, 48, 77), ("rbm", 12, 1024, 21),
                                          ("rbm", 6, 1100, 10)])
def test_pinned_uniforms_are_read_in_place_and_give_the_same_chain(model, N, M, K, monkeypatch):
    """nqs_set_uniforms with a page-locked buffer: the sweep kernels read it over PCIe (no staging copy, a group of proposals
    fetched ahead); pageable memory is staged through HBM.  Same uniforms -> the same accept decisions, bit for bit -- including
    when the feed is longer than max_predrawn_steps (no device buffer is involved) and when a sweep starts in the middle of it."""
    import torch
    rng = np.random.default_rng(N + M)
    n_sweeps = 5
    U = rng.random((n_sweeps * N, K))
    params = synth(model, N, M, np.random.default_rng(3))
    res = []
    for mode in ("pageable", "pinned", "pinned_copy"):
        monkeypatch.setenv("NQS_UNIFORMS_ZEROCOPY", "0" if mode == "pinned_copy" else "1")
        # the pinned run gets a device staging buffer far too small for the feed: it must not need it
        e = _engine(model, N, M, K, H, J, ALPHA, max_predrawn_steps=(2 if mode == "pinned" else U.shape[0]), accept_log=True,
                    sampler_only=True)
        e.set_params(params)
        if mode == "pageable":
            buf = U.copy()
        else:
            t = torch.empty(U.shape, dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = U
            buf = t.numpy()
        e.set_uniforms(buf)
        e.warm_up(2)
        e.do_mcmc_steps(1)
        e.do_mcmc_steps(2)
        res.append((e.get_spinStates(), e.get_theta(), e.get_lnpsi()))
        e.set_uniforms(None)
        e.close()
    for r in res[1:]:
        assert np.array_equal(res[0][0], r[0]), "accept decisions differ between the staged and the in-place uniform feed"
        assert np.array_equal(res[0][1], r[1])
        assert np.array_equal(res[0][2], r[2])
