"""One-pass S*v cluster kernel (csrc/sv_fused.cuh) against (1) the dense formula evaluated in numpy on the engine's own O
and (2) the two-pass kernels, over shapes that exercise every launch plan: cluster sizes 8 / 9 / 10 / 16, 1..9 columns per thread,
ragged last column slice, row counts that do not divide by the number of clusters, fewer rows than clusters.

ref: SMatrixForCG::dot, gpu/include/functor_for_CG.cuh:104-127:  S v = (1/K) O^H (O v) - conj(<O>) (<O> . v) + lambda diag(S) v.
"""
import math

import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu

H, J, ALPHA = -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0

# (model, N, M, K, pinned cluster size or None = the engine's own choice, expected cluster size, expected columns per thread)
SHAPES = [
    ("rbm", 16, 16, 64, 8, 8, 1),       # tiny: P = 288, 36 columns per CTA
    ("rbm", 6, 1, 40, 8, 8, 1),         # P = 13 < 2*8: trailing CTAs own no column at all
    ("rbm", 64, 128, 300, 8, 8, 8),     # cfg2 shape, K not a multiple of the cluster count
    ("rbm", 80, 250, 37, 8, 8, 8),      # P = 20330
    ("rbm", 100, 260, 50, 8, 8, 8),     # P = 26360
    ("rbm", 128, 256, 150, 8, 8, 9),    # cfg3 shape (P = 33152), portable cluster size
    ("rbm", 128, 256, 150, None, 9, 8), # cfg3 shape as shipped: 9-CTA clusters cover 135 SMs
    ("rbm", 64, 128, 300, None, 16, 5), # cfg2 shape as shipped
    ("rbm", 128, 256, 150, 10, 10, 8),
    ("ffnn", 128, 512, 40, None, 16, 9),# cfg4 shape (P = 66560): 16-CTA clusters
    ("rbm", 128, 256, 5, None, 16, 8),  # fewer rows than clusters: the widest cluster covers most SMs
    ("rbm", 24, 40, 33, 8, 8, 1),       # P = 1024: 128 columns per CTA, exactly the 128-thread floor of the fat-warp rule
    ("rbm", 64, 128, 300, 2, 2, 9),     # narrow rows on SMALL clusters (the cfg2 plan): 2 CTAs x 4192 columns
    ("rbm", 24, 40, 130, 1, 1, 8),      # a cluster of ONE CTA: the whole row in one slice, the DSMEM exchange targets itself
    ("rbm", 32, 64, 90, 3, 3, 7),
    ("ffnn", 16, 48, 77, 4, 4, 2),
]


def planned_cpt(P, cs):
    """The engine's rule (engine.cu: plan_sv): few fat warps -- 8 columns per thread down to 1 while at least 128 consumer
    threads remain, then 9 and 10 (wide slices); slices too narrow for 128 threads take the fewest columns that fit."""
    pc = -(-P // cs)

    def need(c):
        return (-(-pc // c) + 31) // 32 * 32

    for c in (8, 7, 6, 5, 4, 3, 2, 1, 9, 10):
        if 128 <= need(c) <= (992 if c <= 3 else 480):
            return c
    for c in range(1, 11):
        if need(c) <= (992 if c <= 3 else 480):
            return c
    return None


def dense_sv(O, v, lam):
    K = O.shape[0]
    aO = O.mean(axis=0)
    diag = (np.abs(O) ** 2).mean(axis=0) - np.abs(aO) ** 2
    z = O @ v
    return (O.conj().T @ z) / K - aO.conj() * (aO @ v) + lam * diag * v, aO, diag


@pytest.mark.parametrize("model,N,M,K,pin_cs,cs,cpt", SHAPES)
def test_fused_sv_matches_dense_and_two_pass(model, N, M, K, pin_cs, cs, cpt, monkeypatch):
    from neural_network_quantum_state_b200 import Engine
    if pin_cs is not None:
        monkeypatch.setenv("NQS_SV_CS", str(pin_cs))     # read by the engine when it plans the S*v launch (nqs_create / enable_sr)
    else:
        monkeypatch.delenv("NQS_SV_CS", raising=False)
    rng = np.random.default_rng(N * 1000 + M)
    out = {}
    for two_pass in (False, True):
        e = Engine(model, N, M, K, H, J, ALPHA, seed=7, two_pass_sv=two_pass)
        e.init_params_random(5)
        e.warm_up(2)
        variant = e.kernel_variant("sv")
        if two_pass:
            assert variant == "two_pass"
        elif pin_cs is not None:
            assert variant.startswith("fused_cs%d_cpt%d_" % (cs, cpt)), variant
            assert planned_cpt(e.P, cs) == cpt
        else:
            # unpinned: the planner takes the cluster size with the smallest estimated time (small clusters for narrow rows);
            # whatever it chose, its columns per thread follow the fat-warp rule
            import re
            m = re.match(r"fused_cs(\d+)_cpt(\d+)_", variant)
            assert m, variant
            assert planned_cpt(e.P, int(m.group(1))) == int(m.group(2)), variant
        O = e.get_lnpsiGradients()
        v = rng.normal(size=e.P) + 1j * rng.normal(size=e.P) if not out else out["v"]
        Sv, aO, diag = e.smatrix_dot(0.37, v)
        want, aO_w, diag_w = dense_sv(O, v, 0.37)
        assert_close(aO, aO_w, what="<O>")
        assert_close(diag, diag_w, atol=1e-11, what="diag")
        assert_close(Sv, want, what="S v (%s)" % variant)
        # twice in a row: the TMA slots / mbarrier phases are re-initialised per launch
        Sv2, _, _ = e.smatrix_dot(0.37, v)
        assert np.array_equal(Sv, Sv2), "S v is not run-to-run deterministic"
        out["v"] = v
        out[two_pass] = Sv
        e.close()
    assert_close(out[False], out[True], rtol=1e-12, what="fused vs two-pass")


def test_fused_sv_inside_cg_matches_two_pass_trajectory():
    """Whole SR steps (CG to tolerance) with the one-pass kernel give the same iteration counts and parameters."""
    from neural_network_quantum_state_b200 import Engine
    res = []
    for two_pass in (False, True):
        e = Engine("rbm", 32, 64, 256, H, J, ALPHA, seed=3, two_pass_sv=two_pass)
        e.init_params_random(11)
        e.warm_up(20)
        its = [e.sr_step(n_mc_steps=1, lr=0.05).cg_iters for _ in range(4)]
        res.append((its, e.get_params()))
        e.close()
    assert res[0][0] == res[1][0]
    assert_close(res[0][1], res[1][1], rtol=1e-9, what="params after 4 SR steps")


@pytest.mark.parametrize("N,M,K", [(16, 16, 64), (64, 128, 300), (128, 256, 150), (24, 40, 33)])
def test_o_generated_inside_first_sv_matches_separate_writer(N, M, K, monkeypatch):
    """nqs_sr_step writes O from inside the first S*v of the CG (sv_fused_kernel GEN: factors staged by TMA, elements formed
    in registers, used and stored; opt-in with NQS_SV_GEN=1, see engine.cu: alloc_sr).  By default the separate writer
    (oderiv_kernel) runs first.  Both must give the same CG trajectory: every later S*v of the step reads the O that the GEN
    launch wrote."""
    from neural_network_quantum_state_b200 import Engine
    res = []
    monkeypatch.setenv("NQS_CG_PERSIST", "0")   # the O-generating first product belongs to the launch-per-iteration path: like with like
    for gen in ("1", "0"):
        monkeypatch.setenv("NQS_SV_GEN", gen)
        e = Engine("rbm", N, M, K, H, J, ALPHA, seed=5)
        e.init_params_random(9)
        e.warm_up(30)
        steps = []
        for it in range(3):
            st = e.sr_step(n_mc_steps=1, lr=0.05, fixed_iters=5, lam=0.3)
            F, dx = e.get_sr_vectors()
            steps.append((st.e_mean, dx.copy()))
        res.append((steps, e.get_params()))
        e.close()
    # (the generating launch has its own cluster geometry, so the partial sums of that one product are grouped differently:
    # equal to rounding, not bit for bit; every later product of the step reads the O it wrote)
    for (e0, dx0), (e1, dx1) in zip(res[0][0], res[1][0]):
        assert e0 == pytest.approx(e1, rel=1e-13)
        assert_close(dx0, dx1, rtol=1e-10, what="dx: O-generating S*v vs separate O writer")
    assert_close(res[0][1], res[1][1], rtol=1e-10, what="params")
