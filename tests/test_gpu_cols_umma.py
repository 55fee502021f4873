"""tcgen05 int8 kernel for O^H z and the SR setup sums (csrc/cols_umma.cuh: S^T C with C = conj(T) z split on the fly into 7
int8 digit planes, int32 accumulators in TMEM) against the fp64 tensor-core kernel it replaces (spin_cols_dmma_kernel) and,
through the structured S*v, against the dense formula on the engine's own O.
ref: SMatrixForCG::dot, gpu/include/functor_for_CG.cuh:104-127; setup sums optimizer.cuh:140-143."""
import numpy as np
import pytest

from helpers import assert_close
from test_gpu_parity import ALPHA, H, J, _engine, synth
from test_gpu_sv_fused import dense_sv

pytestmark = pytest.mark.gpu

# (model, N, M, K): one 64-chain block per CTA / several (the double-buffered tiles are reused) / ragged last block;
# M2 a multiple of the 64-column group / ragged; N = 128 / a multiple of 16 / odd
SHAPES = [
    ("rbm", 128, 256, 300),
    ("rbm", 64, 128, 130),
    ("rbm", 33, 40, 200),
    ("rbm", 16, 16, 512),
    ("rbm", 48, 7, 64),
    ("rbm", 32, 256, 4000),     # 18 chunks of 256 chains: 4 blocks per CTA
    ("rbm", 128, 64, 5000),     # 74 chunks, ragged
    ("ffnn", 16, 48, 77),
    ("ffnn", 128, 96, 2570),
]


@pytest.mark.parametrize("model,N,M,K", SHAPES)
def test_umma_cols_structured_sv_and_setup_match_dmma_and_dense(model, N, M, K, monkeypatch):
    rng = np.random.default_rng(N * 17 + M)
    params = synth(model, N, M, rng)
    P = params.size
    v = rng.normal(size=P) + 1j * rng.normal(size=P)
    v[::5] *= 1e-7
    out = []
    for umma in ("1", "0"):
        monkeypatch.setenv("NQS_COLS_UMMA", umma)
        e = _engine(model, N, M, K, H, J, ALPHA, seed=5, structured_sv=True)
        assert e.kernel_variant("sv").startswith("structured_umma_i8" if umma == "1" else "structured_dmma"), e.kernel_variant("sv")
        e.set_params(params)
        e.warm_up(3)
        Sv, aO, diag = e.smatrix_dot(0.25, v)
        if umma == "1":
            want, aO_w, diag_w = dense_sv(e.get_lnpsiGradients(), v, 0.25)
            assert_close(aO, aO_w, what="<O> vs dense")
            assert_close(diag, diag_w, atol=1e-11, what="diag vs dense")
            assert_close(Sv, want, what="S v vs dense")
        out.append((Sv, aO, diag))
        e.close()
    for a, b, what in zip(out[0], out[1], ("S v", "<O>", "diag S")):
        assert_close(a, b, rtol=1e-12, atol=0.0, what=what + ": tcgen05 int8 vs DMMA")


def test_umma_cols_sr_trajectory_matches_dmma(monkeypatch):
    """Whole SR steps (setup GEMM + CG to tolerance) on either kernel: same iteration counts, same parameters to rounding."""
    res = []
    for umma in ("1", "0"):
        monkeypatch.setenv("NQS_COLS_UMMA", umma)
        monkeypatch.setenv("NQS_ROWS_UMMA", umma)
        e = _engine("rbm", 32, 64, 2048, H, J, ALPHA, seed=3, structured_sv=True)
        e.init_params_random(11)
        e.warm_up(20)
        its = [e.sr_step(n_mc_steps=1, lr=0.05).cg_iters for _ in range(4)]
        res.append((its, e.get_params()))
        e.close()
    assert res[0][0] == res[1][0]
    assert_close(res[0][1], res[1][1], rtol=1e-9, what="params after 4 SR steps")
