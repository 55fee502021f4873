"""tcgen05 int8 rows kernel (csrc/rows_umma.cuh: spins x Ozaki-split B, int32 accumulators in TMEM, int64 recombination)
against the fp64 tensor-core kernel it replaces (spin_rows_dmma_kernel) and the numpy oracle:
  * theta = S W + b and lnpsi = sum log cosh theta (ref: gpu/include/impl_neural_quantum_state.cuh:78,114);
  * z = O v of the structured S*v (functor_for_CG.cuh:104-127 on the factors of O).
The split is error-free up to the rounding of B to 53 bits below its column maximum, so the two kernels agree to a few ulps of
the column scale: bar 1e-13 relative to the largest entry (the oracle bar stays the suite's 1e-10)."""
import numpy as np
import pytest

from helpers import assert_close
from oracle import nqs_oracle as o
from test_gpu_parity import ALPHA, H, J, _engine, synth

pytestmark = pytest.mark.gpu

# (model, N, M, K): M2 = 2M a multiple of the 32-column chunk / ragged last chunk; N a multiple of 32 / of 16 / odd; K ragged
SHAPES = [
    ("rbm", 128, 256, 300),    # cfg3 network
    ("rbm", 64, 128, 130),     # cfg2 network
    ("rbm", 33, 40, 200),      # odd N (byte-wise spin tile fill, zero padding to 64 sites), M2 = 80: last chunk half empty
    ("rbm", 16, 16, 512),      # cfg1 network: one chunk, 32 sites of padding
    ("rbm", 256, 64, 129),     # 8 MMAs per chunk
    ("rbm", 48, 7, 64),        # M2 = 14 < one chunk
    ("ffnn", 16, 48, 77),
    ("ffnn", 128, 96, 257),
]


def _run(model, N, M, K, umma, monkeypatch, structured):
    monkeypatch.setenv("NQS_ROWS_UMMA", "1" if umma else "0")
    e = _engine(model, N, M, K, H, J, ALPHA, seed=5, structured_sv=structured)
    return e


@pytest.mark.parametrize("model,N,M,K", SHAPES)
def test_umma_theta_and_lnpsi_match_dmma_and_oracle(model, N, M, K, monkeypatch):
    rng = np.random.default_rng(N * 31 + M)
    params = synth(model, N, M, rng)
    spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(K, N))
    out = []
    for umma in (True, False):
        e = _run(model, N, M, K, umma, monkeypatch, False)
        e.set_params(params)
        e.initialize(spins)
        assert e.kernel_variant("theta") == ("umma_i8_ozaki7_rows" if umma else "dmma_rows"), e.kernel_variant("theta")
        out.append((e.get_theta(), e.get_lnpsi(), e.get_htilda()))     # htilda: s.J.s through the same kernel (B = J)
        e.close()
    assert_close(out[0][0], out[1][0], rtol=1e-13, atol=0.0, what="theta: tcgen05 int8 vs DMMA")
    assert_close(out[0][1], out[1][1], rtol=1e-12, atol=0.0, what="lnpsi: tcgen05 int8 vs DMMA")
    assert_close(out[0][2], out[1][2], rtol=1e-12, atol=0.0, what="local energy: tcgen05 int8 vs DMMA")
    net = o.make_ansatz(model, N, M, K)
    net.variables = params.copy()
    lnpsi = net.initialize(spins.astype(np.float64))
    assert_close(out[0][0], net.y, what="theta vs oracle")
    assert_close(out[0][1], lnpsi, what="lnpsi vs oracle")


@pytest.mark.parametrize("model,N,M,K", SHAPES)
def test_umma_structured_sv_matches_dmma(model, N, M, K, monkeypatch):
    rng = np.random.default_rng(N * 13 + M)
    params = synth(model, N, M, rng)
    P = params.size
    v = rng.normal(size=P) + 1j * rng.normal(size=P)
    v[::7] *= 1e-9          # columns of very different magnitude: each takes its own power-of-two scale
    v[3::11] *= 1e6
    out = []
    for umma in (True, False):
        e = _run(model, N, M, K, umma, monkeypatch, True)
        e.set_params(params)
        e.warm_up(3)
        out.append(e.smatrix_dot(0.25, v))
        e.close()
    for a, b, what in zip(out[0], out[1], ("S v", "<O>", "diag S")):
        assert_close(a, b, rtol=1e-12, atol=0.0, what=what + ": tcgen05 int8 vs DMMA")


def test_umma_propagates_nonfinite_parameters(monkeypatch):
    """A NaN in W must reach theta (the fp64 GEMM would propagate it; the integer split cannot represent it, so the column's
    scale carries it)."""
    model, N, M, K = "rbm", 32, 24, 128
    params = synth(model, N, M, np.random.default_rng(1))
    params[5 * M + 3] = np.nan
    e = _run(model, N, M, K, True, monkeypatch, False)
    e.set_params(params)
    e.initialize(np.ones((K, N), dtype=np.int8))
    th = e.get_theta()
    assert np.isnan(th[:, 3]).all()
    assert np.isfinite(np.delete(th, 3, axis=1)).all()
    e.close()
