"""pynqs drop-in (ref python/pynqs/sampler.py + gpu/src/pywrapping_sampler.cu): same classes, kwargs, shapes and dtypes; the
sampler follows Sampler4SpinHalf (sequential site order, random initial spins) and the fixed-spin evaluation keeps the plain
RBM's visible-bias quirk of the two-instance PySampler."""
import math
import os

import numpy as np
import pytest

from helpers import assert_close


def test_surface_and_argument_checks_cpu():
    from neural_network_quantum_state_b200.pynqs import sampler, _pynqs_gpu
    for name in ("sRBMSampler", "dRBMSampler", "sRBMTrSymmSampler", "dRBMTrSymmSampler", "sRBMZ2PrSymmSampler",
                 "dRBMZ2PrSymmSampler", "sFFNNSampler", "dFFNNSampler", "sFFNNTrSymmSampler", "dFFNNTrSymmSampler"):
        assert hasattr(_pynqs_gpu, name)                   # ref PYBIND11_MODULE table, pywrapping_sampler.cu:120-132
    with pytest.raises(Exception):
        sampler.RBM(floatType="float64")                   # symmType omitted
    with pytest.raises(Exception):
        sampler.RBM(floatType="float16", symmType="None")
    r = sampler.RBM(floatType="float64", symmType="None")
    with pytest.raises(Exception):
        r.init(nInputs=4, nHiddens=4, nChains=4)           # essential arguments omitted
    # the float instantiations keep the float32 / complex64 surface (the engine computes in fp64)
    for name in ("RBM", "FFNN", "RBMTrSymm", "RBMZ2PrSymm", "FFNNTrSymm"):
        s_cls, d_cls = getattr(_pynqs_gpu, "s%sSampler" % name), getattr(_pynqs_gpu, "d%sSampler" % name)
        assert issubclass(s_cls, d_cls) and s_cls._real is np.float32 and s_cls._complex is np.complex64
        assert d_cls._real is np.float64 and d_cls._complex is np.complex128 and s_cls._model == d_cls._model


def test_reference_scripts_import_path_cpu():
    """python/meas_*.py do `from pynqs import sampler` with the directory that holds pynqs/ on PYTHONPATH (ref README.md:20-22);
    the drop-in must resolve the same way, with the arguments those scripts pass (floatType 'float32', symmType 'tr')."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "neural_network_quantum_state_b200"))
    code = ("from pynqs import sampler\n"
            "r = sampler.RBM(floatType='float32', symmType='tr')\n"
            "print(r._sampler.__name__, r._sampler._model)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.split() == ["sRBMTrSymmSampler", "rbmtrsymm"]


@pytest.mark.gpu
def test_pynqs_sampler_against_oracle(tmp_path):
    from neural_network_quantum_state_b200 import Engine
    from neural_network_quantum_state_b200.pynqs import sampler
    from oracle import nqs_oracle as o
    N, M, K = 10, 20, 96
    rng = np.random.default_rng(8)
    m = o.RBM(N, M, K, rng)
    m.W[...] *= 6.0
    m.a[...] = 0.2 * (rng.normal(size=N) + 1j * rng.normal(size=N))
    prefix = str(tmp_path / "net")
    w = Engine("rbm", N, M, 4, 0.0, 0.0, 0.0, sampler_only=True)
    w.set_params(m.variables)
    w.save(prefix)
    w.load(prefix)
    params = w.get_params()                                 # what a 10-digit file holds
    w.close()
    r = sampler.RBM(floatType="float64", symmType="None")
    r.init(nInputs=N, nHiddens=M, nChains=K, seedNumber=3, seedDistance=1000, path_to_load=prefix, init_mcmc_steps=4)
    r.do_mcmc_steps(3)
    s = r.get_spinStates()
    assert s.shape == (K, N) and s.dtype == np.float64 and set(np.unique(s)) <= {-1.0, 1.0}
    ln = r.get_lnpsi()
    assert ln.shape == (K,) and ln.dtype == np.complex128
    # lnpsi tracked by the sampler == amplitude of the current configuration (every chain accepted at least once by now)
    m.variables = params.copy()
    want = o.logcosh(s @ m.W + m.b[None, :]).sum(axis=1) + s @ m.a
    assert_close(ln, want, what="lnpsi of sampled states")
    # fixed-spin evaluation: the second ansatz instance never had spins set, so the visible-bias term is absent (ref quirk)
    fixed = (2 * rng.integers(0, 2, size=(K, N)) - 1).astype(np.float64)
    got = r.get_lnpsi_for_fixed_spins(fixed)
    assert_close(got, o.logcosh(fixed @ m.W + m.b[None, :]).sum(axis=1), what="lnpsi_for_fixed_spins")
    # distribution sanity: acceptance happened and the chains decorrelated from the initial state
    r2 = sampler.RBM(floatType="float64", symmType="None")
    r2.init(nInputs=N, nHiddens=M, nChains=K, seedNumber=3, seedDistance=1000, path_to_load=prefix, init_mcmc_steps=4)
    r2.do_mcmc_steps(3)
    assert np.array_equal(r2.get_spinStates(), s), "same seedNumber must reproduce the chains"
