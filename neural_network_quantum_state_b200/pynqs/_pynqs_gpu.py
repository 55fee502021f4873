"""`_pynqs_gpu`: the extension-module surface of the reference's Python binding, served by libnqs_b200.so through ctypes.

ref: PYBIND11_MODULE(_pynqs_gpu, m) and PySampler<ansatz, T>, gpu/src/pywrapping_sampler.cu:9-18,29-132.  Each class takes the
same kwargs dict {nInputs, nHiddens, nChains, seedNumber, seedDistance} and offers load / warm_up / do_mcmc_steps /
get_spinStates / get_lnpsi / get_lnpsi_for_fixed_spins with the reference's shapes and dtypes.

Like PySampler it holds TWO ansatz instances: nqs0 is sampled by Sampler4SpinHalf (sequential site order 1,2,..,N-1,0,
gpu/include/impl_meas.cuh:12-21,33-34), nqs1 evaluates given configurations with forward(spins, lnpsi, saveSpinStates=false).
nqs1's own spin register is never set (all zero, as the reference's zero-initialised device vector), so the plain RBM's
visible-bias term -- computed from the MEMBER spins, ref impl_neural_quantum_state.cuh:119-120 -- contributes nothing there;
that reference quirk is kept bit for bit.

All ten classes of the module are served.  The engine computes in fp64 only: the d* classes are the reference's double
instantiations; the s* classes (what the reference's own python/meas_*.py scripts select, floatType = 'float32') keep the
float32 surface -- get_spinStates -> float32, get_lnpsi / get_lnpsi_for_fixed_spins -> complex64 -- on the same fp64
arithmetic, i.e. they are MORE precise than the reference's float instantiation and not bit-comparable to it (a float32 Markov
chain follows another path anyway: the reference draws uniform01_dist<float> there).
"""
from __future__ import annotations

import numpy as np

try:
    from ..engine import Engine
except ImportError:
    # imported as the TOP-LEVEL package `pynqs` -- PYTHONPATH=<repo>/neural_network_quantum_state_b200, the way the reference's
    # scripts find python/pynqs (README.md:20-22, python/meas_*.py: `from pynqs import sampler`)
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from neural_network_quantum_state_b200.engine import Engine


class _PySampler:
    _model = "rbm"
    _real, _complex = np.float64, np.complex128      # dtypes of the returned arrays (py::array_t<T> / std::complex<T> upstream)

    def __init__(self, kwargs: dict):
        self._N, self._M, self._K = int(kwargs["nInputs"]), int(kwargs["nHiddens"]), int(kwargs["nChains"])
        if self._model in ("rbmtrsymm", "ffnntrsymm"):
            self._M *= self._N     # RBMTrSymm / FFNNTrSymm(nInputs, alpha, nChains): "nHiddens" is the number of filters; the engine takes alpha*N
        elif self._model == "rbmz2prsymm":
            self._M *= 4           # RBMZ2PrSymm(nInputs, alpha, nChains): four hidden units per filter
        self._seed, self._seed_distance = int(kwargs["seedNumber"]), int(kwargs["seedDistance"])
        dev = int(kwargs.get("device", 0))
        # Sampler4SpinHalf has no Hamiltonian: h = J = 0; sequential order; O / CG buffers are not allocated
        mk = lambda: Engine(self._model, self._N, self._M, self._K, 0.0, 0.0, 0.0, order="sequential", seed=self._seed,
                            device=dev, sampler_only=True)
        self._nqs0, self._nqs1 = mk(), mk()
        # PySampler's smp_(psi, seedNumber, seedDistance) -> TRNGWrapper<T, trng::yarn2> (gpu/src/pywrapping_sampler.cu:29-40)
        self._nqs0.set_rng("yarn2", self._seed, self._seed_distance)
        # the reference ctor draws clock-seeded random parameters for both instances (load() normally overrides them)
        self._nqs0.init_params_random(self._seed * 2 + 1)
        self._nqs1.init_params_random(self._seed * 2 + 2)

    def load(self, prefix: str):
        self._nqs0.load(str(prefix))
        self._nqs1.set_params(self._nqs0.get_params())          # nqs0.copy_to(nqs1)

    def warm_up(self, nMCSteps: int):
        # Sampler4SpinHalf::initialize_ -> psi.initialize(lnpsi) with random +-1 spins (the reference seeds them from the clock,
        # gpu/include/neural_quantum_state.cuh:239-249; here from seedNumber so that runs are reproducible)
        rng = np.random.default_rng([self._seed, 0x5EED])
        spins = (2 * rng.integers(0, 2, size=(self._K, self._N)) - 1).astype(np.int8)
        self._nqs0.warm_up(int(nMCSteps), spins)

    def do_mcmc_steps(self, nMCSteps: int):
        self._nqs0.do_mcmc_steps(int(nMCSteps))

    def get_spinStates(self) -> np.ndarray:
        return self._nqs0.get_spinStates().astype(self._real).reshape(-1)   # flat [K*N] reals like py::array_t<T>(size)

    def get_lnpsi(self) -> np.ndarray:
        return self._nqs0.get_lnpsi().astype(self._complex, copy=False)

    def get_lnpsi_for_fixed_spins(self, spinStates) -> np.ndarray:
        s = np.asarray(spinStates).reshape(self._K, self._N)
        return self._nqs1.get_lnpsi_for_fixed_spins(np.rint(s).astype(np.int8)).astype(self._complex, copy=False)


class dRBMSampler(_PySampler):
    _model = "rbm"


class dFFNNSampler(_PySampler):
    _model = "ffnn"


class dRBMTrSymmSampler(_PySampler):
    """ref MAKE_PYSAMPLER_MODULE(m, "dRBMTrSymmSampler", spinhalf::RBMTrSymm, double), pywrapping_sampler.cu:125; kwargs["nHiddens"]
    is the number of filters alpha, load(path) reads the single variables file (impl_neural_quantum_state.cuh:484-517)."""
    _model = "rbmtrsymm"


class dRBMZ2PrSymmSampler(_PySampler):
    """ref MAKE_PYSAMPLER_MODULE(m, "dRBMZ2PrSymmSampler", spinhalf::RBMZ2PrSymm, double), pywrapping_sampler.cu:127; kwargs["nHiddens"]
    is the number of filters alpha, load(path) reads the single variables file (impl_neural_quantum_state.cuh:691-722)."""
    _model = "rbmz2prsymm"


class dFFNNTrSymmSampler(_PySampler):
    """ref MAKE_PYSAMPLER_MODULE(m, "dFFNNTrSymmSampler", spinhalf::FFNNTrSymm, double), pywrapping_sampler.cu:131."""
    _model = "ffnntrsymm"


def _single(cls):
    """the float instantiation of MAKE_PYSAMPLER_MODULE (pywrapping_sampler.cu:120-131): float32 / complex64 arrays out"""
    return type("s" + cls.__name__[1:], (cls,), {"_real": np.float32, "_complex": np.complex64, "__doc__": cls.__doc__})


sRBMSampler = _single(dRBMSampler)
sFFNNSampler = _single(dFFNNSampler)
sRBMTrSymmSampler = _single(dRBMTrSymmSampler)
sRBMZ2PrSymmSampler = _single(dRBMZ2PrSymmSampler)
sFFNNTrSymmSampler = _single(dFFNNTrSymmSampler)
