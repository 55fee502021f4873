"""`pynqs.sampler`: same user-facing class as the reference's python/pynqs/sampler.py:12-72 (RBM(floatType=, symmType=) ->
init(...) -> do_mcmc_steps / get_spinStates / get_lnpsi / get_lnpsi_for_fixed_spins), backed by the B200 engine."""
import numpy as np

from . import _pynqs_gpu

_REQUIRED_CTOR = ("floatType", "symmType")
_REQUIRED_INIT = ("nInputs", "nHiddens", "nChains", "seedNumber", "seedDistance", "path_to_load", "init_mcmc_steps")


def argchecker(kwargs, required):
    missing = [name for name in required if name not in kwargs]
    if missing:
        raise Exception("You omit an essential argument registered in :", list(required))


class RBM:
    # (floatType, symmType) -> class of _pynqs_gpu, as the reference's if/elif table (sampler.py:27-40)
    _TABLE = {("float32", "None"): "sRBMSampler", ("float64", "None"): "dRBMSampler",
              ("float32", "tr"): "sRBMTrSymmSampler", ("float64", "tr"): "dRBMTrSymmSampler",
              ("float32", "z2pr"): "sRBMZ2PrSymmSampler", ("float64", "z2pr"): "dRBMZ2PrSymmSampler"}

    def __init__(self, **kwargs):
        argchecker(kwargs, _REQUIRED_CTOR)
        self._floatType, self._symmType = kwargs["floatType"], kwargs["symmType"]
        try:
            self._sampler = getattr(_pynqs_gpu, self._TABLE[(self._floatType, self._symmType)])
        except KeyError:
            raise Exception(" --hint:  floatType: float32 or float64 / symmType: None, tr, z2pr")

    def init(self, **kwargs):
        argchecker(kwargs, _REQUIRED_INIT)
        self._rbm = self._sampler(kwargs)
        self._nInputs, self._nChains = int(kwargs["nInputs"]), int(kwargs["nChains"])
        self._rbm.load("%s" % str(kwargs["path_to_load"]))
        self._rbm.warm_up(int(kwargs["init_mcmc_steps"]))

    def do_mcmc_steps(self, mcmc_steps):
        self._rbm.do_mcmc_steps(mcmc_steps)

    def get_spinStates(self):
        return self._rbm.get_spinStates().reshape([-1, self._nInputs])

    def get_lnpsi(self):
        return self._rbm.get_lnpsi()

    def get_lnpsi_for_fixed_spins(self, spinStates):
        spinStates = np.array(spinStates).astype(self._floatType).reshape([self._nChains, self._nInputs])
        return self._rbm.get_lnpsi_for_fixed_spins(spinStates)
