"""The measurement flows of the reference's python/meas_{smag,renyi,fidelity}.py through the pynqs drop-in, with
the arguments those scripts use (floatType = 'float32', symmType = 'tr'; also float64 / z2pr), against EXACT answers from the
enumeration of all 2^N configurations of a small chain (amplitudes from the numpy oracle): |m|, the second Renyi entropy by the
swap estimator, and the fidelity of two states.  Oracle-independent of any sampler: a wrong stationary distribution, a wrong
fixed-spin amplitude or a broken second instance shows up as a bias."""
import itertools
import math

import numpy as np
import pytest

from oracle import nqs_oracle as o

pytestmark = pytest.mark.gpu
N, AL, K = 8, 2, 4096
CLS = {"tr": o.RBMTrSymm, "z2pr": o.RBMZ2PrSymm}


def _state(symm, seed, scale):
    t = CLS[symm](N, AL, 1, np.random.default_rng(seed))
    v = t.variables * scale
    if symm == "tr":
        v[N * AL] = 0.05 + 0.02j
    return v


def _exact_amplitudes(symm, v):
    conf = np.array(list(itertools.product([1.0, -1.0], repeat=N)))            # first site = slowest index
    m = CLS[symm](N, AL, conf.shape[0])
    m.variables = v.copy()
    return conf, np.exp(m.forward_spins(conf, save=False))


def _sampler(floatType, symm, path, seed):
    from neural_network_quantum_state_b200.pynqs import sampler
    r = sampler.RBM(floatType=floatType, symmType=symm)
    r.init(nInputs=N, nHiddens=AL, nChains=K, seedNumber=seed, seedDistance=123456789, path_to_load=path, init_mcmc_steps=60)
    return r


def _save(tmp_path, symm, v, name):
    m = CLS[symm](N, AL, 1)
    m.variables = v.copy()
    path = str(tmp_path / name)
    m.save(path, 17)
    return path


def _check(samples, exact, what):
    samples = np.asarray(samples, dtype=np.float64)
    mean = samples.mean()
    err = math.sqrt(((samples - mean) ** 2).sum() / (samples.size * (samples.size - 1)))
    assert abs(mean - exact) < 6 * err + 2e-3 * max(1.0, abs(exact)), "%s: %.6f +- %.1e (MC) vs %.6f (exact)" % (what, mean, err, exact)
    return mean, err


@pytest.mark.parametrize("floatType,symm", [("float32", "tr"), ("float64", "tr"), ("float32", "z2pr")])
def test_magnetisation_and_renyi_entropy_match_exact_enumeration(tmp_path, floatType, symm):
    v = _state(symm, 11, 8.0 if symm == "tr" else 6.0)
    conf, psi = _exact_amplitudes(symm, v)
    prob = np.abs(psi) ** 2
    prob /= prob.sum()
    exact_m = float((prob * np.abs(conf.mean(axis=1))).sum())
    ell = 3
    A = psi.reshape(2 ** ell, 2 ** (N - ell)) / math.sqrt((np.abs(psi) ** 2).sum())
    rho = A @ A.conj().T
    exact_tr2 = float(np.real(np.trace(rho @ rho)))
    path = _save(tmp_path, symm, v, "vars")
    rbms = [_sampler(floatType, symm, path, 1 * 123456789), _sampler(floatType, symm, path, 2 * 123456789)]   # meas_renyi.py:41
    mag, tr2 = [], []
    for _ in range(12):
        for r in rbms:
            r.do_mcmc_steps(4)
        s0, s1 = rbms[0].get_spinStates(), rbms[1].get_spinStates()
        assert s0.dtype == np.dtype(floatType) and s0.shape == (K, N)
        l0, l1 = rbms[0].get_lnpsi(), rbms[1].get_lnpsi()
        assert l0.dtype == (np.complex64 if floatType == "float32" else np.complex128)
        mag.append(np.mean(np.abs(np.mean(s0, axis=1))))
        s2, s3 = s0.copy(), s1.copy()
        s2[:, :ell], s3[:, :ell] = s1[:, :ell], s0[:, :ell]                       # swap_operations, meas_renyi.py:30-35
        l2, l3 = rbms[0].get_lnpsi_for_fixed_spins(s2), rbms[1].get_lnpsi_for_fixed_spins(s3)
        tr2.append(np.mean(np.exp(l2.astype(np.complex128) + l3 - l0 - l1)).real)
    _check(mag, exact_m, "<|m|>")
    mean, _ = _check(tr2, exact_tr2, "tr rho_A^2")
    assert abs(-math.log(mean) + math.log(exact_tr2)) < 0.05


@pytest.mark.parametrize("floatType", ["float32", "float64"])
def test_fidelity_of_two_states_matches_exact_overlap(tmp_path, floatType):
    va, vb = _state("tr", 21, 6.0), _state("tr", 22, 6.0)
    vb = 0.7 * va + 0.3 * vb                                                      # neighbouring states: overlap of order one
    _, pa = _exact_amplitudes("tr", va)
    _, pb = _exact_amplitudes("tr", vb)
    exact_f2 = float(abs(np.vdot(pa, pb)) ** 2 / (np.vdot(pa, pa).real * np.vdot(pb, pb).real))
    rbms = [_sampler(floatType, "tr", _save(tmp_path, "tr", va, "a"), 5), _sampler(floatType, "tr", _save(tmp_path, "tr", vb, "b"), 6)]
    f2 = []
    for _ in range(12):
        for r in rbms:
            r.do_mcmc_steps(4)
        s0, s1 = rbms[0].get_spinStates(), rbms[1].get_spinStates()
        l00, l11 = rbms[0].get_lnpsi().astype(np.complex128), rbms[1].get_lnpsi().astype(np.complex128)
        l01, l10 = rbms[0].get_lnpsi_for_fixed_spins(s1), rbms[1].get_lnpsi_for_fixed_spins(s0)     # meas_fidelity.py:47-49
        # NOTE the script's naming: rbms[0] evaluates the samples of rbms[1] and vice versa
        f2.append(np.mean(np.exp(l01 - l11) * np.exp(l10 - l00)).real)
    _check(f2, exact_f2, "F^2")
