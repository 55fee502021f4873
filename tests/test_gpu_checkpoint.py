"""Full-state checkpoint: the reference saves parameters only (gpu/include/optimizer.cuh:154-155,163), the
sidecar of nqs_checkpoint_save carries chains, RNG counter, lambda-schedule state and CG warm start -- a restarted run must
continue bit for bit."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
H, J, ALPHA = -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0


@pytest.mark.parametrize("model,N,M,K,pbc", [("rbm", 16, 32, 200, False), ("ffnn", 12, 24, 96, False), ("rbmtrsymm", 16, 32, 160, True)])
def test_restart_continues_bit_identically(tmp_path, model, N, M, K, pbc):
    from neural_network_quantum_state_b200 import Engine, NQSError
    rng = np.random.default_rng(3)
    spins0 = (2 * rng.integers(0, 2, size=(K, N)) - 1).astype(np.int8)

    def fresh():
        e = Engine(model, N, M, K, H, J, ALPHA, pbc=pbc, seed=21)
        e.init_params_random(5)
        return e

    a = fresh()
    a.warm_up(10, spins0)
    for _ in range(3):
        a.sr_step(n_mc_steps=1, lr=0.02)
    path = str(tmp_path / "state.bin")
    a.save_state(path)
    want = [(a.sr_step(n_mc_steps=1, lr=0.02), a.get_params().copy()) for _ in range(3)]
    want_spins, want_lnpsi = a.get_spinStates(), a.get_lnpsi()
    a.close()

    b = Engine(model, N, M, K, 0.3, -0.2, 1.0, pbc=pbc, seed=999)       # other Hamiltonian and seed: both travel with the state
    b.load_state(path)
    for st_w, p_w in want:
        st = b.sr_step(n_mc_steps=1, lr=0.02)
        assert st.e_mean == st_w.e_mean and st.rsd == st_w.rsd and st.lam == st_w.lam and st.cg_iters == st_w.cg_iters
        assert np.array_equal(b.get_params(), p_w)
    assert np.array_equal(b.get_spinStates(), want_spins)
    assert np.array_equal(b.get_lnpsi(), want_lnpsi)
    b.close()

    c = Engine(model, N, M, K + 1, H, J, ALPHA, pbc=pbc)
    with pytest.raises(NQSError):
        c.load_state(path)                                               # another shape
    with pytest.raises(NQSError):
        c.load_state(str(tmp_path / "missing.bin"))
    c.close()
