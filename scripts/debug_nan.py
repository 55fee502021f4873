import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neural_network_quantum_state_b200 import Engine
from neural_network_quantum_state_b200.init import reference_init
N, M, K = 128, 256, 16384
H, J = -math.cos(math.pi/4), math.sin(math.pi/4)
params = reference_init("rbm", N, M, np.random.default_rng(20261018+3))
for fg in (False, True):
    e = Engine("rbm", N, M, K, H, J, 2.0, seed=20261018, force_generic=fg)
    e.set_params(params)
    e.warm_up(5)
    ht = e.get_htilda()
    ln = e.get_lnpsi()
    print("force_generic", fg, "htilda finite", np.isfinite(ht).all(), "nan count", np.isnan(ht).sum(), "lnpsi finite", np.isfinite(ln).all(), ht[:3])
    for it in range(3):
        st = e.sr_step(n_mc_steps=1, lr=1e-2)
        F, dx = e.get_sr_vectors()
        p = e.get_params()
        print(" step", it, st.e_mean, st.rsd, st.lam, st.cg_iters, st.finite, "F finite", np.isfinite(F).all(), "dx finite", np.isfinite(dx).all(), "params finite", np.isfinite(p).all(), "res2", st.cg_res2, st.cg_rhs2)
    e.close()
