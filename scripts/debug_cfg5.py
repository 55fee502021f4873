import sys, math, numpy as np
sys.path.insert(0, ".")
from neural_network_quantum_state_b200 import Engine
from neural_network_quantum_state_b200.init import reference_init
H, J, A = -math.cos(math.pi/4), math.sin(math.pi/4), 2.0
N, M, K = 128, 256, 2048
params = reference_init("rbm", N, M, np.random.default_rng(20261018+3))
for fg in (False, True):
  for nw in (5, 100):
    e = Engine("rbm", N, M, K, H, J, A, seed=20261018, force_generic=fg, max_predrawn_steps=N)
    e.set_params(params)
    e.warm_up(nw)
    ln = e.get_lnpsi(); ht = e.get_htilda(); sp = e.get_spinStates()
    bad = np.where(~np.isfinite(ht))[0]
    print("generic", fg, "nwarm", nw, e.kernel_variant("sweep"), "lnpsi finite", np.isfinite(ln).all(), "htilda bad", len(bad), ht.mean(), "mag", sp.mean())
    for it in range(3):
        st = e.sr_step(n_mc_steps=1, lr=0.01)
        print("   sr:", st.e_mean, st.finite, st.cg_iters, st.lam, st.rsd)
    e.close()
