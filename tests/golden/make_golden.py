"""Generates tests/golden/*.npz and tests/golden/files/* from the REFERENCE'S OWN CPU implementation
(oracle/_ref/libnqs_ref.so, built by `make -C oracle` from /root/reference/cpu/include; see oracle/ref_harness.cpp).

Run in the development container only (needs /root/reference to build the harness):
    python tests/golden/make_golden.py
The vectors pin oracle/nqs_oracle.py (tests/test_oracle_golden.py) and are what the CUDA engine is compared with on
the GPU box, where /root/reference does not exist.  Inputs are seeded numpy draws; every OUTPUT in the files below
was produced by the reference code, none by the oracle or the engine.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_cpu  # noqa: E402

H_FIELD, J_COUP, ALPHA = -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0   # theta = pi/4, gpu/src/LICH-train_rbm.cu:91


def synth_params(model, N, M, rng):
    """'Trained-like' magnitudes so that acceptance is ~0.3-0.9 and every S_ii > 0."""
    sw = math.sqrt(1.0 / (N + M))
    if model == "rbm":
        W = 0.8 * (rng.normal(0, sw, (N, M)) + 1j * rng.normal(0, sw, (N, M)))
        a = 0.3 * (rng.normal(size=N) + 1j * rng.normal(size=N))
        b = 0.5 * math.sqrt(1.0 / M) * (rng.normal(size=M) + 1j * rng.normal(size=M))
        return np.concatenate([W.ravel(), a, b])
    if model == "rbmtrsymm":       # [w (f*N+i) | a | b (alpha)], M = alpha*N
        al = M // N
        w = 0.8 * (rng.normal(0, 0.35, N * al) + 1j * rng.normal(0, 0.35, N * al))
        b = 0.3 * (rng.normal(size=al) + 1j * rng.normal(size=al))
        return np.concatenate([w, [0.2 - 0.1j], b])
    if model == "ffnntrsymm":      # [wi1 (f*N+i) | b1 (alpha) | w1o (alpha)], M = alpha*N
        al = M // N
        w = rng.normal(0, 0.3, N * al) + 0.1j * rng.normal(0, 0.3, N * al)
        b1 = 0.3 * (rng.normal(size=al) + 1j * rng.normal(size=al))
        w1o = rng.normal(0, 0.5, al) + 0.1j * rng.normal(0, 0.5, al)
        return np.concatenate([w, b1, w1o])
    W = rng.normal(0, sw, (N, M)) + 0.1j * rng.normal(0, sw, (N, M))
    b1 = 0.3 * (rng.normal(size=M) + 1j * rng.normal(size=M))
    w1o = rng.normal(0, math.sqrt(1.0 / M), M) + 0.1j * rng.normal(0, math.sqrt(1.0 / M), M)
    return np.concatenate([W.ravel(), b1, w1o])


def make_case(name, model, N, M, K, pbc, order, seed, n_warm, n_sr, lr, custom_init=False):
    rng = np.random.default_rng(seed)
    params = synth_params(model, N, M, rng)
    steps = (n_warm + n_sr + 2) * N
    U = rng.random((steps, K))
    r = ref_cpu.RefSampler(model, N, M, K, H_FIELD, J_COUP, ALPHA, pbc, order)
    r.set_params(params)
    r.set_uniforms(U)
    out = dict(model=model, N=N, M=M, K=K, pbc=int(pbc), order=order, h=H_FIELD, J=J_COUP, alpha=ALPHA,
               n_warm=n_warm, n_sr=n_sr, lr=lr, params=params, uniforms=U)
    if custom_init:
        s0 = np.where(rng.random((K, N)) < 0.5, 1.0, -1.0)
        r.set_initial_spins(s0)
        out["init_spins"] = s0.astype(np.int8)
    r.warm_up(n_warm)
    out["warm_spins"] = r.get_spins().astype(np.int8)
    out["warm_y"] = r.get_y()
    out["warm_lnpsi"] = r.get_lnpsi()
    out["flip_lnpsi"] = np.stack([r.forward_flip(i) for i in range(N)])      # [N][K]
    out["htilda"] = r.get_htilda()
    O = r.get_gradients()
    out["O"] = O
    lam = 0.37
    aO, diag = r.smatrix_set(O, lam)
    v = rng.normal(size=r.P) + 1j * rng.normal(size=r.P)
    out.update(sm_lambda=lam, sm_aO=aO, sm_diag=diag, sm_v=v, sm_Sv=r.smatrix_dot(v))
    E, rsd, lams, cgs, Fs, dxs = [], [], [], [], [], []
    for _ in range(n_sr):
        st = r.sr_step(1, lr)
        E.append(st["e_mean"]); rsd.append(st["rsd"]); lams.append(st["lam"]); cgs.append(st["cg_iters"])
        Fs.append(st["F"]); dxs.append(st["dx"])
        assert r.smatrix_diag().min() > 1e-6, "zero-variance parameter inside SR step (SURVEY 0.8): %s" % name
    out.update(sr_E=np.array(E), sr_rsd=np.array(rsd), sr_lambda=np.array(lams), sr_cg_iters=np.array(cgs),
               sr_F=np.stack(Fs), sr_dx=np.stack(dxs), final_params=r.get_params(),
               final_spins=r.get_spins().astype(np.int8), final_y=r.get_y(), final_lnpsi=r.get_lnpsi())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    # text parameter files written by the reference's own save() (drop-in file format, precision 10 = GPU default)
    fdir = os.path.join(HERE, "files")
    os.makedirs(fdir, exist_ok=True)
    r2 = ref_cpu.RefSampler(model, N, M, K, H_FIELD, J_COUP, ALPHA, pbc, order)
    r2.set_params(params)
    r2.save(os.path.join(fdir, name + "_"), 10)
    r2.close()
    r.close()
    assert diag.min() > 1e-6, "zero-variance parameter (SURVEY 0.8): pick another seed/size for %s" % name
    print(name, "P=%d" % out["params"].size, "E=", E, "cg=", cgs, "min diag=", diag.min())


if __name__ == "__main__":
    if not ref_cpu.available():
        raise SystemExit("build oracle/_ref first: make -C oracle")
    if "--tied" in sys.argv:
        # the tied-variable ansaetze the CPU tree has (cpu/include/neural_quantum_state.hpp:68-102, 184-217); added in round 2
        # without regenerating the four cases below.  M = expanded width alpha*N; random initial spins (a symmetric start leaves
        # zero-variance columns in the few-parameter O)
        make_case("rbmtrsymm_pbc", "rbmtrsymm", 10, 20, 64, True, "checkerboard", 201, 10, 3, 0.03, custom_init=True)
        make_case("ffnntrsymm_pbc", "ffnntrsymm", 8, 24, 64, True, "checkerboard", 202, 10, 3, 0.03, custom_init=True)
        raise SystemExit(0)
    make_case("rbm_obc", "rbm", 10, 16, 48, False, "checkerboard", 101, 12, 3, 0.05)
    make_case("rbm_pbc_odd_m", "rbm", 8, 13, 40, True, "checkerboard", 112, 14, 2, 0.05)
    make_case("rbm_seq_custom", "rbm", 9, 12, 50, False, "sequential", 103, 10, 2, 0.02, custom_init=True)
    make_case("ffnn_obc", "ffnn", 10, 20, 48, False, "checkerboard", 104, 12, 3, 0.05)
