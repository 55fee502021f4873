"""Second checker for the GPU-only semantics (blend update of lnpsi0, per-site energy, GPU solver settings): the
reference's OWN CUDA driver, gpu/src/LICH-train_rbm.cu compiled for sm_100 by baseline/Makefile, run on this box with the same
parameter files and -- through the Philox TRNG shim -- the same uniforms as libnqs_b200.so.  Skipped where the binary is absent
(it is built in the development container and travels with the snapshot; /root/reference does not exist on the GPU box)."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_cuda  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not ref_cuda.available(), reason="baseline/_ref/LICH-train_rbm-gpu-ref not built")
@pytest.mark.parametrize("L,nh,ns,nwarm,niter", [(16, 16, 512, 30, 6), (24, 48, 1024, 20, 4), (64, 128, 2048, 80, 3)])
def test_sr_trajectory_matches_reference_cuda_build(tmp_path, L, nh, ns, nwarm, niter):
    from neural_network_quantum_state_b200 import Engine
    from neural_network_quantum_state_b200.init import reference_init
    theta = float(ref_cuda.THETA_STR)
    h, J = -math.cos(theta), math.sin(theta)
    seed = 12345
    path = str(tmp_path)
    prefix = ref_cuda.prefix_for(path, L, nh)
    e = Engine("rbm", L, nh, ns, h, J, 2.0, seed=seed)
    e.set_params(reference_init("rbm", L, nh, np.random.default_rng(3)))
    e.save(prefix, 17)                       # full precision: both programs start from the same doubles
    e.load(prefix)
    ref = ref_cuda.run(L, nh, ns, niter, nwarm, seed, path, lr=0.02)
    assert len(ref["energies"]) == niter, ref["stdout_tail"]
    e.warm_up(nwarm)
    for it in range(niter):
        st = e.sr_step(n_mc_steps=1, lr=0.02)
        # the driver prints 7 significant digits
        assert st.e_mean.real == pytest.approx(ref["energies"][it], rel=2e-6, abs=1e-7), (it, st.e_mean.real, ref["energies"])
        assert st.rsd == pytest.approx(ref["rsd"][it], rel=2e-5, abs=1e-7)
    # the driver saved its parameters (10 significant digits) after the last iteration
    want = ref_cuda.load_params(prefix, L, nh)
    have = e.get_params()
    assert np.abs(have - want).max() <= 2e-9 * max(1.0, np.abs(want).max()) + 1e-9 * np.abs(want).max()
    e.close()


@pytest.mark.skipif(not ref_cuda.available("rbmtrsymm"), reason="baseline/_ref/LICH-train_rbmtrsymm-gpu-ref not built")
@pytest.mark.parametrize("L,nf,ns,nwarm,niter", [(16, 2, 512, 30, 6), (32, 3, 1024, 20, 4), (64, 2, 2048, 40, 3)])
def test_trsymm_trajectory_matches_reference_cuda_build(tmp_path, L, nf, ns, nwarm, niter):
    """The translation-symmetric RBM against gpu/src/LICH-train_rbmtrsymm.cu itself: periodic chain, the
    driver's own expansion / gradient kernels (impl_neural_quantum_state.cuh:1487-1553), same uniforms."""
    from neural_network_quantum_state_b200 import Engine
    theta = float(ref_cuda.THETA_STR)
    h, J = -math.cos(theta), math.sin(theta)
    seed = 777
    path = str(tmp_path)
    fname = ref_cuda.prefix_for(path, L, nf, driver="rbmtrsymm")
    e = Engine("rbmtrsymm", L, nf * L, ns, h, J, 2.0, pbc=True, seed=seed)
    assert e.P == L * nf + 1 + nf
    e.init_params_random(5)
    e.save(fname, 17)
    e.load(fname)
    ref = ref_cuda.run(L, nf, ns, niter, nwarm, seed, path, lr=0.02, driver="rbmtrsymm")
    assert len(ref["energies"]) == niter, ref["stdout_tail"]
    e.warm_up(nwarm)
    for it in range(niter):
        st = e.sr_step(n_mc_steps=1, lr=0.02)
        assert st.e_mean.real == pytest.approx(ref["energies"][it], rel=2e-6, abs=1e-7), (it, st.e_mean.real, ref["energies"])
        assert st.rsd == pytest.approx(ref["rsd"][it], rel=2e-5, abs=1e-7)
    want = ref_cuda.load_vars(fname)
    have = e.get_params()
    assert want.size == have.size
    assert np.abs(have - want).max() <= 3e-9 * max(1.0, np.abs(want).max())
    e.close()


@pytest.mark.skipif(not (ref_cuda.available("rbm") and ref_cuda.available("rbmtrsymm")), reason="reference CUDA drivers not built")
@pytest.mark.parametrize("driver,width", [("rbm", ("nh", 32)), ("rbmtrsymm", ("nf", 2))])
def test_cli_driver_prints_the_reference_drivers_table(tmp_path, driver, width):
    """Drop-in at the command line: the same argument list to the reference's own program and to this repository's
    LICH-train_*-gpu, same parameter file(s) to start from -- the iteration tables on stdout must agree to the printed digits."""
    import re
    import shutil
    import subprocess
    from neural_network_quantum_state_b200 import Engine, build
    from neural_network_quantum_state_b200.init import reference_init
    L, ns, niter, nwarm = 16, 512, 5, 40
    theta = float(ref_cuda.THETA_STR)
    dirs = [tmp_path / "ref", tmp_path / "ours"]
    for d in dirs:
        d.mkdir()
    if driver == "rbm":
        e = Engine("rbm", L, width[1], 4, 0.0, 0.0, 0.0, sampler_only=True)
        e.set_params(reference_init("rbm", L, width[1], np.random.default_rng(8)))
    else:
        e = Engine("rbmtrsymm", L, width[1] * L, 4, 0.0, 0.0, 0.0, sampler_only=True)
        e.init_params_random(8)
    for d in dirs:
        e.save(ref_cuda.prefix_for(str(d), L, width[1], driver=driver), 17)
    e.close()
    args = ["-L=%d" % L, "-%s=%d" % width, "-ns=%d" % ns, "-niter=%d" % niter, "-alpha=2", "-theta=%s" % ref_cuda.THETA_STR,
            "-ver=0", "-nwarm=%d" % nwarm, "-dev=0", "-lr=0.02", "-rsd=1e-30", "-seed=99"]
    ref_bin = ref_cuda.BINARY if driver == "rbm" else ref_cuda.BINARY_TRSYMM
    our_bin = os.path.join(build.BIN_DIR, "LICH-train_rbm-gpu" if driver == "rbm" else "LICH-train_rbmtrsymm-gpu")
    outs = []
    # this pair of reference binaries draws Philox uniforms (baseline/shim_cuda); the yarn2 pair is tests/test_gpu_yarn2.py
    env = dict(os.environ, NQS_RNG="philox")
    for b, d in zip((ref_bin, our_bin), dirs):
        r = subprocess.run([b] + args + ["-path=%s" % d], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        rows = [ln.split() for ln in r.stdout.splitlines() if re.match(r"^\s*\d+\s+\S+\s+\S+\s*$", ln)]
        outs.append([(int(a), float(b_), float(c)) for a, b_, c in rows])
        assert "# of loop\t<H>" in r.stdout and "# elapsed time:" in r.stdout
    assert len(outs[0]) == niter and len(outs[1]) == niter
    for (n0, e0, r0), (n1, e1, r1) in zip(*outs):
        assert n0 == n1
        assert e1 == pytest.approx(e0, rel=3e-6, abs=2e-7)
        assert r1 == pytest.approx(r0, rel=3e-5, abs=2e-7)
