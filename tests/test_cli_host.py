"""C++ host programs over the C ABI: LICH-train_rbm-gpu / LICH-train_ffnn-gpu (ref gpu/src/LICH-train_rbm.cu) and the
reference's command-line conventions (ref cpu/include/argparse.hpp:14-230).  CPU tests cover everything up to the first CUDA
call (the programs must fail LOUDLY without a GPU -- there is no CPU fallback); the GPU test runs a training and checks the
stdout table, the parameter files and that the trajectory equals the same run driven through the Python host."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def exe(name="LICH-train_rbm-gpu"):
    from neural_network_quantum_state_b200 import build
    build.build()
    path = os.path.join(build.BIN_DIR, name)
    assert os.path.exists(path)
    return path


def run(args, **kw):
    return subprocess.run(args, capture_output=True, text=True, timeout=600, **kw)


REQUIRED = ["-L=8", "-nh=8", "-ns=16", "-niter=2", "-alpha=2", "-theta=0.785", "-ver=1", "-dev=0", "-rsd=0.001"]


def test_help_lists_reference_options_and_defaults():
    r = run([exe(), "--help"])
    assert r.returncode == 1                      # the reference exits 1 after printing the table (argparse.hpp:48)
    for opt in ("L", "nh", "ns", "niter", "alpha", "theta", "ver", "nwarm", "nms", "dev", "lr", "rsd", "path", "seed", "ifprefix"):
        assert re.search(r"^\s*%s : " % opt, r.stdout, re.M), opt
    assert "nwarm : # of MCMC steps for warming-up (default : 100)" in r.stdout
    assert "(default : 1e-2)" in r.stdout and "(default : None)" in r.stdout
    assert "-option1=value1 -option2=value2" in r.stdout


def test_trsymm_driver_help_lists_its_own_options_and_defaults():
    """ref gpu/src/LICH-train_rbmtrsymm.cu:17-40: -nf instead of -nh, nwarm 500 and rsd 1e-3 by default."""
    r = run([exe("LICH-train_rbmtrsymm-gpu"), "--help"])
    assert r.returncode == 1
    assert re.search(r"^\s*nf : # of filters", r.stdout, re.M)
    assert not re.search(r"^\s*nh : ", r.stdout, re.M)
    assert "nwarm : # of MCMC steps for warming-up (default : 500)" in r.stdout
    assert "(default : 1e-3)" in r.stdout


def test_missing_and_malformed_options_exit_1_with_reference_messages():
    r = run([exe(), "-L=8"])
    assert r.returncode == 1
    assert "# error(in) ---> The following option is missing. : nh" in r.stderr
    assert "# error(in) ---> The following option is missing. : rsd" in r.stderr   # rsd has NO default in the reference
    assert "nwarm" not in r.stderr                                                   # defaults are filled in
    r = run([exe()] + REQUIRED + ["-L=9"])
    assert r.returncode == 1 and "# error(in-2) ---> The doubly occupied option is found! : L" in r.stderr
    r = run([exe()] + [a for a in REQUIRED if not a.startswith("-nh")] + ["-nh:8"])
    assert r.returncode == 1 and "# error(in-3)" in r.stderr
    r = run([exe()] + [a for a in REQUIRED if not a.startswith("-nh")] + ["-nh="])
    assert r.returncode == 1 and "# error(in-1) ---> Put the option correctly! : nh" in r.stderr
    r = run([exe()] + [a for a in REQUIRED if not a.startswith("-nh")] + ["-nh=8,"])
    assert r.returncode == 1 and "remove ',' at the last part" in r.stderr


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="CPU-only behaviour")
def test_without_gpu_the_driver_fails_loudly_after_echoing_the_arguments():
    r = run([exe()] + REQUIRED)
    assert r.returncode == 1
    assert "#===== updated arguments =====" in r.stdout
    assert re.search(r"^#\s+lr : 1e-2$", r.stdout, re.M) and re.search(r"^# ifprefix : None$", r.stdout, re.M)
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("structured_env", [False, True])
@pytest.mark.parametrize("prog,model,sufs", [("LICH-train_rbm-gpu", "rbm", ("Dw.dat", "Da.dat", "Db.dat")),
                                              ("LICH-train_ffnn-gpu", "ffnn", ("Dw1.dat", "Dw2.dat", "Db1.dat"))])
def test_training_run_matches_python_host_and_writes_reference_files(tmp_path, prog, model, sufs, structured_env):
    """structured_env: the reference-compatible CLI cannot pass engine flags; NQS_STRUCTURED_SV=1 in its environment switches the
    S*v to the tensor-core GEMMs on the factors of O.  Same table, same files (the Python run stays on the explicit-O path)."""
    from neural_network_quantum_state_b200 import Engine
    from neural_network_quantum_state_b200.init import reference_init
    L, nh, ns, niter, alpha, theta, seed = 12, 24, 256, 6, 2.0, 0.785398, 7
    tag = "RBMLICH" if model == "rbm" else "FFNNLICH"
    prefix = str(tmp_path / ("%s-L%dNH%dA2T0.785398V3" % (tag, L, nh)))     # trailing zeros stripped like the reference
    params = reference_init(model, L, nh, np.random.default_rng(4))
    e = Engine(model, L, nh, ns, -math.cos(theta), math.sin(theta), alpha, seed=seed)
    # the driver's sampler draws trng::yarn2 with seedDistance = niter*nms*L*ns (ref gpu/src/LICH-train_rbm.cu:82,97)
    e.set_rng("yarn2", seed, niter * 2 * L * ns)
    e.set_params(params)
    e.save(prefix)                      # the driver finds these files and loads them instead of its clock-seeded init
    e.load(prefix)                      # 10 significant digits survive the round trip: start both runs from the same numbers
    e.warm_up(20)
    want = [e.sr_step(n_mc_steps=2, lr=0.05) for _ in range(niter)]
    r = run([exe(prog), "-L=%d" % L, "-nh=%d" % nh, "-ns=%d" % ns, "-niter=%d" % niter, "-alpha=2", "-theta=%s" % theta, "-ver=3",
             "-dev=0", "-rsd=1e-9", "-nwarm=20", "-nms=2", "-lr=0.05", "-seed=%d" % seed, "-path=%s" % tmp_path],
            env=dict(os.environ, NQS_STRUCTURED_SV="1" if structured_env else "0"))
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert "# of loop\t<H>" in lines
    rows = [ln for ln in lines if re.match(r"^\s+\d+\s+\S+\s+\S+$", ln)]
    assert len(rows) == niter
    for n, (ln, st) in enumerate(zip(rows, want)):
        assert len(ln) == 5 + 16 + 16                                        # setw(5) setw(16) setw(16)
        it, en, rsd = ln.split()
        assert int(it) == n + 1
        assert float(en) == pytest.approx(st.e_mean.real, rel=2e-6)         # setprecision(7)
        assert float(rsd) == pytest.approx(st.rsd, rel=2e-6)
    assert any(ln.startswith("# elapsed time: ") and ln.endswith("(sec)") for ln in lines)
    for suf in sufs:
        assert os.path.exists(prefix + suf)
    e2 = Engine(model, L, nh, 4, 0.0, 0.0, 0.0, sampler_only=True)
    e2.load(prefix)
    np.testing.assert_allclose(e2.get_params(), e.get_params(), rtol=1e-8, atol=1e-12)
    e.close(); e2.close()


@pytest.mark.gpu
def test_missing_parameter_file_is_not_an_error(tmp_path):
    r = run([exe()] + REQUIRED + ["-path=%s" % tmp_path, "-nwarm=5"])
    assert r.returncode == 0, r.stderr
    assert "is not exist..." in r.stdout                                     # ref impl_neural_quantum_state.cuh:247-251
    assert os.path.exists(str(tmp_path / "RBMLICH-L8NH8A2T0.785V1Dw.dat"))


@pytest.mark.gpu
def test_bad_device_number(tmp_path):
    r = run([exe()] + [a for a in REQUIRED if not a.startswith("-dev")] + ["-dev=99"])
    assert r.returncode == 1 and "# error ---> dev(99) >= # of devices" in r.stderr


@pytest.mark.gpu
def test_cpp_sampler4spinhalf(tmp_path):
    """The C++ host mirror of Sampler4SpinHalf (ref gpu/include/meas.cuh:11-28, impl_meas.cuh:5-41) driven the way the reference's
    measurement programs drive it: tracked lnpsi == amplitude of the sampled configuration == forward(spins) of a second instance
    == the numpy oracle's amplitude."""
    import shutil
    import subprocess
    from neural_network_quantum_state_b200 import build
    from oracle import nqs_oracle as o
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe_path = str(tmp_path / "sampler4spinhalf_check")
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-o", exe_path, os.path.join(root, "tests", "sampler4spinhalf_check.cpp"),
                        "-L" + build.PKG_DIR, "-lnqs_b200", "-Wl,-rpath," + build.PKG_DIR], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    N, al, K = 12, 2, 48
    m = o.RBMTrSymm(N, al, K, np.random.default_rng(6))
    m.variables *= 5.0
    m.variables[N * al] = 0.1 - 0.05j
    path = str(tmp_path / "vars")
    m.save(path, 17)
    r = subprocess.run([exe_path, path, str(N), str(al), str(K)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = np.array([[float(t) for t in ln.split()] for ln in r.stdout.splitlines()])
    assert rows.shape == (K, N + 4)
    spins = rows[:, :N]
    assert set(np.unique(spins)) <= {-1.0, 1.0} and 0 < np.abs(spins.mean(axis=1)).mean() < 1
    tracked, fixed = rows[:, N] + 1j * rows[:, N + 1], rows[:, N + 2] + 1j * rows[:, N + 3]
    want = m.forward_spins(spins, save=False)
    np.testing.assert_allclose(fixed, want, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(tracked, want, rtol=1e-10, atol=1e-12)
