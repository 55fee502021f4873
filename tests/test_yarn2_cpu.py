"""trng::yarn2 (the reference's source of Metropolis uniforms, gpu/include/trng4cuda.cuh:14-65).  TRNG4 is a
third-party library absent from /root/reference (pinned v4.22, cmake/FindTRNG4.cmake:46-48): the stream is restated from the
published algorithm three times, separately -- oracle/yarn2.py (python integers), csrc/yarn2.cuh (the product, tables + binary
matrix powers) and baseline/shim_yarn2 (the library's class shape, compiled into the reference's CUDA drivers).  PARITY UNPINNED
against the real library; these tests pin the algebra of the published generator and the three restatements to each other."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import yarn2 as y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRIMES_OF_M_MINUS_1 = [2, 3, 7, 11, 31, 151, 331]      # 2^31 - 2 = 2 * 3^2 * 7 * 11 * 31 * 151 * 331


def test_constants_are_the_published_ones_by_their_algebra():
    """A wrongly remembered constant would almost surely fail these: the output base must generate F_m^*, and L'Ecuyer's
    multipliers make x^2 - a0 x - a1 primitive over F_m (period m^2 - 1)."""
    n = y.M - 1
    for q in PRIMES_OF_M_MINUS_1:
        while n % q == 0:
            n //= q
    assert n == 1
    assert all(pow(y.GEN, (y.M - 1) // q, y.M) != 1 for q in PRIMES_OF_M_MINUS_1)
    ident, order = (1, 0, 0, 1), y.M * y.M - 1           # m^2 - 1 = (m - 1) 2^31
    assert y._matpow(order) == ident
    assert all(y._matpow(order // q) != ident for q in PRIMES_OF_M_MINUS_1)


def test_jump_equals_stepping_and_is_additive():
    for seed, n in [(0, 1), (1, 15), (12345, 16), (2**31 - 1, 17), (2**40 + 3, 1000)]:
        a, b = y.Yarn2(seed), y.Yarn2(seed)
        a.jump(n)
        for _ in range(n):
            b.step()
        assert (a.r0, a.r1) == (b.r0, b.r1)
    a, b = y.Yarn2(99), y.Yarn2(99)
    a.jump(2**45 + 12345)
    b.jump(2**45)
    b.jump(12345)
    assert (a.r0, a.r1) == (b.r0, b.r1)
    # the jump argument is a 64-bit unsigned in the reference: 2ul * seedDistance * k wraps
    a, b = y.Yarn2(5), y.Yarn2(5)
    a.jump(2 * (2**62) * 3)
    b.jump((2 * (2**62) * 3) % 2**64)
    assert (a.r0, a.r1) == (b.r0, b.r1)


def test_seed_and_default_state():
    e = y.Yarn2()
    assert (e.r0, e.r1) == (0, 1)
    e.seed(y.M + 5)
    assert (e.r0, e.r1) == (5, 1)
    e = y.Yarn2()                                           # first draw from the default state: r0 = a1
    assert e.next_int() == pow(y.GEN, y.A1, y.M)


def test_uniforms_in_unit_interval_and_vectorised_form_matches_scalar():
    u = y.chain_uniforms(seed=7, seed_distance=1000, n_chains=9, steps=50, chain_offset=3, skip=11)
    assert u.shape == (50, 9) and (u >= 0).all() and (u < 1).all()
    for k in range(9):
        e = y.Yarn2(7)
        e.jump(2 * 1000 * (3 + k))
        e.jump(11)
        assert np.array_equal(u[:, k], np.array([e.uniform01() for _ in range(50)]))
    # chains are windows of ONE sequence: chain k+1 starts 2*seedDistance draws after chain k
    w = y.chain_uniforms(seed=7, seed_distance=4, n_chains=3, steps=20)
    assert np.array_equal(w[8:, 0], w[:12, 1]) and np.array_equal(w[8:, 1], w[:12, 2])
    # crude equidistribution of a long window
    z = y.chain_uniforms(seed=1, seed_distance=0, n_chains=1, steps=20000)[:, 0]
    assert abs(z.mean() - 0.5) < 0.01 and abs(z.var() - 1 / 12) < 0.005


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not found")
def test_product_header_and_reference_shim_agree_with_the_oracle(tmp_path):
    """csrc/yarn2.cuh's generator arithmetic (executed on the host; no kernel runs) and baseline/shim_yarn2 (what the reference
    CUDA drivers are compiled against) print bit-identical uniforms to oracle/yarn2.py."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "yarn2_host_check")
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    r = subprocess.run([nvcc, "-O1", "-std=c++17", "-arch=sm_100", "-o", exe, os.path.join(ROOT, "tests", "yarn2_host_check.cu")],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    for seed, dist, chain, skip in [(0, 0, 0, 0), (7, 1000, 5, 0), (12345, 2097152 * 100, 16383, 640), (2**63 + 9, 2**61 + 1, 77, 3)]:
        r = subprocess.run([exe, str(seed), str(dist), str(chain), str(skip), "40"], capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stderr)
        got = np.array([[float.fromhex(t) for t in ln.split()] for ln in r.stdout.splitlines()])
        e = y.Yarn2(seed)
        e.jump((2 * dist * chain) % 2**64)
        e.jump(skip)
        want = np.array([e.uniform01() for _ in range(40)])
        assert np.array_equal(got[:, 0], want), "csrc/yarn2.cuh"
        assert np.array_equal(got[:, 1], want), "baseline/shim_yarn2"


def test_committed_known_answers_of_the_restatement():
    """tests/golden/yarn2_kat.json: regression vectors of the RESTATEMENT (not of TRNG4) together with the 6-line program that
    would pin it against the library once TRNG4 v4.22 can be obtained."""
    import json
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "yarn2_kat.json")))
    c0, c1, c2 = kat["cases"]
    e = y.Yarn2()
    assert [e.next_int() for _ in range(8)] == c0["ints"]
    e = y.Yarn2(12345)
    e.jump(2 * 1000 * 7)
    assert [float.hex(e.uniform01()) for _ in range(8)] == c1["hex"]
    e = y.Yarn2(2**64 - 1)
    assert (e.r0, e.r1) == (y.M - 1, 1)                     # int64(2^64-1) = -1 -> m-1
    e.jump(1 << 40)
    assert [e.next_int() for _ in range(4)] == c2["ints"]
    assert "trng/yarn2.hpp" in kat["check_program"]


def test_jump_properties_hypothesis():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 2**64 - 1), st.integers(0, 2**63), st.integers(0, 2**63 - 1))
    def prop(seed, a, b):
        e1, e2, e3 = y.Yarn2(seed), y.Yarn2(seed), y.Yarn2(seed)
        e1.jump(a)
        e1.jump(b)
        e2.jump(b)
        e2.jump(a)
        e3.jump((a + b) % 2**64)                             # a + b < 2^64 here: no wrap
        assert (e1.r0, e1.r1) == (e2.r0, e2.r1) == (e3.r0, e3.r1)
        assert 0 <= e1.r0 < y.M and 0 <= e1.r1 < y.M
        u = e1.uniform01()
        assert 0.0 <= u < 1.0

    prop()
