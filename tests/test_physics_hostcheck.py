"""CPU: the product's device physics header (athena-gamma_b200/csrc/ab_physics.cuh), compiled
for the host by a test-only shim, against the C oracle on random states.  Bit-exact.
This pins the arithmetic of the CUDA kernels before any GPU time is spent; the GPU parity
tests (tests/test_gpu_*.py) then cover the kernels themselves."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
import util

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "hostcheck", "libphysics_host.so")


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "physics_host.cpp")
    hdr = os.path.join(HERE, "..", "athena-gamma_b200", "csrc", "ab_physics.cuh")
    if (not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src),
                                                              os.path.getmtime(hdr))):
        subprocess.run(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off",
                        "-x", "c++", src, "-o", SO], check=True)
    L = C.CDLL(SO)
    dp = C.POINTER(C.c_double)
    L.hc_riemann.argtypes = [C.c_int, C.c_int, C.c_long, dp, dp, dp, dp, dp, C.c_double,
                             C.c_double, C.c_double, dp, dp]
    L.hc_riemann_iso.argtypes = [C.c_int, C.c_int, C.c_long, dp, dp, dp, C.c_double, C.c_double,
                                 dp]
    L.hc_recon_char.argtypes = [C.c_int, C.c_int, C.c_long, dp, dp, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_double, dp, dp]
    L.hc_plm.argtypes = [C.c_long, C.c_int, dp, dp, dp, C.c_double, C.c_double, dp, dp]
    L.hc_ppm.argtypes = [C.c_long, C.c_int, dp, dp, dp, dp, dp, dp, dp]
    L.hc_fastdiv_mismatches.argtypes = [C.c_int, C.c_long, C.c_long, C.c_long]
    L.hc_fastdiv_mismatches.restype = C.c_long
    return L


def test_multiply_high_division_is_exact(hc):
    """FastDiv (ab_types.h): the index decode of every flattened kernel and of k_copy_boxes.
    t / d for 0 <= t < 2^31: every t up to 2^20 and windows at the top of the range for the
    divisors that occur (row lengths 2..600, 513/514/516/517, powers of two and their
    neighbours), strided sweeps of the whole range for a sample of large divisors."""
    rng = np.random.default_rng(5)
    top = 2**31 - 1
    small = list(range(1, 601)) + [2**k + e for k in range(10, 31) for e in (-1, 0, 1)]
    for d in small:
        if d > top:
            continue
        assert hc.hc_fastdiv_mismatches(d, 0, 2**20, 1) == 0, d
        assert hc.hc_fastdiv_mismatches(d, top - 2**18, top, 1) == 0, d
    for d in [int(x) for x in rng.integers(2, top, 300)] + [top, top - 1, 3, 7, 641, 65537]:
        assert hc.hc_fastdiv_mismatches(d, 0, top, 65521) == 0, d
        for k in (1, 2, 3, 1000):                  # around multiples of d: where a wrong constant shows
            lo = max(0, min(top, k*d) - 40)
            assert hc.hc_fastdiv_mismatches(d, lo, min(top, lo + 80), 1) == 0, (d, k)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def random_states(rng, n, mhd, regime):
    nw = 7 if mhd else 5
    w = np.zeros((nw, n))
    w[0] = np.exp(rng.uniform(-3, 3, n))
    vs = {"subsonic": 0.3, "supersonic": 5.0, "mixed": 2.0}[regime]
    w[1:4] = rng.normal(0, vs, (3, n))
    w[4] = np.exp(rng.uniform(-3, 3, n))
    if mhd:
        w[5:7] = rng.normal(0, 1.0, (2, n))
    return w


@pytest.mark.parametrize("solver,mhd", [("hllc", False), ("hlle", False), ("roe", False),
                                        ("lhllc", False), ("hlld", True), ("hlle", True),
                                        ("roe", True), ("lhlld", True), ("llf", False),
                                        ("llf", True)])
@pytest.mark.parametrize("regime", ["subsonic", "supersonic", "mixed"])
def test_riemann_matches_oracle(hc, solver, mhd, regime):
    rng = np.random.default_rng(1234)
    n = 20000
    wl = random_states(rng, n, mhd, regime)
    wr = random_states(rng, n, mhd, regime)
    # a share of near-identical states (smooth flow) and of Bx ~ 0 (degenerate HLLD branches)
    wr[:, : n // 4] = wl[:, : n // 4] * (1 + 1e-6 * rng.normal(size=(wl.shape[0], n // 4)))
    bx = rng.normal(0, 1.0, n)
    bx[n // 2: n // 2 + n // 8] = 0.0
    bx[n // 2 + n // 8: n // 2 + n // 4] *= 1e-9
    dvn = rng.normal(0, 1.0, n)
    dvt = rng.normal(0, 1.0, n)
    dvn[: n // 3] = 0.0
    fo, wo = oracle.riemann(solver, mhd, wl, wr, bx, 5.0 / 3.0, dt=0.01, dx=0.1, dvn=dvn, dvt=dvt)
    fh = np.zeros_like(wl)
    wh = np.zeros(n)
    hc.hc_riemann(oracle.SOLVER[solver], int(mhd), n, _dp(wl), _dp(wr), _dp(bx), _dp(dvn),
                  _dp(dvt), 5.0 / 3.0, 0.01, 0.1, _dp(fh), _dp(wh))
    util.assert_bitwise(fh, fo, "flux %s mhd=%s" % (solver, mhd))
    if mhd:
        util.assert_bitwise(wh, wo, "ct weight")


@pytest.mark.parametrize("solver,mhd", [("hlle", False), ("hlle", True), ("hlld", True),
                                        ("llf", False), ("llf", True), ("roe", False),
                                        ("roe", True)])
@pytest.mark.parametrize("regime", ["subsonic", "supersonic", "mixed"])
def test_isothermal_riemann_matches_oracle(hc, solver, mhd, regime):
    rng = np.random.default_rng(4321)
    n = 20000
    wl = random_states(rng, n, mhd, regime)
    wr = random_states(rng, n, mhd, regime)
    wr[:, : n // 4] = wl[:, : n // 4] * (1 + 1e-6 * rng.normal(size=(wl.shape[0], n // 4)))
    wr[:, n // 4: n // 3] = wl[:, n // 4: n // 3]          # identical states: degenerate waves
    bx = rng.normal(0, 1.0, n)
    bx[n // 2: n // 2 + n // 8] = 0.0
    bx[n // 2 + n // 8: n // 2 + n // 4] *= 1e-9
    fo = oracle.riemann_iso(solver, mhd, wl, wr, bx, 0.7)
    fh = np.zeros_like(wl)
    hc.hc_riemann_iso(oracle.SOLVER[solver], int(mhd), n, _dp(wl), _dp(wr), _dp(bx), 0.7,
                      oracle.DEFAULT_FLOOR, _dp(fh))
    keep = [0, 1, 2, 3] + ([5, 6] if mhd else [])           # slot 4 (energy) is unused
    util.assert_bitwise(fh[keep], fo[keep], "iso flux %s mhd=%s" % (solver, mhd))


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("mhd", [False, True])
def test_characteristic_reconstruction_matches_oracle(hc, order, mhd):
    """xorder = 2c / 3c: eigenvector projections + limiter + back-projection of one cell"""
    rng = np.random.default_rng(99 + order)
    n = 20000
    q = np.zeros((5, 7, n))
    base = random_states(rng, n, True, "mixed")
    for o in range(5):
        q[o] = base*(1.0 + 0.3*rng.normal(size=base.shape))
        q[o][0] = np.abs(q[o][0]) + 1e-3
        q[o][4] = np.abs(q[o][4]) + 1e-3
    q[:, :, : n // 5] = q[2:3, :, : n // 5]                 # uniform stencil
    q[:, 5:7, n // 5: n // 4] = 0.0                         # no transverse field (bt = 0)
    bx = rng.normal(0, 1.0, n)
    bx[n // 2: n // 2 + n // 10] = 0.0
    args = (order, int(mhd), n, _dp(q), _dp(bx), 5.0/3.0, 0.5, 0.5, oracle.DEFAULT_FLOOR,
            oracle.DEFAULT_FLOOR)
    po, mo = np.zeros((7, n)), np.zeros((7, n))
    ph, mh = np.zeros((7, n)), np.zeros((7, n))
    oracle.lib().ao_recon_char(*args, _dp(po), _dp(mo))
    hc.hc_recon_char(*args, _dp(ph), _dp(mh))
    nw = 7 if mhd else 5
    util.assert_bitwise(ph[:nw], po[:nw], "plus")
    util.assert_bitwise(mh[:nw], mo[:nw], "minus")


def test_plm_ppm_match_oracle(hc):
    rng = np.random.default_rng(7)
    n = 50000
    q = [np.ascontiguousarray(rng.normal(0, 1, (7, n)) + (rng.random((7, n)) < 0.1) * 5)
         for _ in range(5)]
    # sprinkle exact extrema / flat regions
    q[2][:, :1000] = q[1][:, :1000]
    q[3][:, 500:1500] = q[2][:, 500:1500]
    ol, orr = oracle.plm(q[1], q[2], q[3], 0.5000000000000001, 0.4999999999999999)
    hl, hr = np.zeros_like(q[2]), np.zeros_like(q[2])
    hc.hc_plm(n, 7, _dp(q[1]), _dp(q[2]), _dp(q[3]), 0.5000000000000001, 0.4999999999999999,
              _dp(hl), _dp(hr))
    util.assert_bitwise(hl, ol, "plm plus")
    util.assert_bitwise(hr, orr, "plm minus")
    ol, orr = oracle.ppm(q[0], q[1], q[2], q[3], q[4], dfloor=-1e300, pfloor=-1e300)
    hc.hc_ppm(n, 7, _dp(q[0]), _dp(q[1]), _dp(q[2]), _dp(q[3]), _dp(q[4]), _dp(hl), _dp(hr))
    util.assert_bitwise(hl, ol, "ppm plus")
    util.assert_bitwise(hr, orr, "ppm minus")
