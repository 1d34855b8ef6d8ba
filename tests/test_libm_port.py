"""tgx_libm.h (the exp/log the CUDA forward-backward kernels use) must be bit-identical to the
system libm — the functions Rust's f64::exp / f64::ln call in the reference.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libm_port_shim.so")


@pytest.fixture(scope="module")
def shim():
    src = os.path.join(HERE, "libm_port_shim.c")
    hdr = os.path.join(HERE, "..", "tokengeex_b200", "csrc", "tgx_libm.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        flags = ["-O2", "-ffp-contract=off", "-fPIC", "-shared"]
        if "fma" in open("/proc/cpuinfo").read().split("flags", 1)[-1].split("\n", 1)[0].split():
            flags.append("-mfma")
        subprocess.check_call(["gcc", *flags, "-o", SO, src, "-lm"])
    L = C.CDLL(SO)
    L.shim_exp.restype = C.c_double
    L.shim_exp.argtypes = [C.c_double]
    L.shim_log.restype = C.c_double
    L.shim_log.argtypes = [C.c_double]
    L.shim_compare.restype = C.c_uint64
    L.shim_compare.argtypes = [C.POINTER(C.c_double), C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    return L


def compare(L, xs, which):
    xs = np.ascontiguousarray(xs, np.float64)
    bad = C.c_double(0.0)
    n = L.shim_compare(xs.ctypes.data_as(C.POINTER(C.c_double)), xs.size, which, C.byref(bad))
    return int(n), float(bad.value)


def test_exp_matches_libm_bitwise(shim):
    rng = np.random.default_rng(1)
    sets = [
        -rng.random(2_000_000) * 50.0,               # log_sum_exp: exp(vmin - vmax), d in [-50, 0]
        -rng.random(1_000_000) * 800.0,              # exp(a + s + b - z): down into the subnormal range
        rng.normal(0, 1e-3, 200_000),                # near zero, both signs
        rng.uniform(-745.2, -700.0, 200_000),        # subnormal results (specialcase)
        rng.uniform(-1100.0, 720.0, 200_000),        # underflow / overflow edges
        np.array([0.0, -0.0, 1.0, -1.0, -50.0, -708.3964185322641, -745.1332191019411, -745.2, -1e300,
                  709.782712893384, 710.0, 1e-300, -1e-300, np.inf, -np.inf, 2.0 ** -54, -2.0 ** -55, 512.0,
                  -512.0, 1023.9, -1023.9, 1024.0, -1024.0]),
    ]
    for xs in sets:
        n, bad = compare(shim, xs, 0)
        assert n == 0, (n, bad)
    assert np.isnan(shim.shim_exp(float("nan")))


def test_log_matches_libm_bitwise(shim):
    rng = np.random.default_rng(2)
    d = -rng.random(2_000_000) * 50.0
    sets = [
        np.exp(d) + 1.0,                             # exactly the log_sum_exp arguments, (1, 2]
        rng.uniform(0.9, 1.1, 1_000_000),            # both sides of the near-1 polynomial window
        np.exp(rng.uniform(-700, 700, 1_000_000)),   # whole normal range (digamma's ln on device later)
        rng.uniform(0.0, 5e-308, 100_000),           # subnormals
        np.array([1.0, 0.9375, 1.0 + float.fromhex("0x1.09p-4"), np.nextafter(0.9375, 0), np.nextafter(1.0 + float.fromhex("0x1.09p-4"), 2), 2.0,
                  0.5, 1e-310, 5e-324, np.inf, 0.0, -0.0, 1e308]),
    ]
    for xs in sets:
        n, bad = compare(shim, xs, 1)
        assert n == 0, (n, bad)
    assert np.isnan(shim.shim_log(-1.0)) and np.isnan(shim.shim_log(float("nan")))


def test_softplus_matches_libm_bitwise(shim):
    """log(exp(d) + 1.0), the inner term of log_sum_exp (src/lattice.rs:321-333), as one specialised function."""
    rng = np.random.default_rng(3)
    sets = [
        -rng.random(3_000_000) * 50.0,                # the domain log_sum_exp produces
        -rng.random(500_000) * 3.0,                   # both sides of the near-1 window of log (d ~ -2.74)
        -np.exp(rng.uniform(-45, 4.2, 500_000)),      # tiny |d| up to ~-66: the |d| < 2^-54 select, the fallback beyond -60
        np.array([0.0, -0.0, -2.0 ** -54, -2.0 ** -55, -1e-300, -36.7, -36.8, -37.5, -50.0, -60.0, -60.000001, -700.0,
                  -2.7436, -2.7437, 1.0, 5.0, -np.inf]),
    ]
    for xs in sets:
        n, bad = compare(shim, xs, 2)
        assert n == 0, (n, bad)
    shim.shim_softplus.restype = C.c_double
    shim.shim_softplus.argtypes = [C.c_double]
    assert np.isnan(shim.shim_softplus(float("nan")))
