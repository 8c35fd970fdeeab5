"""CPU: the product's host/device math headers (csrc/exact_math.h, csrc/solver_core.h) compiled with g++ and
driven by loops that mirror the kernels' control flow (tests/hostsim/hostsim.cpp), bit for bit against the oracle.
Covers the exact arithmetic, the border rules and the strip-marching iteration for both lane widths."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from rvdd_release_b200 import synth

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def hostsim():
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    deps = [src] + [os.path.join(ROOT, "rvdd-release_b200", "csrc", f) for f in ("exact_math.h", "solver_core.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out, src, "-lm"],
                       check=True)
    L = C.CDLL(out)
    L.hs_tvl1flow.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int]
    L.hs_tvl1flow.restype = C.c_int
    L.hs_gauss.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_double]
    L.hs_resample.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_float, C.c_float]
    return L


@pytest.mark.parametrize("h,w,iso,nwarps,scalar", [
    (90, 160, "iso3200", 64, 0),      # 4-pixel lanes, several strips
    (90, 160, "iso3200", 7, 1),       # scalar lanes, fewer warps than column segments
    (97, 131, "iso12800", 40, 0),     # odd sizes -> scalar path on every level
    (120, 200, "clean", 2368, 0),     # more warps than rows: 1-row strips (every row is a halo row)
])
def test_solver_math_bit_exact(port, hostsim, h, w, iso, nwarps, scalar):
    I0, I1 = synth.gray_pair(h, w, iso)
    ref, it_ref, _, _, _ = port.tvl1flow_traced(I0, I1, err_mode=1)
    u = np.zeros((2, h, w), np.float32)
    it = np.zeros(32 * 5, np.int32)
    S = hostsim.hs_tvl1flow(I0, I1, u, w, h, nwarps, it, scalar)
    assert np.array_equal(it[:S * 5].reshape(S, 5), it_ref)
    assert np.array_equal(u, ref)


def test_gauss_and_resample_bit_exact(port, hostsim):
    rng = np.random.RandomState(5)
    for ny, nx in [(23, 45), (40, 64), (33, 130)]:
        a = (rng.rand(ny, nx) * 255).astype(np.float32)
        for sigma in (0.8, float(np.float32(0.6 * np.sqrt(3.0)))):
            out = np.empty_like(a)
            hostsim.hs_gauss(a, out, nx, ny, sigma)
            assert np.array_equal(out, port.gaussian(a, sigma))
        z = port.zoom_out(a)
        blurred = port.gaussian(a, float(np.float32(0.6 * np.sqrt(3.0))))
        out = np.empty_like(z)
        hostsim.hs_resample(blurred, nx, ny, out, z.shape[1], z.shape[0], 0.5, 0.5)
        assert np.array_equal(out, z)
