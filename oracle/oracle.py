"""ctypes access to the CPU oracles.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (rvdd-release_b200/) never does.

Two libraries, both built by oracle/Makefile:

* ``PortLib``  -- oracle/_build/libtvl1_port.so, our C restatement (oracle/tvl1_port.c), with iteration traces.
* ``RefLib``   -- oracle/_ref/libref_{serial,omp}.so, the UNMODIFIED reference sources
  (/root/reference/libBridge.cpp + 3rdparty/tvl1flow/*.c) compiled where they lie.  Every stage is reachable
  because the reference exports all of its functions.

The torch/numpy restatement of the reference warp lives in oracle/warp_ref.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(quiet=True):
    """(Re)build the port always, and the reference objects when /root/reference is present."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class PortParams(C.Structure):
    _fields_ = [("tau", C.c_float), ("lambda_", C.c_float), ("theta", C.c_float),
                ("nscales", C.c_int), ("fscale", C.c_int), ("zfactor", C.c_float),
                ("nwarps", C.c_int), ("epsilon", C.c_float), ("err_mode", C.c_int)]


class PortLib:
    """Our restatement.  Shapes are (ny, nx) float32 arrays."""

    def __init__(self, path=None):
        path = path or os.path.join(HERE, "_build", "libtvl1_port.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.port_tvl1flow.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.port_tvl1flow.restype = None
        L.port_tvl1flow_traced.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _i32p, _f32p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.port_tvl1flow_traced.restype = C.c_int
        L.port_clamp_nscales.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
        L.port_clamp_nscales.restype = C.c_int
        L.port_pyramid_sizes.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, _i32p, _i32p]
        L.port_normalize.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int]
        L.port_gaussian.argtypes = [_f32p, C.c_int, C.c_int, C.c_double]
        L.port_zoom_out.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_float]
        L.port_zoom_in.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.port_zoom_size.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float]
        L.port_bicubic_warp.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.port_centered_gradient.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.port_forward_gradient.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.port_divergence.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.port_warp_constants.argtypes = [_f32p] * 10 + [C.c_int, C.c_int]
        L.port_iteration.argtypes = [_f32p] * 10 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _f32p,
                                                    C.POINTER(C.c_double)]
        L.port_iteration.restype = C.c_float
        L.port_solve_scale.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.POINTER(PortParams),
                                       C.c_int, C.c_void_p]
        L.port_default_params.argtypes = [C.POINTER(PortParams)]
        L.port_multiscale.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.POINTER(PortParams), C.c_void_p]

    def nscales(self, nx, ny, zfactor=0.5, req=100):
        return self.L.port_clamp_nscales(nx, ny, zfactor, req)

    def pyramid_sizes(self, nx, ny, zfactor=0.5):
        S = self.nscales(nx, ny, zfactor)
        a, b = np.zeros(S, np.int32), np.zeros(S, np.int32)
        self.L.port_pyramid_sizes(nx, ny, zfactor, S, a, b)
        return list(zip(a.tolist(), b.tolist()))

    def tvl1flow(self, I0, I1):
        ny, nx = I0.shape
        u = np.zeros((2, ny, nx), np.float32)
        self.L.port_tvl1flow(_c(I0), _c(I1), u, nx, ny)
        return u

    def multiscale(self, I0, I1, tau=0.25, lam=0.15, theta=0.3, nscales=100, fscale=0, zfactor=0.5, nwarps=5, epsilon=0.01):
        """Dual_TVL1_optic_flow_multiscale (tvl1flow_lib.c:343-472) with explicit parameters; nscales is clamped the
        way libBridge.cpp:131-138 clamps it."""
        ny, nx = I0.shape
        P = PortParams()
        self.L.port_default_params(C.byref(P))
        P.tau, P.lambda_, P.theta, P.fscale, P.zfactor, P.nwarps, P.epsilon = tau, lam, theta, fscale, zfactor, nwarps, epsilon
        P.nscales = self.L.port_clamp_nscales(nx, ny, zfactor, nscales)
        if P.nscales < P.fscale:
            P.fscale = P.nscales
        u = np.zeros((2, ny, nx), np.float32)
        self.L.port_multiscale(_c(I0), _c(I1), u[0], u[1], nx, ny, C.byref(P), None)
        return u

    def tvl1flow_traced(self, I0, I1, err_mode=0, err_cap=0):
        """-> (flow (2,ny,nx), iters (S, nwarps), last_err (S, nwarps), err_f32, err_f64)."""
        ny, nx = I0.shape
        u = np.zeros((2, ny, nx), np.float32)
        iters = np.zeros(32 * 5, np.int32)
        last = np.zeros(32 * 5, np.float32)
        ef = np.zeros(max(err_cap, 1), np.float32)
        ed = np.zeros(max(err_cap, 1), np.float64)
        n = C.c_int(0)
        S = self.L.port_tvl1flow_traced(_c(I0), _c(I1), u, nx, ny, err_mode, iters, last,
                                        ef.ctypes.data if err_cap else None, ed.ctypes.data if err_cap else None,
                                        err_cap, C.addressof(n))
        return u, iters[:S * 5].reshape(S, 5), last[:S * 5].reshape(S, 5), ef[:n.value], ed[:n.value]

    def normalize(self, a, b):
        an, bn = np.empty_like(_c(a)), np.empty_like(_c(b))
        self.L.port_normalize(_c(a), _c(b), an, bn, a.size)
        return an, bn

    def gaussian(self, img, sigma):
        out = _c(img).copy()
        self.L.port_gaussian(out, img.shape[1], img.shape[0], float(sigma))
        return out

    def zoom_out(self, img, factor=0.5):
        ny, nx = img.shape
        nxx, nyy = C.c_int(), C.c_int()
        self.L.port_zoom_size(nx, ny, C.byref(nxx), C.byref(nyy), factor)
        out = np.empty((nyy.value, nxx.value), np.float32)
        self.L.port_zoom_out(_c(img), out, nx, ny, factor)
        return out

    def zoom_in(self, img, nxx, nyy):
        out = np.empty((nyy, nxx), np.float32)
        self.L.port_zoom_in(_c(img), out, img.shape[1], img.shape[0], nxx, nyy)
        return out

    def bicubic_warp(self, img, u, v, border_out=True):
        out = np.empty_like(_c(img))
        self.L.port_bicubic_warp(_c(img), _c(u), _c(v), out, img.shape[1], img.shape[0], int(border_out))
        return out

    def centered_gradient(self, img):
        dx, dy = np.empty_like(_c(img)), np.empty_like(_c(img))
        self.L.port_centered_gradient(_c(img), dx, dy, img.shape[1], img.shape[0])
        return dx, dy

    def forward_gradient(self, img):
        dx, dy = np.empty_like(_c(img)), np.empty_like(_c(img))
        self.L.port_forward_gradient(_c(img), dx, dy, img.shape[1], img.shape[0])
        return dx, dy

    def divergence(self, a, b):
        out = np.empty_like(_c(a))
        self.L.port_divergence(_c(a), _c(b), out, a.shape[1], a.shape[0])
        return out

    def warp_constants(self, I0, I1, u1, u2):
        """-> (I1wx, I1wy, grad, rho_c) for one warp, as tvl1flow_lib.c:143-159."""
        ny, nx = I0.shape
        I1x, I1y = self.centered_gradient(I1)
        gx, gy, g2, rc = (np.empty((ny, nx), np.float32) for _ in range(4))
        self.L.port_warp_constants(_c(I0), _c(I1), I1x, I1y, _c(u1), _c(u2), gx, gy, g2, rc, nx, ny)
        return gx, gy, g2, rc

    def iteration(self, u1, u2, p, gx, gy, g2, rc, theta=np.float32(0.3), l_t=None, taut=None):
        """One in-place primal-dual iteration on copies -> (u1, u2, p(4), err_f32, sum_f64)."""
        ny, nx = u1.shape
        f = np.float32
        l_t = f(f(0.15) * f(0.3)) if l_t is None else l_t
        taut = f(f(0.25) / f(0.3)) if taut is None else taut
        u1, u2 = _c(u1).copy(), _c(u2).copy()
        p = [_c(q).copy() for q in p]
        scratch = np.empty(3 * nx * ny, np.float32)
        wide = C.c_double()
        e = self.L.port_iteration(u1, u2, p[0], p[1], p[2], p[3], _c(gx), _c(gy), _c(g2), _c(rc), nx, ny,
                                  theta, l_t, taut, scratch, C.byref(wide))
        return u1, u2, p, e, wide.value


class RefLib:
    """The unmodified reference, compiled by oracle/Makefile into oracle/_ref/."""

    def __init__(self, kind="serial"):
        path = os.path.join(HERE, "_ref", "libref_%s.so" % kind)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (needs /root/reference to build)")
        L = self.L = C.CDLL(path)
        self.kind = kind
        L.tvl1flow.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]           # libBridge.cpp:44
        L.tvl1flow.restype = None
        L.Dual_TVL1_optic_flow_multiscale.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float,
                                                      C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int,
                                                      C.c_float, C.c_bool]        # tvl1flow_lib.c:343
        L.Dual_TVL1_optic_flow.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float,
                                           C.c_float, C.c_int, C.c_float, C.c_bool]   # tvl1flow_lib.c:91
        L.image_normalization.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int]   # tvl1flow_lib.c:301
        L.gaussian.argtypes = [_f32p, C.c_int, C.c_int, C.c_double]              # mask.c:214
        L.zoom_out.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_float]        # zoom.c:41
        L.zoom_in.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]  # zoom.c:85
        L.zoom_size.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float]
        L.bicubic_interpolation_warp.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_bool]
        L.centered_gradient.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]   # mask.c:149
        L.forward_gradient.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]    # mask.c:98
        L.divergence.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]          # mask.c:40

    def tvl1flow(self, I0, I1):
        ny, nx = I0.shape
        u = np.zeros((2, ny, nx), np.float32)
        self.L.tvl1flow(_c(I0), _c(I1), u, nx, ny)
        return u

    def multiscale(self, I0, I1, tau=0.25, lam=0.15, theta=0.3, nscales=100, fscale=0, zfactor=0.5, nwarps=5, epsilon=0.01):
        """The reference's Dual_TVL1_optic_flow_multiscale itself (tvl1flow_lib.c:343), nscales clamped by the caller
        exactly as libBridge.cpp:131-138 does."""
        import math
        ny, nx = I0.shape
        N = np.float32(1 + math.log(math.hypot(nx, ny) / 16.0) / math.log(1 / float(np.float32(zfactor))))
        if N < nscales:
            nscales = int(N)
        if nscales < fscale:
            fscale = nscales
        u = np.zeros((2, ny, nx), np.float32)
        self.L.Dual_TVL1_optic_flow_multiscale(_c(I0), _c(I1), u[0], u[1], nx, ny, tau, lam, theta, nscales, fscale, zfactor,
                                               nwarps, epsilon, False)
        return u

    def solve_scale(self, I0, I1, u1, u2, tau=0.25, lam=0.15, theta=0.3, warps=5, eps=0.01):
        ny, nx = I0.shape
        u1, u2 = _c(u1).copy(), _c(u2).copy()
        self.L.Dual_TVL1_optic_flow(_c(I0), _c(I1), u1, u2, nx, ny, tau, lam, theta, warps, eps, False)
        return u1, u2

    def normalize(self, a, b):
        an, bn = np.empty_like(_c(a)), np.empty_like(_c(b))
        self.L.image_normalization(_c(a), _c(b), an, bn, a.size)
        return an, bn

    def gaussian(self, img, sigma):
        out = _c(img).copy()
        self.L.gaussian(out, img.shape[1], img.shape[0], float(sigma))
        return out

    def zoom_out(self, img, factor=0.5):
        ny, nx = img.shape
        nxx, nyy = C.c_int(), C.c_int()
        self.L.zoom_size(nx, ny, C.byref(nxx), C.byref(nyy), factor)
        out = np.empty((nyy.value, nxx.value), np.float32)
        self.L.zoom_out(_c(img), out, nx, ny, factor)
        return out

    def zoom_in(self, img, nxx, nyy):
        out = np.empty((nyy, nxx), np.float32)
        self.L.zoom_in(_c(img), out, img.shape[1], img.shape[0], nxx, nyy)
        return out

    def bicubic_warp(self, img, u, v, border_out=True):
        out = np.empty_like(_c(img))
        self.L.bicubic_interpolation_warp(_c(img), _c(u), _c(v), out, img.shape[1], img.shape[0], bool(border_out))
        return out

    def centered_gradient(self, img):
        dx, dy = np.empty_like(_c(img)), np.empty_like(_c(img))
        self.L.centered_gradient(_c(img), dx, dy, img.shape[1], img.shape[0])
        return dx, dy

    def forward_gradient(self, img):
        dx, dy = np.empty_like(_c(img)), np.empty_like(_c(img))
        self.L.forward_gradient(_c(img), dx, dy, img.shape[1], img.shape[0])
        return dx, dy

    def divergence(self, a, b):
        out = np.empty_like(_c(a))
        self.L.divergence(_c(a), _c(b), out, a.shape[1], a.shape[0])
        return out
