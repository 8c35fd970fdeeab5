"""ctypes binding of libBridge.so (include/rvdd_bridge.h) with torch tensors in and out.

This is the only place the package touches the C ABI.  Tensors cross it as ``data_ptr()`` + sizes + the current
CUDA stream; nothing is staged through the host.  There is NO fallback: if the library is missing or no CUDA
device is usable, the calls raise.
"""
import ctypes as C
import os

import numpy as np
import torch

PKG = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(PKG, "lib", "libBridge.so")
TRACE_SCALES = 16
NWARPS_DEFAULT = 5


class TVL1Params(C.Structure):
    """rvdd_tvl1_params (libBridge.cpp:27-36 defaults)."""
    _fields_ = [("tau", C.c_float), ("lambda_", C.c_float), ("theta", C.c_float), ("nscales", C.c_int),
                ("fscale", C.c_int), ("zfactor", C.c_float), ("nwarps", C.c_int), ("epsilon", C.c_float)]


class BridgeError(RuntimeError):
    pass


_SIGNATURES = {
    # name: (restype, argtypes) -- one entry per symbol declared in include/rvdd_bridge.h
    "tvl1flow": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "rvdd_default_params": (None, [C.POINTER(TVL1Params)]),
    "rvdd_pyramid": (C.c_int, [C.c_int, C.c_int, C.POINTER(TVL1Params), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rvdd_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "rvdd_destroy": (C.c_int, [C.c_void_p]),
    "rvdd_set_groups": (C.c_int, [C.c_void_p, C.c_int]),
    "rvdd_set_watchdog": (C.c_int, [C.c_void_p, C.c_longlong]),
    "rvdd_set_fuse": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rvdd_last_solver_fused": (C.c_int, [C.c_void_p]),
    "rvdd_last_error": (C.c_char_p, []),
    "rvdd_abi_version": (C.c_int, []),
    "rvdd_gray_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "rvdd_tvl1_flow_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_int, C.POINTER(TVL1Params), C.c_void_p, C.c_void_p, C.c_void_p]),
    "rvdd_solver_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvdd_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "rvdd_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "rvdd_profile_scales": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "rvdd_profile_phases": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "rvdd_demosaic_ha_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p]),
    "rvdd_remosaick_gray_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_float,
                                          C.c_float, C.c_void_p]),
    "rvdd_selftest_fastmath": (C.c_int, [C.c_ulonglong, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]),
    "rvdd_debug_level_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rvdd_warp_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
                      + [C.c_longlong] * 8 + [C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "rvdd_upsample2_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "rvdd_flow_and_warp_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_int, C.POINTER(TVL1Params), C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "rvdd_flow_and_warp_host_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.POINTER(TVL1Params), C.c_void_p,
                                                 C.c_void_p, C.c_void_p]),
    "rvdd_flow_and_warp_host_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "rvdd_flow_and_warp_host_submit_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    C.c_void_p, C.c_void_p, C.c_int, C.POINTER(TVL1Params), C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load_library(path=None):
    """dlopen libBridge.so and declare every prototype.  Raises if the library has not been built."""
    path = path or os.environ.get("RVDD_BRIDGE_LIB") or DEFAULT_LIB
    if not os.path.exists(path):
        raise BridgeError("%s not found: build it with `python rvdd-release_b200/build.py` "
                          "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise BridgeError("%s must be a CUDA tensor (no CPU fallback in this path)" % name)
    if t.dtype != torch.float32:
        raise BridgeError("%s must be float32" % name)


class Bridge:
    """One context (device workspace) on the current CUDA device."""

    def __init__(self, libpath=None, groups=0):
        self.lib = load_library(libpath)
        if self.lib.rvdd_abi_version() != 1:
            raise BridgeError("libBridge.so ABI mismatch")
        if not torch.cuda.is_available():
            raise BridgeError("no CUDA device: the alignment path has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.init()
        torch.zeros(1, device=self.device)          # make sure the primary context exists
        h = C.c_void_p()
        self._ck(self.lib.rvdd_create(C.byref(h)))
        self.ctx = h
        if groups:
            self.set_groups(groups)

    def _ck(self, rc):
        if rc != 0:
            raise BridgeError("libBridge: %s (code %d)" % (self.lib.rvdd_last_error().decode(), rc))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.rvdd_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_groups(self, n):
        self._ck(self.lib.rvdd_set_groups(self.ctx, int(n)))

    def set_fuse(self, mode="auto", min_px=-1):
        """Solver instantiation: 'auto' (default), 'never' or 'always' run two iterations per pass on levels of at least
        min_px pixels (rvdd_set_fuse).  Same bits either way."""
        self._ck(self.lib.rvdd_set_fuse(self.ctx, {"auto": 0, "never": 1, "always": 2}[mode], int(min_px)))

    def last_solver_fused(self):
        """True if the last tvl1_flow launch used the two-iterations-per-pass solver instantiation."""
        return self.lib.rvdd_last_solver_fused(self.ctx) == 1

    def set_watchdog(self, ticks):
        """Solver watchdog in SM clock ticks (default 4e9); a launch whose watchdog fires returns NaN flows."""
        self._ck(self.lib.rvdd_set_watchdog(self.ctx, int(ticks)))

    # ------------------------------------------------------------------ geometry
    def pyramid(self, nx, ny, params=None):
        nxs, nys = (C.c_int * TRACE_SCALES)(), (C.c_int * TRACE_SCALES)()
        S = self.lib.rvdd_pyramid(nx, ny, C.byref(params) if params else None, nxs, nys)
        return [(nxs[s], nys[s]) for s in range(S)]

    # ------------------------------------------------------------------ device entry points
    def gray(self, frames):
        """[n, h, w, c] packed HWC frames -> [n, h, w] gray (library.py:162-170)."""
        _check_cuda_f32(frames, "frames")
        frames = frames.contiguous()
        n, h, w, c = frames.shape
        out = torch.empty((n, h, w), dtype=torch.float32, device=frames.device)
        self._ck(self.lib.rvdd_gray_dev(frames.data_ptr(), out.data_ptr(), n, h, w, c, _stream_ptr(frames.device)))
        return out

    def tvl1_flow(self, gray, src, tgt, params=None, trace=False, check=False):
        """Batched TV-L1.  gray [nframes, ny, nx]; pair k: I0 = gray[tgt[k]], I1 = gray[src[k]].

        Returns flow [npairs, 2, ny, nx] (ch 0 = x-displacement) and, with trace=True, the per-(scale, warp)
        inner-iteration counts [npairs, 16, nwarps] (int32, on the device).  check=True synchronises and raises
        if the solver watchdog fired."""
        _check_cuda_f32(gray, "gray")
        gray = gray.contiguous()
        nframes, ny, nx = gray.shape
        src = np.ascontiguousarray(src, dtype=np.int32)
        tgt = np.ascontiguousarray(tgt, dtype=np.int32)
        k = int(src.size)
        flow = torch.empty((k, 2, ny, nx), dtype=torch.float32, device=gray.device)
        nw = params.nwarps if params else NWARPS_DEFAULT
        iters = torch.zeros((k, TRACE_SCALES, nw), dtype=torch.int32, device=gray.device) if trace else None
        self._ck(self.lib.rvdd_tvl1_flow_dev(self.ctx, gray.data_ptr(), nframes, nx, ny, src.ctypes.data,
                                             tgt.ctypes.data, k, C.byref(params) if params else None,
                                             flow.data_ptr(), iters.data_ptr() if trace else None,
                                             _stream_ptr(gray.device)))
        if check:
            self.check(gray.device)
        return (flow, iters) if trace else flow

    def profile(self, enable=True):
        """Bracket every solver launch with CUDA events (roofline report of bench.py)."""
        self._ck(self.lib.rvdd_profile(self.ctx, 1 if enable else 0))

    def profile_read(self, cap=4096):
        """-> list of solver-kernel durations in ms since the last read."""
        buf = (C.c_float * cap)()
        n = self.lib.rvdd_profile_read(self.ctx, buf, cap)
        if n < 0:
            raise BridgeError("rvdd_profile_read failed")
        return [buf[i] for i in range(n)]

    def profile_scales(self):
        """-> mean ms a pair of the last profiled launch spent at each pyramid level (index 0 = finest)."""
        buf = (C.c_float * TRACE_SCALES)()
        n = self.lib.rvdd_profile_scales(self.ctx, buf, TRACE_SCALES)
        return [buf[i] for i in range(max(n, 0))]

    def profile_phases(self):
        """-> [(ms in the warp-constants phases, ms in the iteration loops)] per pyramid level of the last profiled launch."""
        buf = (C.c_float * (2 * TRACE_SCALES))()
        n = self.lib.rvdd_profile_phases(self.ctx, buf, 2 * TRACE_SCALES)
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(max(n, 0))]

    def debug_level(self, pair, which, level, nx, ny):
        """Pyramid level of the last tvl1_flow call (test hook, rvdd_debug_level_dev)."""
        out = torch.empty((ny, nx), dtype=torch.float32, device=self.device)
        self._ck(self.lib.rvdd_debug_level_dev(self.ctx, pair, which, level, out.data_ptr(), _stream_ptr(self.device)))
        return out

    def check(self, device=None):
        self._ck(self.lib.rvdd_solver_status(self.ctx, _stream_ptr(device or self.device)))

    def warp(self, x, flow, interp="bicubic", flow_mul=1.0, want_mask=True, out=None):
        """flow_utils.warp on the device.  x [B, C, H, W] (any strides), flow [B, 2, H, W] or [B, 2, H/2, W/2]."""
        _check_cuda_f32(x, "x")
        _check_cuda_f32(flow, "flow")
        if interp not in ("bicubic", "bilinear"):
            raise BridgeError("interp must be 'bicubic' or 'bilinear' (got %r)" % (interp,))
        B, Cc, H, W = x.shape
        flow = flow.contiguous()
        if flow.shape[0] != B or flow.shape[1] != 2:
            raise BridgeError("flow must be [B, 2, h, w]")
        fh, fw = int(flow.shape[2]), int(flow.shape[3])
        if out is None:
            out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
        elif out.data_ptr() == x.data_ptr():
            raise BridgeError("warp: out must not alias x")
        mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device) if want_mask else None
        xs, os_ = x.stride(), out.stride()
        self._ck(self.lib.rvdd_warp_dev(x.data_ptr(), flow.data_ptr(), out.data_ptr(),
                                        mask.data_ptr() if want_mask else None, B, Cc, H, W,
                                        xs[0], xs[1], xs[2], xs[3], os_[0], os_[1], os_[2], os_[3], fh, fw,
                                        float(flow_mul), 1 if interp == "bicubic" else 0, _stream_ptr(x.device)))
        return out, mask

    def upsample2(self, t, mul=1.0):
        """upsample_factor_2 on [..., C, H, W] (flow_utils.py:159-174)."""
        _check_cuda_f32(t, "tensor")
        t = t.contiguous()
        *rem, c, h, w = t.shape
        planes = int(np.prod(rem, dtype=np.int64)) * c if rem else c
        out = torch.empty((*rem, c, 2 * h, 2 * w), dtype=torch.float32, device=t.device)
        self._ck(self.lib.rvdd_upsample2_dev(t.data_ptr(), out.data_ptr(), planes, h, w, float(mul),
                                             _stream_ptr(t.device)))
        return out

    def demosaic(self, x, pattern="gbrg"):
        """HamiltonAdam(pattern).forward (util/Hamilton_Adam_demo.py:249-289): [B, 4k, H, W] packed raw -> [B, 3k, 2H, 2W]."""
        _check_cuda_f32(x, "x")
        x = x.contiguous()
        B, c4, H, W = x.shape
        if c4 % 4:
            raise BridgeError("demosaic: the channel count must be a multiple of 4 (packed Bayer frames)")
        out = torch.empty((B, 3 * (c4 // 4), 2 * H, 2 * W), dtype=torch.float32, device=x.device)
        self._ck(self.lib.rvdd_demosaic_ha_dev(x.data_ptr(), out.data_ptr(), B * (c4 // 4), H, W, pattern.encode(),
                                               _stream_ptr(x.device)))
        return out

    def remosaick_gray(self, rgb, pattern="gbrg", add=1.0, mul=0.5):
        """mean over the 4 packed channels of (remosaick(rgb) + add) * mul: [B, 3, 2H, 2W] -> [B, H, W]."""
        _check_cuda_f32(rgb, "rgb")
        rgb = rgb.contiguous()
        B, c, H2, W2 = rgb.shape
        if c != 3 or H2 % 2 or W2 % 2:
            raise BridgeError("remosaick_gray: expected [B, 3, 2H, 2W]")
        out = torch.empty((B, H2 // 2, W2 // 2), dtype=torch.float32, device=rgb.device)
        self._ck(self.lib.rvdd_remosaick_gray_dev(rgb.data_ptr(), out.data_ptr(), B, H2 // 2, W2 // 2, pattern.encode(),
                                                  float(add), float(mul), _stream_ptr(rgb.device)))
        return out

    # ------------------------------------------------------------------ host entry point (end to end)
    def flow_and_warp_host(self, frames, src, tgt, params=None, want_warp=True, trace=False, flow_out=None,
                           warped_out=None):
        """compute_flow_and_warp for a batch of pairs with HOST buffers (rvdd_flow_and_warp_host).

        frames: float32 [nframes, h, w, c] numpy array or CPU tensor (pinned memory makes the copies async).
        Returns (flow [npairs, h, w, 2], warped [npairs, h, w, c] or None, iters or None) as CPU tensors."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames, dtype=np.float32))
        if frames.is_cuda or frames.dtype != torch.float32:
            raise BridgeError("frames must be a float32 host array/tensor")
        frames = frames.contiguous()
        n, h, w, c = frames.shape
        src = np.ascontiguousarray(src, dtype=np.int32)
        tgt = np.ascontiguousarray(tgt, dtype=np.int32)
        k = int(src.size)
        flow = flow_out if flow_out is not None else torch.empty((k, h, w, 2), dtype=torch.float32)
        warped = None
        if want_warp:
            warped = warped_out if warped_out is not None else torch.empty((k, h, w, c), dtype=torch.float32)
        nw = params.nwarps if params else NWARPS_DEFAULT
        iters = torch.zeros((k, TRACE_SCALES, nw), dtype=torch.int32) if trace else None
        self._ck(self.lib.rvdd_flow_and_warp_host(self.ctx, frames.data_ptr(), n, h, w, c, src.ctypes.data,
                                                  tgt.ctypes.data, k, C.byref(params) if params else None,
                                                  flow.data_ptr(), warped.data_ptr() if want_warp else None,
                                                  iters.data_ptr() if trace else None))
        return flow, warped, iters


    def submit_host(self, slot, frames, src, tgt, flow_out, warped_out=None, params=None, discard_warp=False):
        """Asynchronous half of flow_and_warp_host on staging slot 0 / 1 (rvdd_flow_and_warp_host_submit_ex).  `frames`,
        `flow_out`, `warped_out` are CPU float32 tensors (pin them to let the copies overlap) that must stay alive and
        untouched until wait_host(slot).  discard_warp=True (with warped_out=None) warps on the device without downloading
        the result -- the reference's gen_warp=False case (base_dataset.py:178-189)."""
        n, h, w, c = frames.shape
        src = np.ascontiguousarray(src, dtype=np.int32)
        tgt = np.ascontiguousarray(tgt, dtype=np.int32)
        mode = 1 if warped_out is not None else (2 if discard_warp else 0)
        self._ck(self.lib.rvdd_flow_and_warp_host_submit_ex(
            self.ctx, int(slot), frames.data_ptr(), n, h, w, c, src.ctypes.data, tgt.ctypes.data, int(src.size),
            C.byref(params) if params else None, flow_out.data_ptr(),
            warped_out.data_ptr() if warped_out is not None else None, None, mode))

    def wait_host(self, slot):
        self._ck(self.lib.rvdd_flow_and_warp_host_wait(self.ctx, int(slot)))


_DEFAULT = None


def default_bridge():
    """Process-wide Bridge on the current device (created on first use)."""
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Bridge()
    return _DEFAULT
