"""GPU (-m gpu): the LITERAL drop-in boundary.  The reference binds one symbol of ./build/libBridge.so with ctypes
(library.py:145-148) and calls it with host float buffers (library.py:150-175); these tests perform that binding and that
call sequence verbatim -- not through rvdd_release_b200.bridge -- on 4-, 3- and 1-channel inputs and compare the result
with the oracle bit for bit.  Also here: the largest configured geometry (3840x2160) against a golden vector made with
the compiled reference, the per-device context of `tvl1flow`, and the cross-stream ordering of the shared workspace."""
import ctypes
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from rvdd_release_b200 import synth

pytestmark = pytest.mark.gpu


def rgb2gray(rgb):
    """skimage.color.rgb2gray (the import at library.py:14; skimage is not in this image): rgb @ [0.2125, 0.7154, 0.0721]
    in the array's own float type."""
    return rgb @ np.array([0.2125, 0.7154, 0.0721], dtype=rgb.dtype)


class ReferenceCPPbridge(object):
    """library.py:143-175 of the reference, copied call for call (this is the code a user of the reference runs)."""

    def __init__(self, libpath):
        self.libBridge = ctypes.cdll.LoadLibrary(libpath)
        self.libBridge.tvl1flow.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        self.libBridge.tvl1flow.restype = None

    def TVL1_flow(self, Im1, Im2):
        h, w = Im1.shape[:2]
        h1, w1 = Im2.shape[:2]
        assert h1 == h and w1 == w
        I1 = np.zeros(h * w, dtype=ctypes.c_float)
        I2 = np.zeros(h * w, dtype=ctypes.c_float)
        flow = np.zeros(2 * h * w, dtype=ctypes.c_float)
        if Im1.shape[2] == 3:
            I1[:] = rgb2gray(Im1).flatten()[:]
            I2[:] = rgb2gray(Im2).flatten()[:]
        elif Im1.shape[2] == 4:
            I1[:] = np.mean(Im1, axis=2).flatten()[:]
            I2[:] = np.mean(Im2, axis=2).flatten()[:]
        elif Im1.shape[2] == 1:
            I1[:] = Im1.flatten()[:]
            I2[:] = Im2.flatten()[:]
        floatp = ctypes.POINTER(ctypes.c_float)
        self.libBridge.tvl1flow(I1.ctypes.data_as(floatp), I2.ctypes.data_as(floatp), flow.ctypes.data_as(floatp),
                                ctypes.c_int(w), ctypes.c_int(h))
        self.gray = (I1.reshape(h, w), I2.reshape(h, w))
        return flow.reshape(2, h, w).transpose(1, 2, 0)


def _installed_lib(libpath):
    """The copy build.py puts where the reference looks for it: ./build/libBridge.so (flow_utils.py:129,145)."""
    root = os.path.dirname(os.path.dirname(os.path.dirname(libpath)))
    inst = os.path.join(root, "build", "libBridge.so")
    return inst if os.path.exists(inst) else libpath


@pytest.mark.parametrize("channels", [4, 3, 1])
@pytest.mark.parametrize("h,w,iso", [(72, 128, "iso3200"), (97, 131, "iso12800")])
def test_literal_tvl1flow_symbol(libpath, bridge, port, channels, h, w, iso):
    seq = synth.sequence(2, h, w, iso).numpy()
    if channels == 3:
        seq = np.ascontiguousarray(seq[..., :3])
    elif channels == 1:
        seq = np.ascontiguousarray(seq.mean(axis=3, keepdims=True, dtype=np.float32))
    cpp = ReferenceCPPbridge(_installed_lib(libpath))
    flow = cpp.TVL1_flow(seq[1], seq[0])                     # Im1 = target, Im2 = source (flow_utils.py:149)
    assert flow.shape == (h, w, 2) and flow.dtype == np.float32
    ref, it_ref, _, _, _ = port.tvl1flow_traced(cpp.gray[0], cpp.gray[1], err_mode=0)
    assert int(it_ref.sum()) > 0 and np.count_nonzero(flow) > 0
    assert np.array_equal(flow.transpose(2, 0, 1), ref)
    # our own CPPbridge mirror gives the same bits on the same inputs (4-ch takes the gray-on-GPU path)
    from rvdd_release_b200.library import CPPbridge
    mine = CPPbridge(libpath).TVL1_flow(seq[1], seq[0])
    assert np.array_equal(mine, flow)


def test_literal_tvl1flow_1280x720(libpath, bridge, port):
    """The headline geometry through the one symbol the reference binds, one pair per call."""
    I0, I1 = synth.gray_pair(720, 1280, "iso3200")
    lib = ctypes.cdll.LoadLibrary(_installed_lib(libpath))
    lib.tvl1flow.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.tvl1flow.restype = None
    u = np.zeros(2 * 720 * 1280, dtype=ctypes.c_float)
    floatp = ctypes.POINTER(ctypes.c_float)
    lib.tvl1flow(I0.ctypes.data_as(floatp), I1.ctypes.data_as(floatp), u.ctypes.data_as(floatp), ctypes.c_int(1280),
                 ctypes.c_int(720))
    ref = port.tvl1flow(I0, I1)
    assert np.array_equal(u.reshape(2, 720, 1280), ref)


def test_tvl1flow_follows_the_current_device(libpath, bridge, port):
    """`tvl1flow` keeps one context per device and runs on the device current in the calling thread."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    I0, I1 = synth.gray_pair(72, 128, "iso3200")
    ref = port.tvl1flow(I0, I1)
    lib = ctypes.cdll.LoadLibrary(libpath)
    lib.tvl1flow.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2
    lib.tvl1flow.restype = None
    for dev in (1, 0, 1):
        with torch.cuda.device(dev):
            before = torch.cuda.memory_stats(dev)  # noqa: F841  (touches the device: makes it current for the runtime)
            torch.zeros(1, device="cuda:%d" % dev)
            u = np.zeros((2, 72, 128), np.float32)
            lib.tvl1flow(I0.ctypes.data, I1.ctypes.data, u.ctypes.data, 128, 72)
            assert np.array_equal(u, ref), dev


def test_flow_3840x2160_against_reference_golden(bridge):
    """Config-5 geometry, 9 scales: flow and iteration counts against tests/golden/large_tvl1_2160x3840_exact.npz (made by
    tests/golden/make_large_golden.py with the compiled reference; inputs regenerated here and hash-checked first)."""
    g = np.load(os.path.join(GOLDEN, "large_tvl1_2160x3840_exact.npz"))
    h, w = int(g["h"]), int(g["w"])
    I0, I1 = synth.exact_gray_pair(h, w)

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    assert sha(I0) == str(g["sha_I0"]) and sha(I1) == str(g["sha_I1"]), "synthetic inputs are not bit-reproducible here"
    gray = torch.from_numpy(np.stack([I0, I1])).cuda()
    flow, iters = bridge.tvl1_flow(gray, [1], [0], trace=True, check=True)
    flow, iters = flow[0].cpu().numpy(), iters[0].cpu().numpy()
    S = g["iters"].shape[0]
    assert S == 9 and np.array_equal(iters[:S], g["iters"]), (iters[:S].sum(1), g["iters"].sum(1))
    sub = flow[:, ::24, ::24]
    epe = float(np.sqrt(((sub - g["flow_sub"]) ** 2).sum(0)).mean())
    assert epe <= 0.01, epe
    assert np.array_equal(sub, g["flow_sub"])
    assert sha(flow) == str(g["sha_flow"])


def test_workspace_is_ordered_across_streams(bridge, port):
    """Two calls on different streams share the context's workspace: the second must wait for the first's kernels."""
    seq = synth.sequence(4, 180, 320, "iso3200").numpy().mean(axis=3, dtype=np.float32)
    gray = torch.from_numpy(seq).cuda()
    refs = [port.tvl1flow(seq[t], seq[t - 1]) for t in (1, 2, 3)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        for k, st in enumerate((s1, s2, s1)):
            with torch.cuda.stream(st):
                outs.append((k, bridge.tvl1_flow(gray, [k], [k + 1])))
    torch.cuda.synchronize()
    bridge.check()
    for k, f in outs:
        assert np.array_equal(f[0].cpu().numpy(), refs[k]), k


def test_watchdog_poisons_the_result(bridge):
    """A launch whose watchdog fires must not return plausible-looking stale values: the flows come back as NaN and the
    status call reports it.  (A 1-tick limit makes every multi-CTA barrier time out.)"""
    from rvdd_release_b200 import bridge as B
    seq = synth.sequence(2, 360, 640, "iso3200").numpy().mean(axis=3, dtype=np.float32)
    gray = torch.from_numpy(seq).cuda()
    b = B.Bridge()
    try:
        b.set_watchdog(1)
        flow = b.tvl1_flow(gray, [0], [1])
        torch.cuda.synchronize()
        assert bool(torch.isnan(flow).all())
        with pytest.raises(B.BridgeError):
            b.check()
        b.set_watchdog(4000000000)
        flow = b.tvl1_flow(gray, [0], [1], check=True)
        assert bool(torch.isfinite(flow).all())
    finally:
        b.close()


def test_both_solver_instantiations_are_bit_exact(libpath, port):
    """The solver exists twice: one iteration per pass, and two iterations per pass on big levels (speculative exact stop with a
    one-iteration replay).  The library picks per launch ('auto': from the previous launch's iteration counts); forced either
    way, and with the fused pass on EVERY level (min_px = 0: tiny strips, odd pyramids), the reference's bits and iteration
    counts must come out."""
    from rvdd_release_b200 import bridge as B
    b = B.Bridge(libpath)
    try:
        cases = [(180, 320, "iso3200"), (97, 132, "iso12800"), (360, 640, "clean"), (720, 1280, "iso3200")]
        refs = []
        for h, w, iso in cases:
            I0, I1 = synth.gray_pair(h, w, iso)
            ref, it_ref, _, _, _ = port.tvl1flow_traced(I0, I1, err_mode=0)
            refs.append((torch.from_numpy(np.stack([I0, I1])).cuda(), ref, it_ref))
        for mode, min_px in (("always", 0), ("always", 600000), ("never", -1), ("auto", -1), ("auto", -1)):
            b.set_fuse(mode, min_px)
            for (h, w, iso), (gray, ref, it_ref) in zip(cases, refs):
                flow, iters = b.tvl1_flow(gray, [1], [0], trace=True, check=True)
                assert np.array_equal(iters[0, :it_ref.shape[0]].cpu().numpy(), it_ref), (mode, min_px, h, w)
                assert np.array_equal(flow[0].cpu().numpy(), ref), (mode, min_px, h, w)
    finally:
        b.close()


def test_flow_edge_cases_and_errors(bridge, port):
    """The reference's failure modes on this path kill the process (mask.c:229-232 abort when the Gaussian is wider than the
    image, xmalloc.c:15-17 exit); the library reports them instead.  Plus: an empty batch, the smallest image the presmoothing
    kernel accepts, a constant pair (normalisation with max == min copies, tvl1flow_lib.c:327-334) and bad pair indices."""
    from rvdd_release_b200 import bridge as B
    dev = "cuda"
    # empty batch: nothing to do, empty result
    g = torch.zeros(2, 32, 48, device=dev)
    out = bridge.tvl1_flow(g, [], [])
    assert tuple(out.shape) == (0, 2, 32, 48)
    # smaller than the 9-tap presmoothing kernel: an error, not an abort
    with pytest.raises(B.BridgeError):
        bridge.tvl1_flow(torch.zeros(2, 4, 64, device=dev), [0], [1])
    # pair index out of range
    with pytest.raises(B.BridgeError):
        bridge.tvl1_flow(g, [0], [2])
    # a zoom factor so close to 1 that the reference would use more than 16 scales
    p = B.TVL1Params()
    bridge.lib.rvdd_default_params(ctypes.byref(p))
    p.zfactor = 0.97
    with pytest.raises(B.BridgeError):
        bridge.tvl1_flow(torch.rand(2, 360, 640, device=dev), [0], [1], params=p)
    # smallest sizes with more than one scale / exactly one scale, against the oracle
    for h, w in ((17, 23), (12, 16), (9, 40)):
        I0, I1 = synth.gray_pair(h, w, "iso3200")
        ref, it_ref, _, _, _ = port.tvl1flow_traced(I0, I1, err_mode=0)
        flow, iters = bridge.tvl1_flow(torch.from_numpy(np.stack([I0, I1])).cuda(), [1], [0], trace=True, check=True)
        assert np.array_equal(iters[0, :it_ref.shape[0]].cpu().numpy(), it_ref), (h, w)
        assert np.array_equal(flow[0].cpu().numpy(), ref), (h, w)
    # constant images: den == 0 -> copied unnormalised, zero flow
    c = np.full((40, 56), 7.0, np.float32)
    ref = port.tvl1flow(c, c)
    flow = bridge.tvl1_flow(torch.from_numpy(np.stack([c, c])).cuda(), [1], [0], check=True)
    assert np.array_equal(flow[0].cpu().numpy(), ref) and np.count_nonzero(ref) == 0


def test_auto_kernel_choice_follows_the_iteration_counts(libpath):
    """'auto' launches the two-iterations-per-pass instantiation only when the previous launch's inner loops were long (noisy
    frames) and the batch gives every warp long strips; clean frames and small batches stay on the plain kernel.  Either way
    the result of a launch does not depend on which instantiation computed it."""
    from rvdd_release_b200 import bridge as B
    b = B.Bridge(libpath)
    try:
        noisy = b.gray(synth.sequence(17, 720, 1280, "iso3200", device="cuda"))
        clean = b.gray(synth.sequence(17, 720, 1280, "clean", device="cuda"))
        src, tgt = np.arange(16), np.arange(1, 17)
        b.set_fuse("auto")
        first = b.tvl1_flow(noisy, src, tgt)
        assert not b.last_solver_fused()                       # nothing known yet: plain kernel
        torch.cuda.synchronize()
        second = b.tvl1_flow(noisy, src, tgt)
        assert b.last_solver_fused()                           # ~19 inner iterations per warp on the finest level
        assert torch.equal(first, second)
        torch.cuda.synchronize()
        b.tvl1_flow(noisy, src[:2], tgt[:2])
        assert not b.last_solver_fused()                       # 2 pairs on the whole GPU: strips too short to fuse
        torch.cuda.synchronize()
        b.tvl1_flow(clean, src, tgt)                           # (still decided from the noisy launch before it)
        torch.cuda.synchronize()
        b.tvl1_flow(clean, src, tgt)
        assert not b.last_solver_fused()                       # ~1 inner iteration per warp: plain kernel
    finally:
        b.close()
