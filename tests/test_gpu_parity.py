"""GPU (-m gpu): the CUDA path, called through the C ABI of libBridge.so, against the oracle on the same seeded
inputs.  Integer-like quantities (iteration counts per scale and warp) must coincide; float results are compared
bit for bit where the arithmetic is rounding-exact by construction and with the north-star tolerances otherwise
(mean end-point error <= 0.01 px for flows, <= 1e-4 relative for warped pixels)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import warp_ref
from rvdd_release_b200 import synth

pytestmark = pytest.mark.gpu
EPE_TOL = 0.01          # px, BASELINE.json north_star
WARP_RTOL = 1e-4        # relative, BASELINE.json north_star


def epe(a, b):
    return float(np.sqrt(((a - b) ** 2).sum(axis=-3)).mean())


def run_pairs(bridge, pairs, groups=0, trace=True):
    """pairs: list of (I0, I1) numpy gray images of one size -> flows [k,2,h,w], iters [k,S,5]."""
    k = len(pairs)
    gray = torch.from_numpy(np.stack([im for pr in pairs for im in pr])).cuda()
    tgt, src = np.arange(0, 2 * k, 2), np.arange(1, 2 * k, 2)
    bridge.set_groups(groups)
    flow, iters = bridge.tvl1_flow(gray, src, tgt, trace=True, check=True)
    bridge.set_groups(0)
    S = len(bridge.pyramid(gray.shape[2], gray.shape[1]))
    return flow.cpu().numpy(), iters.cpu().numpy()[:, :S, :]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tvl1_*.npz"))))
def test_flow_matches_reference_golden(bridge, port, path):
    g = np.load(path)
    flow, iters = run_pairs(bridge, [(g["I0"], g["I1"])])
    _, it_ref, _, _, _ = port.tvl1flow_traced(g["I0"], g["I1"], err_mode=0)
    assert np.array_equal(iters[0], it_ref), (iters[0], it_ref)
    assert epe(flow[0], g["flow"]) <= EPE_TOL
    assert np.array_equal(flow[0], g["flow"])       # rounding-exact arithmetic: identical bits


def test_pyramid_levels_bit_exact(bridge, port):
    I0, I1 = synth.gray_pair(90, 160, "iso3200")
    run_pairs(bridge, [(I0, I1)])
    a, b = port.normalize(I0, I1)
    a, b = port.gaussian(a, 0.8), port.gaussian(b, 0.8)
    for lvl, (nx, ny) in enumerate(bridge.pyramid(160, 90)):
        assert np.array_equal(bridge.debug_level(0, 0, lvl, nx, ny).cpu().numpy(), a), lvl
        assert np.array_equal(bridge.debug_level(0, 1, lvl, nx, ny).cpu().numpy(), b), lvl
        a, b = port.zoom_out(a), port.zoom_out(b)


@pytest.mark.parametrize("h,w,iso", [(97, 131, "iso12800"), (180, 320, "iso3200"), (135, 240, "clean"), (64, 260, "iso3200")])
def test_flow_various_sizes(bridge, port, h, w, iso):
    I0, I1 = synth.gray_pair(h, w, iso)
    ref, it_ref, _, _, _ = port.tvl1flow_traced(I0, I1, err_mode=0)
    flow, iters = run_pairs(bridge, [(I0, I1)])
    assert np.array_equal(iters[0], it_ref)
    assert epe(flow[0], ref) <= EPE_TOL
    assert np.array_equal(flow[0], ref)


def test_batch_groups_and_determinism(bridge, port):
    """A batch of different pairs gives the same bits whatever the number of solver groups, and twice in a row."""
    seq = synth.sequence(6, 90, 160, "iso3200").numpy().mean(axis=3, dtype=np.float32)
    pairs = [(seq[t], seq[t - 1]) for t in range(1, 6)] + [(seq[t], seq[t + 1]) for t in range(0, 3)]
    refs = [port.tvl1flow_traced(a, b, err_mode=0) for a, b in pairs]
    base = None
    for groups in (1, 3, 8, 0):
        flow, iters = run_pairs(bridge, pairs, groups=groups)
        if base is None:
            base = flow
        assert np.array_equal(flow, base), groups
        for k, r in enumerate(refs):
            assert np.array_equal(iters[k], r[1]), (groups, k)
            assert epe(flow[k], r[0]) <= EPE_TOL
    flow2, _ = run_pairs(bridge, pairs, groups=8)
    assert np.array_equal(flow2, base)


def test_flow_1280x720_against_oracle(bridge, port):
    """The headline geometry: one noisy 1280x720 pair, 7 scales, against the oracle (a few seconds of CPU)."""
    I0, I1 = synth.gray_pair(720, 1280, "iso3200")
    ref, it_ref, _, _, _ = port.tvl1flow_traced(I0, I1, err_mode=0)
    flow, iters = run_pairs(bridge, [(I0, I1)])
    assert iters.shape[1] == 7
    mism = int((iters[0] != it_ref).sum())
    e = epe(flow[0], ref)
    print("1280x720: EPE %.6f px, iteration-count mismatches %d, iters/scale %s" % (e, mism, iters[0].sum(1).tolist()))
    assert e <= EPE_TOL
    assert mism == 0 and np.array_equal(flow[0], ref)


@pytest.mark.parametrize("h,w,iso", [(1080, 1920, "iso12800"), (360, 640, "iso3200")])
def test_flow_other_pipeline_geometries(bridge, port, h, w, iso):
    """1920x1080 (the packed-raw grid of the 3840x2160 config, 8 scales incl. odd widths 15x9) and 640x360 (the
    true REDS packed-raw geometry, 6 scales) against the oracle."""
    I0, I1 = synth.gray_pair(h, w, iso)
    ref, it_ref, _, _, _ = port.tvl1flow_traced(I0, I1, err_mode=0)
    flow, iters = run_pairs(bridge, [(I0, I1)])
    assert iters.shape[1] == it_ref.shape[0]
    assert epe(flow[0], ref) <= EPE_TOL
    assert np.array_equal(iters[0], it_ref) and np.array_equal(flow[0], ref)


def test_flow_3840x2160_properties(bridge):
    """Largest configured geometry (9 scales): no oracle run (minutes of CPU); size-independent properties instead --
    finite output, the synthetic motion is recovered, the result does not depend on the batch it was computed in."""
    seq = synth.sequence(2, 2160, 3840, "iso3200").numpy().mean(axis=3, dtype=np.float32)
    flow, iters = run_pairs(bridge, [(seq[1], seq[0])])
    assert iters.shape[1] == 9 and np.isfinite(flow).all()
    inner = flow[0][:, 64:-64, 64:-64]
    assert -4.5 < inner[0].mean() < -0.5 and 0.0 < inner[1].mean() < 3.0     # same sign as test_flow_sign_convention
    both, _ = run_pairs(bridge, [(seq[1], seq[0]), (seq[0], seq[1])], groups=2)
    assert np.array_equal(both[0], flow[0])


def test_zero_motion_and_linearity_properties(bridge):
    """Size-independent properties at full size: identical frames give exactly zero flow; the flow of a pair does
    not depend on what else is in the batch."""
    I0, _ = synth.gray_pair(720, 1280, "iso3200")
    J0, J1 = synth.gray_pair(720, 1280, "iso12800", t=2)
    flow, _ = run_pairs(bridge, [(I0, I0), (J0, J1)])
    assert np.count_nonzero(flow[0]) == 0
    solo, _ = run_pairs(bridge, [(J0, J1)])
    assert np.array_equal(solo[0], flow[1])


def test_dropin_cppbridge_host_call(libpath, bridge, port):
    """library.CPPbridge(libpath).TVL1_flow(Im1, Im2) with numpy HWC inputs, exactly as the reference is used."""
    from rvdd_release_b200.library import CPPbridge
    seq = synth.sequence(2, 72, 128, "iso3200").numpy()
    flow = CPPbridge(libpath).TVL1_flow(seq[1], seq[0])
    assert flow.shape == (72, 128, 2) and flow.dtype == np.float32
    ref = port.tvl1flow(np.mean(seq[1], axis=2), np.mean(seq[0], axis=2))
    assert np.array_equal(flow.transpose(2, 0, 1), ref)


def test_fast_exact_primitives_selftest(bridge):
    """The straight-line division / hypot used in the iteration are bit-identical to IEEE division and the
    double-precision square root on ~4e8 pseudo-random operand pairs (or they declare themselves unsure)."""
    import ctypes as C
    cnt = (C.c_ulonglong * 6)()
    assert bridge.lib.rvdd_selftest_fastmath(12345, 592, 3000, cnt) == 0
    hyp_n, hyp_bad, hyp_mis, div_n, div_rej, div_mis = list(cnt)
    print("hypot: %d trials, %.3f%% fallbacks, %d mismatches; div: %d trials, %.3f%% rejected, %d mismatches"
          % (hyp_n, 100.0 * hyp_bad / hyp_n, hyp_mis, div_n, 100.0 * div_rej / div_n, div_mis))
    assert hyp_n > 4e8 and hyp_mis == 0 and div_mis == 0
    assert hyp_bad / hyp_n < 0.2 and div_rej / div_n < 0.2


def test_gray_kernel_matches_numpy(bridge):
    seq = synth.sequence(3, 40, 64, "iso12800")
    g = bridge.gray(seq.cuda()).cpu().numpy()
    assert np.array_equal(g, np.mean(seq.numpy(), axis=3))


# ------------------------------------------------------------------------------------------------ warp

def rel_err(a, b):
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "warp_*.npz"))))
def test_warp_matches_reference_golden(bridge, path):
    from rvdd_release_b200 import flow_utils
    g = np.load(path)
    x, flow = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["flow"]).cuda()
    yb, m = flow_utils.warp(x, flow, "bicubic")
    yl, _ = flow_utils.warp(x, flow, "bilinear")
    assert rel_err(yb.cpu().numpy(), g["bicubic"]) <= WARP_RTOL
    assert rel_err(yl.cpu().numpy(), g["bilinear"]) <= WARP_RTOL
    assert np.array_equal(m.cpu().numpy(), g["mask"]) and m.device == x.device
    up = flow_utils.upsample_factor_2(torch.from_numpy(g["half_flow"]).cuda(), 2)
    assert rel_err(up.cpu().numpy(), g["up2x2"]) <= WARP_RTOL
    # fused x2 upsampling of a half-resolution flow inside the warp (recurrent_model.py:129 + :151)
    yh, _ = flow_utils.warp(x, torch.from_numpy(g["half_flow"]).cuda(), "bicubic", flow_mul=2.0)
    assert rel_err(yh.cpu().numpy(), g["bicubic_half"]) <= 5 * WARP_RTOL


@pytest.mark.parametrize("C,H,W", [(3, 720, 1280), (48, 180, 320), (4, 360, 640)])
def test_warp_against_oracle_sizes(bridge, C, H, W):
    from rvdd_release_b200 import flow_utils
    g = torch.Generator().manual_seed(C)
    x = torch.randn(1, C, H, W, generator=g)
    flow = 4.0 * torch.randn(1, 2, H, W, generator=g)
    ref, mref = warp_ref.warp(x, flow, "bicubic")
    y, m = flow_utils.warp(x.cuda(), flow.cuda(), "bicubic")
    assert rel_err(y.cpu().numpy(), ref.numpy()) <= WARP_RTOL
    assert np.array_equal(m.cpu().numpy(), mref.numpy())


@pytest.mark.parametrize("C,H,W,interp", [(3, 360, 640, "bicubic"), (48, 90, 200, "bicubic"), (5, 77, 131, "bilinear"),
                                            (11, 64, 96, "bicubic")])
def test_warp_smooth_flow_staged_path(bridge, C, H, W, interp):
    """Smooth flows (what TV-L1 produces) take the shared-memory staged path of the tiled kernel, including tiles at
    the image border where taps are clamped; large constant shifts push whole tiles out of the image."""
    from rvdd_release_b200 import flow_utils
    g = torch.Generator().manual_seed(H)
    x = torch.randn(2, C, H, W, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    f0 = torch.stack((2.5 + 1.5 * torch.sin(yy / 17.0), -1.5 + torch.cos(xx / 23.0)), 0)
    f1 = torch.stack((-40.0 + 0.01 * xx, 25.0 + 0.02 * yy), 0)             # far outside on two sides
    flow = torch.stack((f0, f1), 0)
    ref, mref = warp_ref.warp(x, flow, interp)
    y, m = flow_utils.warp(x.cuda(), flow.cuda(), interp)
    assert rel_err(y.cpu().numpy(), ref.numpy()) <= WARP_RTOL
    assert np.array_equal(m.cpu().numpy(), mref.numpy())


def test_warp_channel_slice_and_identity(bridge):
    """The feature-recurrence call site warps a channel slice of a bigger tensor (recurrent_model.py:295-297);
    zero flow is the identity up to the normalise / un-normalise round trip of the sampling grid, which the
    reference has too (flow_utils.py:93-94 + grid_sampler_unnormalize)."""
    from rvdd_release_b200 import flow_utils
    g = torch.Generator().manual_seed(7)
    feat = torch.randn(2, 96, 36, 52, generator=g).cuda()
    flow = (3.0 * torch.randn(2, 2, 36, 52, generator=g)).cuda()
    sl = feat[:, 48:96]
    y, _ = flow_utils.warp(sl, flow, "bicubic")
    ref, _ = warp_ref.warp(sl.cpu().contiguous(), flow.cpu(), "bicubic")
    assert rel_err(y.cpu().numpy(), ref.numpy()) <= WARP_RTOL
    ident, m = flow_utils.warp(sl, torch.zeros_like(flow), "bicubic")
    assert rel_err(ident.cpu().numpy(), sl.cpu().numpy()) <= 1e-5 and bool((m == 1).all())
    iref, _ = warp_ref.warp(sl.cpu().contiguous(), torch.zeros_like(flow).cpu(), "bicubic")
    assert rel_err(ident.cpu().numpy(), iref.numpy()) <= WARP_RTOL
    with pytest.raises(Exception):
        flow_utils.warp(sl.cpu(), flow.cpu(), "bicubic")            # no CPU fallback


def test_single_warp_and_compute_flow_and_warp(libpath, bridge, port):
    """numpy-level API of util/flow_utils.py:105-156 and the host-buffer batch entry point."""
    from rvdd_release_b200 import flow_utils
    seq = synth.sequence(3, 72, 128, "iso3200").numpy()
    warped, mask, flow = flow_utils.compute_flow_and_warp(seq[0], seq[1])          # img1 = source, img2 = target
    ref_flow = port.tvl1flow(np.mean(seq[1], axis=2), np.mean(seq[0], axis=2)).transpose(1, 2, 0)
    assert np.array_equal(flow, ref_flow)
    ref_w, _ = warp_ref.warp(torch.from_numpy(seq[0].transpose(2, 0, 1)[None].copy()),
                             torch.from_numpy(ref_flow.transpose(2, 0, 1)[None].copy()), "bicubic")
    assert rel_err(warped, ref_w[0].numpy().transpose(1, 2, 0)) <= WARP_RTOL
    # batch, host buffers in / out
    f, w, it = bridge.flow_and_warp_host(seq, src=[0, 2], tgt=[1, 1], trace=True)
    assert np.array_equal(f[0].numpy(), ref_flow)
    assert rel_err(w[0].numpy(), ref_w[0].numpy().transpose(1, 2, 0)) <= WARP_RTOL
    ref2 = port.tvl1flow(np.mean(seq[1], axis=2), np.mean(seq[2], axis=2)).transpose(1, 2, 0)
    assert np.array_equal(f[1].numpy(), ref2)
    assert int(it[0].sum()) > 0


def test_online_flow_from_denoised_frame(bridge, port):
    """validate.py:16-38 (--val_flow_from_denoised) without the GPU -> CPU -> GPU round trip: remosaick + gray +
    TV-L1 on the device against the same steps done the reference's way on the host."""
    from rvdd_release_b200 import flow_utils
    seq = synth.sequence(2, 60, 96, "iso3200")
    noisy = (seq[1] / 4095.0 * 2 - 1).permute(2, 0, 1)[None].contiguous()              # T: 2x - 1, CHW
    g = torch.Generator().manual_seed(3)
    den = torch.rand(1, 3, 120, 192, generator=g) * 2 - 1                               # a "denoised" RGB frame
    den[:, 1, 0::2, 0::2] = (seq[0][:, :, 0] / 4095.0 * 2 - 1)                          # make it resemble frame t-1
    den[:, 2, 0::2, 1::2] = (seq[0][:, :, 1] / 4095.0 * 2 - 1)
    den[:, 0, 1::2, 0::2] = (seq[0][:, :, 2] / 4095.0 * 2 - 1)
    den[:, 1, 1::2, 1::2] = (seq[0][:, :, 3] / 4095.0 * 2 - 1)
    flow = flow_utils.compute_flows_from_denoised(den.cuda(), noisy.cuda())
    assert tuple(flow.shape) == (1, 1, 2, 60, 96)
    # host restatement: remosaick (Hamilton_Adam_demo.py:237-246), singleiT (library.py:67), mean of 4, tvl1flow
    y = torch.stack((den[:, 1, 0::2, 0::2], den[:, 2, 0::2, 1::2], den[:, 0, 1::2, 0::2], den[:, 1, 1::2, 1::2]), 1)
    img1 = ((y[0] + 1.) / 2.).permute(1, 2, 0).numpy()
    img2 = ((noisy[0] + 1.) / 2.).permute(1, 2, 0).numpy()
    ref = port.tvl1flow(np.mean(img2, axis=2), np.mean(img1, axis=2))
    assert np.array_equal(flow[0, 0].cpu().numpy(), ref)


def test_pipelined_host_submissions(bridge, port):
    """rvdd_flow_and_warp_host_submit / _wait: two batches in flight on the two staging slots give the same bits as the
    blocking call; resubmitting a busy slot is refused."""
    from rvdd_release_b200.bridge import BridgeError
    seqs = [synth.sequence(3, 64, 96, iso, noise_seed=7 * i) for i, iso in enumerate(("iso3200", "iso12800", "clean"))]
    src, tgt = [0, 1], [1, 2]
    want = [bridge.flow_and_warp_host(s, src, tgt)[:2] for s in seqs]
    frames = [s.pin_memory() for s in seqs]
    flows = [torch.empty((2, 64, 96, 2)).pin_memory() for _ in seqs]
    warps = [torch.empty((2, 64, 96, 4)).pin_memory() for _ in seqs]
    bridge.submit_host(0, frames[0], src, tgt, flows[0], warps[0])
    bridge.submit_host(1, frames[1], src, tgt, flows[1], warps[1])
    with pytest.raises(BridgeError):
        bridge.submit_host(0, frames[2], src, tgt, flows[2], warps[2])
    bridge.wait_host(0)
    bridge.submit_host(0, frames[2], src, tgt, flows[2], warps[2])
    bridge.wait_host(1)
    bridge.wait_host(0)
    for k in range(3):
        assert torch.equal(flows[k], want[k][0]) and torch.equal(warps[k], want[k][1])
    g = np.mean(seqs[1].numpy(), axis=3)
    assert np.array_equal(flows[1][0].numpy().transpose(2, 0, 1), port.tvl1flow(g[1], g[0]))


def test_precompute_driver_on_gpu(tmp_path, bridge, port):
    """The offline flow cache end to end on the GPU: frame TIFFs in, <from>_<to>.tif flows out (pipelined batches),
    identical to what the reference's createWarpedInputData would write (base_dataset.py:134-191)."""
    from rvdd_release_b200 import flowio, precompute
    videos = []
    for v, iso in enumerate(("iso3200", "iso12800", "clean")):
        d = tmp_path / "noisy" / ("seq%d" % v)
        d.mkdir(parents=True)
        seq = synth.sequence(4, 48, 80, iso, noise_seed=11 * v).numpy()
        for f in range(4):
            flowio.write_tif(str(d / ("%03d.tif" % f)), seq[f])
        videos.append((seq, str(d)))
    listed = precompute.list_videos(str(tmp_path / "noisy"))
    files = precompute.precompute_dataset(listed, str(tmp_path / "flow"), None, 2, 1, max_pairs_per_batch=4)
    assert len(files) == 3 * 6
    for v, (seq, _) in enumerate(videos):
        g = np.mean(seq, axis=3)
        past = flowio.read_tif(str(tmp_path / "flow" / ("seq%d" % v) / "001_002.tif"))        # source 001 -> target 002
        assert np.array_equal(past.transpose(2, 0, 1), port.tvl1flow(g[2], g[1]))
        fut = flowio.read_tif(str(tmp_path / "flow" / ("seq%d" % v) / "003_002.tif"))         # future source 003 -> target 002
        assert np.array_equal(fut.transpose(2, 0, 1), port.tvl1flow(g[2], g[3]))
    assert precompute.precompute_dataset(listed, str(tmp_path / "flow"), None, 2, 1) == []    # resume: nothing to do


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "demosaic_*.npz"))))
def test_demosaic_matches_reference_golden(bridge, path):
    """Hamilton-Adams kernel against the reference's own module output (util/Hamilton_Adam_demo.py:249-289)."""
    from rvdd_release_b200.hamilton_adam import HamiltonAdam
    g = np.load(path)
    ha = HamiltonAdam(str(g["pattern"]))
    y = ha(torch.from_numpy(g["x"]).cuda())
    assert tuple(y.shape) == g["y"].shape
    assert float((y.cpu() - torch.from_numpy(g["y"])).abs().max()) <= 1e-6
    r = ha.remosaick(torch.from_numpy(g["y"][:, :3]).cuda())
    assert torch.equal(r.cpu(), torch.from_numpy(g["remosaick"]))


@pytest.mark.parametrize("H,W,pattern", [(720, 1280, "gbrg"), (33, 47, "gbrg"), (1, 1, "gbrg"), (2, 70, "rggb"), (45, 3, "bggr"),
                                         (16, 32, "grbg")])
def test_demosaic_against_oracle_sizes(bridge, H, W, pattern):
    """Full pipeline size (1440 x 2560 output), ragged tiles, a single Bayer cell: bit-level agreement with the oracle."""
    from oracle import demosaic_ref
    seq = synth.sequence(2, H, W, "iso3200", noise_seed=5)
    x = (2.0 * (seq / 4095.0) - 1.0).permute(0, 3, 1, 2).contiguous()                  # [2, 4, H, W] in [-1, 1]
    y = bridge.demosaic(x.cuda(), pattern).cpu().numpy()
    ref = demosaic_ref.hamilton_adam(x.numpy(), pattern)
    assert y.shape == ref.shape == (2, 3, 2 * H, 2 * W)
    assert np.abs(y - ref).max() <= 1e-6
    # stacked frames [1, 8, H, W] are demosaicked frame by frame (recurrent_model.py:126 passes 4k channels)
    y2 = bridge.demosaic(x.reshape(1, 8, H, W).cuda(), pattern).cpu().numpy()
    assert np.array_equal(y2.reshape(2, 3, 2 * H, 2 * W), y)
    # remosaick + (x + 1) / 2 + mean of 4, fused, against the oracle's remosaick and numpy
    gray = bridge.remosaick_gray(torch.from_numpy(ref).cuda(), pattern, add=1.0, mul=0.5).cpu().numpy()
    packed = (demosaic_ref.remosaick(ref, pattern) + np.float32(1)) / np.float32(2)
    want = (((packed[:, 0] + packed[:, 1]) + packed[:, 2]) + packed[:, 3]) * np.float32(0.25)
    assert np.array_equal(gray, want)
    if pattern == "gbrg":     # the CFA samples survive demosaic -> remosaick untouched
        assert np.array_equal(demosaic_ref.remosaick(y, pattern), x.numpy())


@pytest.mark.parametrize("interp", ["bicubic", "bilinear"])
def test_warp_channel_innermost_frames(bridge, interp):
    """Frames in their on-disk (H, W, 4) layout, viewed as [B, 4, H, W] (single_warp, flow_utils.py:105-122): the
    128-bit gather kernel, with a half-resolution flow, a mask, and both channel-innermost and plane outputs."""
    from rvdd_release_b200 import flow_utils
    g = torch.Generator().manual_seed(11)
    B, H, W = 3, 70, 102
    hwc = torch.randn(B, H, W, 4, generator=g)
    x = hwc.permute(0, 3, 1, 2)                                   # strides (H*W*4, 1, W*4, 4)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    flow = torch.stack((3.0 * torch.sin(yy / 9.0) + 0.3 * torch.randn(H, W, generator=g),
                        -2.0 + 4.0 * torch.cos(xx / 13.0)), 0)[None].repeat(B, 1, 1, 1).contiguous()
    flow[1] += 60.0                                               # one image sampled far outside
    ref, mref = warp_ref.warp(x.contiguous(), flow, interp)
    y, m = flow_utils.warp(x.cuda(), flow.cuda(), interp)
    assert rel_err(y.cpu().numpy(), ref.numpy()) <= WARP_RTOL and np.array_equal(m.cpu().numpy(), mref.numpy())
    out_hwc = torch.empty(B, H, W, 4, device="cuda")
    bridge.warp(x.cuda(), flow.cuda(), interp, want_mask=False, out=out_hwc.permute(0, 3, 1, 2))
    assert torch.equal(out_hwc.permute(0, 3, 1, 2), y)
    half = (0.5 * flow[:, :, ::2, ::2]).contiguous()
    rh, _ = warp_ref.warp(x.contiguous(), warp_ref.upsample_factor_2(half, 2), interp)
    yh, _ = flow_utils.warp(x.cuda(), half.cuda(), interp, flow_mul=2.0)
    assert rel_err(yh.cpu().numpy(), rh.numpy()) <= 5 * WARP_RTOL


@pytest.mark.parametrize("kw", [dict(zfactor=0.7, nwarps=3, tau=0.2, lam=0.1, theta=0.25, epsilon=0.02),
                                dict(zfactor=0.5, nwarps=2, nscales=3, lam=0.3, theta=0.4, epsilon=0.005),
                                dict(zfactor=0.35, nwarps=4, fscale=1),
                                dict(zfactor=0.8, nscales=4, nwarps=1)])
def test_flow_with_other_parameters(bridge, port, kw):
    """Same parameters as the reference, whatever they are (rvdd_tvl1_params): zoom factors other than 0.5 take the
    generic Gaussian + bicubic resampling path and other tap radii, fscale > 0 skips the finest level."""
    from rvdd_release_b200.bridge import TVL1Params
    I0, I1 = synth.gray_pair(72, 110, "iso3200")
    want = port.multiscale(I0, I1, **kw)
    p = TVL1Params()
    bridge.lib.rvdd_default_params(p)
    for k, v in kw.items():
        setattr(p, "lambda_" if k == "lam" else k, v)
    gray = torch.from_numpy(np.stack([I0, I1])).cuda()
    flow = bridge.tvl1_flow(gray, [1], [0], params=p, check=True)[0].cpu().numpy()
    assert epe(flow, want) <= EPE_TOL
    assert np.array_equal(flow, want)


def test_demosaic_argument_errors(bridge):
    """Error behaviour of the new entry points: status code + message, never a crash or a silent fallback."""
    from rvdd_release_b200.bridge import BridgeError
    from rvdd_release_b200.hamilton_adam import HamiltonAdam
    x = torch.zeros(1, 4, 8, 8, device="cuda")
    with pytest.raises(BridgeError):
        bridge.demosaic(x, "rgbg")                    # red and blue must sit on a diagonal
    with pytest.raises(BridgeError):
        bridge.demosaic(torch.zeros(1, 6, 8, 8, device="cuda"), "gbrg")
    with pytest.raises(BridgeError):
        bridge.demosaic(x.cpu(), "gbrg")              # no CPU fallback
    with pytest.raises(ValueError):
        HamiltonAdam("xyzw")
    with pytest.raises(BridgeError):
        bridge.remosaick_gray(torch.zeros(1, 3, 7, 8, device="cuda"))
    assert bridge.demosaic(torch.zeros(0, 4, 8, 8, device="cuda")).shape == (0, 3, 16, 16)


@pytest.mark.parametrize("interp", ["bicubic", "bilinear"])
def test_warp_hwc4_tma_staged_variant(bridge, monkeypatch, interp):
    """RVDD_WARP_TMA=1: the (H, W, 4) warp with its tap box fetched by the TMA unit (measured slower than the L1 gathers and
    therefore opt-in) against the oracle, including border tiles and a tile whose flow does not fit the box."""
    monkeypatch.setenv("RVDD_WARP_TMA", "1")
    B, H, W = 3, 75, 150
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, 4, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    f0 = torch.stack((2.5 + 1.5 * torch.sin(yy / 17.0), -1.5 + torch.cos(xx / 23.0)), 0)
    f1 = torch.stack((-40.0 + 0.01 * xx, 25.0 + 0.02 * yy), 0)             # far outside on two sides
    f2 = 6.0 * torch.randn(2, H, W, generator=g)                            # rough: boxes overflow -> direct gathers
    flow = torch.stack((f0, f1, f2), 0)
    xc = x.permute(0, 3, 1, 2)
    ref, mref = warp_ref.warp(xc.contiguous(), flow, interp)
    out = torch.empty(B, H, W, 4, device="cuda").permute(0, 3, 1, 2)
    y, m = bridge.warp(xc.cuda(), flow.cuda(), interp, out=out)
    assert rel_err(y.cpu().numpy(), ref.numpy()) <= WARP_RTOL
    assert np.array_equal(m.cpu().numpy(), mref.numpy())
    monkeypatch.delenv("RVDD_WARP_TMA")
    y2, _ = bridge.warp(xc.cuda(), flow.cuda(), interp)
    assert rel_err(y2.cpu().numpy(), ref.numpy()) <= WARP_RTOL


def test_kernel_variants_are_bit_equal(bridge, monkeypatch):
    """The shipped fast paths against the plain ones they replaced, bit for bit: demosaic interior tiles vs the general
    border path (RVDD_DEMOSAIC_GENERAL=1), (H, W, 4) bicubic warp two rows per thread vs one pixel per thread
    (RVDD_WARP_HWC_1PX=1).  Sizes with interior tiles, ragged edges, odd widths and a misaligned view."""
    g = torch.Generator().manual_seed(11)
    for (B, H, W) in [(2, 100, 136), (1, 37, 65), (1, 64, 70)]:
        x = (torch.rand(B, 4, H, W, generator=g) * 2 - 1).cuda()
        flat = torch.zeros(B * 4 * H * W + 1, device="cuda")
        flat[1:] = x.reshape(-1)
        views = [x, flat[1:].view(B, 4, H, W)]                 # the second: rows not 8-byte aligned -> scalar staging
        for pat in ("gbrg", "rggb"):
            fast = [bridge.demosaic(v, pat).cpu() for v in views]
            monkeypatch.setenv("RVDD_DEMOSAIC_GENERAL", "1")
            general = bridge.demosaic(x, pat).cpu()
            monkeypatch.delenv("RVDD_DEMOSAIC_GENERAL")
            assert torch.equal(fast[0], general) and torch.equal(fast[1], general)
    for (B, H, W) in [(2, 75, 150), (1, 33, 31), (1, 2, 40)]:
        x = torch.randn(B, H, W, 4, generator=g).cuda().permute(0, 3, 1, 2)
        for flow in (2.5 + 0.05 * torch.randn(B, 2, H, W, generator=g), 5.0 * torch.randn(B, 2, H, W, generator=g)):
            two, m2 = bridge.warp(x, flow.cuda(), "bicubic")
            monkeypatch.setenv("RVDD_WARP_HWC_1PX", "1")
            one, m1 = bridge.warp(x, flow.cuda(), "bicubic")
            monkeypatch.delenv("RVDD_WARP_HWC_1PX")
            assert torch.equal(two, one) and torch.equal(m2, m1)
