"""CPU: the oracle (oracle/tvl1_port.c, oracle/warp_ref.py) against the golden vectors produced by the reference
itself (tests/golden/make_golden.py) and, where oracle/_ref is present, against the compiled reference directly."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import warp_ref
from rvdd_release_b200 import synth


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tvl1_*.npz"))))
def test_port_matches_reference_golden(port, path):
    g = np.load(path)
    I0, I1 = g["I0"], g["I1"]
    h, w = I0.shape
    assert np.array_equal(port.tvl1flow(I0, I1), g["flow"])                 # bit-exact, end to end
    a, b = port.normalize(I0, I1)
    assert np.array_equal(a, g["norm0"]) and np.array_equal(b, g["norm1"])
    gs = port.gaussian(a, 0.8)
    assert np.array_equal(gs, g["gauss08"])
    z = port.zoom_out(gs)
    assert np.array_equal(z, g["zoom_out"])
    assert np.array_equal(port.zoom_in(z, w, h), g["zoom_in"])
    dx, dy = port.centered_gradient(gs)
    assert np.array_equal(dx, g["cgx"]) and np.array_equal(dy, g["cgy"])
    assert np.array_equal(port.bicubic_warp(gs, g["flow"][0], g["flow"][1], True), g["bicubic_warp"])


def test_port_matches_compiled_reference_stages(port, reflib):
    rng = np.random.RandomState(3)
    for ny, nx in [(23, 45), (12, 20), (37, 64), (31, 33)]:
        a = (rng.rand(ny, nx) * 200).astype(np.float32)
        b = (rng.rand(ny, nx) * 180 + 3).astype(np.float32)
        u, v = (b - 90) / 30, (a - 100) / 40
        for x, y in [(port.normalize(a, b), reflib.normalize(a, b)),
                     (port.centered_gradient(a), reflib.centered_gradient(a)),
                     (port.forward_gradient(a), reflib.forward_gradient(a))]:
            assert all(np.array_equal(p, q) for p, q in zip(x, y))
        assert np.array_equal(port.gaussian(a, 0.8), reflib.gaussian(a, 0.8))
        assert np.array_equal(port.gaussian(a, 1.0392), reflib.gaussian(a, 1.0392))
        assert np.array_equal(port.zoom_out(a), reflib.zoom_out(a))
        assert np.array_equal(port.zoom_in(a, 2 * nx - 1, 2 * ny), reflib.zoom_in(a, 2 * nx - 1, 2 * ny))
        assert np.array_equal(port.divergence(a, b), reflib.divergence(a, b))
        assert np.array_equal(port.bicubic_warp(a, u, v, True), reflib.bicubic_warp(a, u, v, True))
        assert np.array_equal(port.bicubic_warp(a, u, v, False), reflib.bicubic_warp(a, u, v, False))


@pytest.mark.parametrize("h,w,iso", [(90, 160, "iso3200"), (75, 101, "iso12800"), (120, 200, "clean")])
def test_port_matches_compiled_reference_end_to_end(port, reflib, h, w, iso):
    I0, I1 = synth.gray_pair(h, w, iso)
    ref = reflib.tvl1flow(I0, I1)
    assert np.array_equal(port.tvl1flow(I0, I1), ref)
    # the double-sum stopping rule (what the CUDA path uses) takes the same decisions on these inputs
    wide, it_w, _, _, _ = port.tvl1flow_traced(I0, I1, err_mode=1)
    _, it_f, _, _, _ = port.tvl1flow_traced(I0, I1, err_mode=0)
    assert np.array_equal(it_w, it_f) and np.array_equal(wide, ref)


def test_flow_sign_convention(port):
    """I1(x + u) ~ I0(x): the recovered flow follows the synthetic motion."""
    I0, I1 = synth.gray_pair(120, 200, "clean")
    u = port.tvl1flow(I0, I1)
    assert -4.5 < u[0, 20:-20, 20:-20].mean() < -1.5 and 0.5 < u[1, 20:-20, 20:-20].mean() < 3.0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "warp_*.npz"))))
def test_warp_ref_matches_reference_golden(path):
    g = np.load(path)
    x, flow = torch.from_numpy(g["x"]), torch.from_numpy(g["flow"])
    yb, m = warp_ref.warp(x, flow, "bicubic")
    yl, _ = warp_ref.warp(x, flow, "bilinear")
    assert np.array_equal(yb.numpy(), g["bicubic"]) and np.array_equal(yl.numpy(), g["bilinear"])
    assert np.array_equal(m.numpy(), g["mask"])
    up = warp_ref.upsample_factor_2(torch.from_numpy(g["half_flow"]), 2)
    assert np.array_equal(up.numpy(), g["up2x2"])
    # independent numpy restatement of the ATen semantics (A = -0.75, centre unclipped, taps clamped)
    man = warp_ref.grid_sample_manual(g["x"][0].astype(np.float64), g["flow"][0])
    assert np.abs(man - g["bicubic"][0]).max() < 2e-4 * max(1.0, np.abs(g["bicubic"]).max())


@pytest.mark.parametrize("name,iso,feat_future", [("pipeline_convunet_iso3200", "iso3200", False),
                                                  ("pipeline_convunet_feat_future_iso12800", "iso12800", True)])
def test_oracle_reproduces_reference_pipeline_psnr(port, name, iso, feat_future):
    """The end-to-end fixtures (tests/golden/make_pipeline_golden.py: reference flows, reference demosaic and warp,
    shipped recurrent-convunet checkpoints, with and without feature recurrence + future frame) replayed on the CPU
    with the ORACLE's flow, warp and demosaic in the loop: same flows bit for bit, same PSNR per frame.  The GPU tests
    (test_gpu_pipeline.py) run the same loops with the CUDA path."""
    from oracle import demosaic_ref
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    net = torch.jit.load(os.path.join(GOLDEN, name + "_denoiser.pt"), map_location="cpu").eval()
    nfr, h, w = (int(v) for v in d["geometry"])
    seq = synth.sequence(nfr, h, w, iso)
    assert float(seq.numpy().astype(np.float64).sum()) == float(d["frames_checksum"])
    g = np.mean(seq.numpy(), axis=3)
    gt = torch.from_numpy(d["gt"].astype(np.float32))[:, None].repeat(1, 3, 1, 1)

    def up_flow(tgt, src, want):
        flow = port.tvl1flow(g[tgt], g[src])
        assert np.array_equal(flow.transpose(1, 2, 0), want)
        return warp_ref.upsample_factor_2(torch.from_numpy(flow[None]), 2)

    with torch.no_grad():
        n = [torch.from_numpy(demosaic_ref.hamilton_adam((2.0 * (seq[t] / 4095.0) - 1.0).permute(2, 0, 1)[None].numpy()))
             for t in range(nfr)]
        lastden, psnrs = n[0], []
        lastfeat = torch.zeros(1, 48, 2 * h, 2 * w)
        for t in range(1, nfr - 1 if feat_future else nfr):
            up = up_flow(t, t - 1, d["flows"][t - 1])
            netinput = torch.cat((warp_ref.warp(lastden, up, "bicubic")[0], n[t]), 1)
            if feat_future:
                fup = up_flow(t, t + 1, d["future_flows"][t - 1])
                netinput = torch.cat((netinput, warp_ref.warp(n[t + 1], fup, "bicubic")[0]), 1)
                den, lastfeat = net(netinput, warp_ref.warp(lastfeat, up, "bicubic")[0])
            else:
                den = net(netinput)
            lastden = den.clone()
            psnrs.append(float(10 * torch.log10(4.0 / torch.mean((den - gt[t:t + 1]) ** 2))))
    assert np.max(np.abs(np.array(psnrs) - d["psnr"])) <= 1e-3
    assert float((den[0] - torch.from_numpy(d["denoised_last"])).abs().max()) <= 1e-5


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "demosaic_*.npz"))))
def test_demosaic_ref_matches_reference_golden(path):
    """oracle/demosaic_ref.py against vectors produced by the reference's own HamiltonAdam module (all 4 patterns)."""
    from oracle import demosaic_ref
    g = np.load(path)
    y = demosaic_ref.hamilton_adam(g["x"], str(g["pattern"]))
    assert y.shape == g["y"].shape and np.abs(y - g["y"]).max() <= 1e-6
    assert np.array_equal(demosaic_ref.remosaick(g["y"][:, :3], "gbrg"), g["remosaick"])


PARAM_SETS = [dict(zfactor=0.7, nwarps=3, tau=0.2, lam=0.1, theta=0.25, epsilon=0.02),
              dict(zfactor=0.5, nwarps=2, nscales=3, lam=0.3, theta=0.4, epsilon=0.005),
              dict(zfactor=0.35, nwarps=4, fscale=1),
              dict(zfactor=0.8, nscales=4, nwarps=1)]


@pytest.mark.parametrize("kw", PARAM_SETS)
def test_port_matches_compiled_reference_with_other_parameters(port, reflib, kw):
    """tau, lambda, theta, nscales, fscale, zfactor, nwarps, epsilon away from the bridge's defaults (the north star's
    "same parameters" clause): the port still equals the reference's Dual_TVL1_optic_flow_multiscale bit for bit,
    including the generic zoom_out path (zoom factors other than 0.5: fractional bicubic resampling)."""
    I0, I1 = synth.gray_pair(72, 110, "iso3200")
    assert np.array_equal(port.multiscale(I0, I1, **kw), reflib.multiscale(I0, I1, **kw))
