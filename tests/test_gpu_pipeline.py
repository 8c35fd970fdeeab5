"""End-to-end gate of the north star: the PSNR of a shipped trained-nets checkpoint, run through the recurrent
inference loop with OUR flow and OUR warp on the GPU, stays within 0.02 dB of the reference pipeline.

The reference side (flows by the compiled reference C, HamiltonAdam demosaic, util/flow_utils.warp, the
recurrent-convunet-iso3200 checkpoint, recurrence of models/recurrent_model.py:233-345) was run on the CPU by
tests/golden/make_pipeline_golden.py; its per-frame PSNR, flows and last denoised frame are the fixture.  The denoiser
and the demosaic are out of this repository's scope: they are loaded as the TorchScript traces that script exported.
"""
import os

import numpy as np
import pytest
import torch

from rvdd_release_b200 import flow_utils, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _psnr(a, b, max_val=2.0):                                       # util/util.py:9-20
    return float(10.0 * torch.log10(max_val * max_val / torch.mean((a - b) ** 2)))


@pytest.fixture(scope="module")
def fixture():
    d = np.load(os.path.join(GOLDEN, "pipeline_convunet_iso3200.npz"))
    net = torch.jit.load(os.path.join(GOLDEN, "pipeline_convunet_iso3200_denoiser.pt"), map_location="cuda").eval()
    # the demosaic trace froze its mask tensors on the CPU (Hamilton_Adam_demo.py:201-224 builds them with torch.zeros):
    # it runs there on these tiny frames and its output moves to the GPU
    ha_cpu = torch.jit.load(os.path.join(GOLDEN, "pipeline_hamilton_adams_gbrg_48x80.pt"), map_location="cpu").eval()
    return d, net, (lambda x: ha_cpu(x.cpu()).cuda())


@pytest.mark.parametrize("fused_upsample,own_demosaic", [(False, False), (True, False), (True, True)])
def test_recurrent_convunet_psnr_within_0p02_db(bridge, fixture, fused_upsample, own_demosaic):
    d, net, ha = fixture
    nfr, h, w = (int(v) for v in d["geometry"])
    seq = synth.sequence(nfr, h, w, "iso3200")
    assert float(seq.numpy().astype(np.float64).sum()) == float(d["frames_checksum"]), "synthetic input drifted"
    frames = seq.cuda()
    gt = torch.from_numpy(d["gt"].astype(np.float32)).cuda()[:, None].repeat(1, 3, 1, 1)       # [nfr, 3, 2h, 2w]

    # offline flows t-1 -> t for the whole sequence in one batch (data/base_dataset.py:134-191)
    gray = bridge.gray(frames)
    flows = bridge.tvl1_flow(gray, src=list(range(nfr - 1)), tgt=list(range(1, nfr)), check=True)
    want = torch.from_numpy(d["flows"]).permute(0, 3, 1, 2)
    assert torch.equal(flows.cpu(), want), "flows differ from the reference C"

    torch.backends.cudnn.allow_tf32 = False                         # fp32 convolutions, as on the reference's CPU run
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        if own_demosaic:             # csrc/demosaic.cu instead of the traced reference module
            from rvdd_release_b200.hamilton_adam import HamiltonAdam
            ha = HamiltonAdam("gbrg")
        n = [ha((2.0 * (frames[t] / 4095.0) - 1.0).permute(2, 0, 1)[None].contiguous()) for t in range(nfr)]
        lastden = n[0]
        psnrs, den = [], None
        for t in range(1, nfr):
            if fused_upsample:       # x2 upsampling of the half-resolution flow inside the gather (recurrent_model.py:129)
                warped, _ = flow_utils.warp(lastden, flows[t - 1:t], "bicubic", flow_mul=2.0)
            else:                    # the reference's two calls
                up = flow_utils.upsample_factor_2(flows[t - 1:t], multiply_by=2)
                warped, _ = flow_utils.warp(lastden, up, "bicubic")
            den = net(torch.cat((warped, n[t]), 1))
            lastden = den.clone()
            psnrs.append(_psnr(den, gt[t:t + 1]))
    ref = d["psnr"]
    assert np.max(np.abs(np.array(psnrs) - ref)) <= 0.02, (psnrs, ref.tolist())
    # much tighter than the PSNR gate: the last denoised frame itself (cuDNN vs CPU convolutions included)
    assert float((den[0].cpu() - torch.from_numpy(d["denoised_last"])).abs().max()) <= 2e-3
