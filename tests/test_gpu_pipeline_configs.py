"""GPU (-m gpu): the north star's PSNR gate AT THE CONFIGURED GEOMETRIES (BASELINE.json configs 2, 3 and 5): the recurrent
inference loop of models/recurrent_model.py:233-345 with OUR flow (one batched solver call for the whole sequence), OUR
demosaic and OUR warps (FrameAligner: half-resolution flows, x2 upsampling fused, outputs written straight into the
network input) around the shipped checkpoints, against fixtures the reference pipeline itself produced on the CPU
(tests/golden/make_config_golden.py):

    c2        recurrent-convunet-iso3200                  30 frames of 1280x720 packed raw, network at 2560x1440
    c3        recurrent-convunet+feat-future-iso12800     the same + future frame + 48-channel feature warp
    c5        recurrent-ConvNeXtUnet+feat-future-iso3200  5 frames of 1920x1080 packed raw, network at 3840x2160
    cn_small  the ConvNeXt checkpoint on 6 frames of 80x48

Gates: flows bit-identical to the reference C (SHA-256 per flow), per-frame PSNR within 0.02 dB."""
import hashlib
import os

import numpy as np
import pytest
import torch

from rvdd_release_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACES = {"c2": "pipeline_convunet_iso3200_denoiser.pt", "c3": "pipeline_convunet_feat_future_iso12800_denoiser.pt",
          "c5": "pipeline_convnext_feat_future_iso3200_denoiser.pt", "cn_small": "pipeline_convnext_feat_future_iso3200_denoiser.pt"}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _psnr(a, b, max_val=2.0):                                       # util/util.py:9-20
    return float(10.0 * torch.log10(max_val * max_val / torch.mean((a - b) ** 2)))


def test_exact_sequence_is_device_independent(bridge):
    a = synth.exact_sequence(3, 96, 160, "iso12800")
    b = synth.exact_sequence(3, 96, 160, "iso12800", device="cuda")
    assert torch.equal(a, b.cpu())


@pytest.mark.parametrize("name", ["cn_small", "c2", "c3", "c5"])
def test_checkpoint_psnr_at_configured_geometry(bridge, name):
    from rvdd_release_b200.hamilton_adam import HamiltonAdam
    from rvdd_release_b200.recurrent_align import FrameAligner
    d = np.load(os.path.join(GOLDEN, "config_%s.npz" % name))
    nfr, h, w = (int(v) for v in d["geometry"])
    iso = str(d["iso"])
    feat_future = "future_flow_sha" in d.files
    frames = synth.exact_sequence(nfr, h, w, iso, device="cuda")
    assert _sha(frames.cpu().numpy()) == str(d["sha_frames"]), "synthetic input is not bit-reproducible here"

    # ---- all flows of the sequence in ONE batched solver call (data/base_dataset.py:134-249 order: past, then future)
    T = nfr - 2 if feat_future else nfr - 1                       # frames 1 .. T are denoised
    src = list(range(0, T)) + (list(range(2, T + 2)) if feat_future else [])
    tgt = list(range(1, T + 1)) * (2 if feat_future else 1)
    flows = bridge.tvl1_flow(bridge.gray(frames), src=src, tgt=tgt, check=True)
    hw2 = flows.permute(0, 2, 3, 1).contiguous().cpu().numpy()    # the (h, w, 2) arrays TVL1_flow returns / the files hold
    for k in range(T):
        sub = hw2[k][::16, ::16]
        epe = float(np.sqrt(((sub - d["flow_sub"][k]) ** 2).sum(-1)).mean())
        assert epe <= 0.01, ("past flow", k, epe)
        assert _sha(hw2[k]) == str(d["flow_sha"][k]), ("past flow differs from the reference C", k, epe)
        if feat_future:
            assert _sha(hw2[T + k]) == str(d["future_flow_sha"][k]), ("future flow differs from the reference C", k)

    # ---- the recurrent loop around the shipped checkpoint
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = torch.jit.load(os.path.join(GOLDEN, TRACES[name]), map_location="cuda").eval()
    ha = HamiltonAdam("gbrg")
    cfg = synth.ISO[iso]
    al = FrameAligner(depth=1, future_depth=1 if feat_future else 0, feature_channels=48 if feat_future else 0)
    psnrs, means, den = [], [], None
    with torch.no_grad():
        n = ha((2.0 * synth._div(frames, 4095.0) - 1.0).permute(0, 3, 1, 2).contiguous())   # [nfr, 3, 2h, 2w], one launch
        al.reset(n[0:1])
        for t in range(1, T + 1):
            if feat_future:
                netinput, featinput = al.step(n[t:t + 1], flows[t - 1:t], [n[t + 1:t + 2]], [flows[T + t - 1:T + t]])
                den, feat = net(netinput, featinput)
                al.update(den.clone(), feat.clone())
            else:
                netinput, _ = al.step(n[t:t + 1], flows[t - 1:t])
                den = net(netinput)
                al.update(den.clone())
            clean = (cfg["lo"] + synth.exact_clean_frame(t, h, w, device="cuda") * (cfg["hi"] - cfg["lo"])).float()
            gt = (2.0 * ha.pack_in_one(clean.permute(2, 0, 1)[None]) / 4095.0 - 1.0)[:, None].repeat(1, 3, 1, 1)[0]
            psnrs.append(_psnr(den, gt))
            means.append(float(den.double().mean()))
    ref = d["psnr"]
    dev = float(np.max(np.abs(np.array(psnrs) - ref)))
    print("%s: reference PSNR %s\n    ours %s\n    max |diff| %.5f dB" % (name, np.round(ref, 4).tolist(), np.round(psnrs, 4).tolist(), dev))
    assert dev <= 0.02, (psnrs, ref.tolist())
    assert np.max(np.abs(np.array(means) - d["denoised_mean"])) <= 1e-3
    last = den[0, :, ::8, ::8].cpu().numpy()
    assert float(np.abs(last - d["denoised_last_sub"]).max()) <= 2e-2      # recurrence-amplified cuDNN-vs-CPU rounding; PSNR is the gate


def test_batched_multi_sequence_driver_matches_fixture(bridge):
    """Config 5's driver shape at a small size: S sequences advance in lock step as the batch dimension of flow -> warp ->
    denoiser (rvdd_release_b200.infer.run_sequences).  Sequence 0 is the cn_small fixture's, so its PSNRs must match the
    reference pipeline's; sequence 1 has another noise realisation and must come out the same whether it runs alone or in
    the batch."""
    from rvdd_release_b200 import infer
    from rvdd_release_b200.hamilton_adam import HamiltonAdam
    d = np.load(os.path.join(GOLDEN, "config_cn_small.npz"))
    nfr, h, w = (int(v) for v in d["geometry"])
    iso = str(d["iso"])
    net = torch.jit.load(os.path.join(GOLDEN, TRACES["cn_small"]), map_location="cuda").eval()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    frames = torch.stack([synth.exact_sequence(nfr, h, w, iso, device="cuda", noise_seed=s) for s in (0, 1, 2)])
    cfg, ha = synth.ISO[iso], HamiltonAdam("gbrg")

    def gt(t):
        clean = (cfg["lo"] + synth.exact_clean_frame(t, h, w, device="cuda") * (cfg["hi"] - cfg["lo"])).float()
        return (2.0 * ha.pack_in_one(clean.permute(2, 0, 1)[None]) / 4095.0 - 1.0)[:, None].repeat(1, 3, 1, 1)

    with torch.no_grad():
        outs, tm = infer.run_sequences(frames, net, future_depth=1, feature_channels=48, denoiser_batch=2)
        solo, _ = infer.run_sequences(frames[1:2], net, future_depth=1, feature_channels=48)
    assert len(outs) == nfr - 2 and tm["pairs"] == 3 * 2 * (nfr - 2) and tm["frames"] == 3 * (nfr - 2)
    ps = np.array([[float(infer.psnr(o[s:s + 1], gt(t + 1))) for t, o in enumerate(outs)] for s in range(3)])
    assert np.max(np.abs(ps[0] - d["psnr"])) <= 0.02, (ps[0].tolist(), d["psnr"].tolist())
    for t in range(nfr - 2):
        assert float((outs[t][1:2] - solo[t]).abs().max()) <= 1e-4, t          # batch composition does not matter
    past, fut = infer.compute_all_flows(frames, 1)
    assert past.shape == (3, nfr - 2, 2, h, w) and fut.shape == (3, nfr - 2, 1, 2, h, w)
    hw2 = past[0].permute(0, 2, 3, 1).contiguous().cpu().numpy()
    assert all(_sha(hw2[k]) == str(d["flow_sha"][k]) for k in range(nfr - 2))
