"""Generate the end-to-end pipeline fixtures by RUNNING THE REFERENCE ITSELF (build container only: needs
/root/reference and oracle/_ref/libref_omp.so; the committed files are what travels to the GPU box).

The north star asks that the PSNR of the shipped trained-nets checkpoints stays within 0.02 dB of the reference
pipeline.  This script runs that pipeline on a small synthetic sequence with the reference's own code on the CPU:

    packed noisy raw --(library.CPPbridge.TVL1_flow -> compiled reference C)--> flows t-1 -> t
    HamiltonAdam('gbrg') demosaic, upsample_factor_2(flow, 2), warp(lastden, flow, 'bicubic')    (util/flow_utils.py)
    netDenoise = convunet-mode=fixedfeatures + trained-nets/recurrent-convunet-iso3200            (networks/unet.py)
    recurrence exactly as models/recurrent_model.py:233-345 with D = 1, fD = 0 (scripts/test-recurrent-convunet.sh)
and, second fixture, convunet-mode=fixedfeatures+feat + trained-nets/recurrent-convunet+feat-future-iso12800 with
--feature_rec --future_patch_depth 1 (scripts/test-recurrent-feat-future-convunet.sh): the 48-channel feature map of
the previous frame is warped as well, and the next noisy frame is warped by a second flow t+1 -> t

and stores per-frame PSNR (util/util.py:9-20, max_val 2.0) and the denoised frames.  The denoiser and the demosaic are
NOT part of this repository's scope, and their Python sources cannot travel, so they are exported as TorchScript
traces (weights + aten graph); tests/test_gpu_pipeline.py runs the same loop on the GPU with OUR flow and OUR warp in
place of the reference's and compares PSNR frame by frame.

    python tests/golden/make_pipeline_golden.py
"""
import ctypes
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
from rvdd_release_b200 import synth  # noqa: E402

H, W, NFR, ISO_NAME = 48, 80, 6, "iso3200"      # packed-raw geometry; the network runs at 96 x 160


def _shim_missing_modules():
    """library.py imports iio / skimage, networks/new_unet.py imports opt_einsum: none is installed, none is used on
    this path (arrays are passed in memory, inputs have 4 channels, the network is the conv UNet)."""
    iio = types.ModuleType("iio")
    iio.read = lambda p: (_ for _ in ()).throw(RuntimeError("iio shim: no file IO in this script"))
    iio.write = iio.read
    sk = types.ModuleType("skimage")
    sk.__path__ = []
    skc = types.ModuleType("skimage.color")
    skc.rgb2gray = lambda a: (a[..., 0] * 0.2125 + a[..., 1] * 0.7154 + a[..., 2] * 0.0721)
    skio = types.ModuleType("skimage.io")
    sk.color, sk.io = skc, skio
    oe = types.ModuleType("opt_einsum")
    oe.contract = lambda *a, **k: torch.einsum(*a)
    for name, mod in (("iio", iio), ("skimage", sk), ("skimage.color", skc), ("skimage.io", skio), ("opt_einsum", oe)):
        sys.modules.setdefault(name, mod)
    if not hasattr(np, "int"):
        np.int = int                                  # data/*.py still use the removed alias


class _FeatWrap(torch.nn.Module):
    """(netinput, warped previous features) -> (denoised, features of this frame): set_rec_features + forward +
    get_current_features of networks/unet.py:808-825 as one traceable call."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, netinput, feat):
        self.net.set_rec_features([feat])
        den = self.net(netinput)
        return den, self.net.get_current_features()[0]


def _trace(net, example, wrap=False):
    # networks/unet.py:163 builds its padding buffer with torch.zeros(size).to(x.device), which a trace freezes to the
    # CPU.  At this geometry (96 x 160, divisible by 2^depth) the padding is the identity, so FOR THE EXPORTED TRACE ONLY
    # the helper is replaced by a size-checked pass-through; the golden numbers come from the unmodified network, and
    # the trace is asserted equal to it by the caller.
    import networks.unet as _unet

    def _same_size_pad(size, x):
        assert tuple(size) == tuple(x.size()), "trace export assumes no padding"
        return x
    _orig_pad, _unet.zero_pad_features = _unet.zero_pad_features, _same_size_pad
    try:
        return torch.jit.trace(_FeatWrap(net) if wrap else net, example)
    finally:
        _unet.zero_pad_features = _orig_pad


def reference_pipeline(feat_future=False):
    """feat_future=False: recurrent-convunet-iso3200 (scripts/test-recurrent-convunet.sh, config 2).
    feat_future=True : recurrent-convunet+feat-future-iso12800 (scripts/test-recurrent-feat-future-convunet.sh,
    config 3): feature recurrence (48-channel warp) and one future frame (a second flow per frame)."""
    _shim_missing_modules()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import library                                    # the reference's own modules
    from networks.unet import get_UNet_cls
    from util.flow_utils import upsample_factor_2, warp
    from util.Hamilton_Adam_demo import HamiltonAdam
    from util.util import psnr

    lib = os.path.join(ROOT, "oracle", "_ref", "libref_omp.so")
    assert os.path.exists(lib), "build the compiled reference first: make -C oracle"
    bridge = library.CPPbridge(lib)
    iso = "iso12800" if feat_future else "iso3200"

    seq = synth.sequence(NFR, H, W, iso).numpy()                                        # (NFR, H, W, 4) raw values
    clean = torch.stack([synth.clean_frame(t, H, W) for t in range(NFR)], 0)
    f0 = clean[0]
    cfg = synth.ISO[iso]
    clean = (cfg["lo"] + (clean - f0.min()) / (f0.max() - f0.min()) * (cfg["hi"] - cfg["lo"])).clamp(0, 4095).float()

    T, _ = library.define_transforms()
    ha = HamiltonAdam("gbrg")
    if feat_future:
        net = get_UNet_cls("fixedfeatures+feat")(in_channels=9, out_channels=3, depth=4)
        ckpt = "recurrent-convunet+feat-future-iso12800_net_Denoise.pth"
    else:
        net = get_UNet_cls("fixedfeatures")(in_channels=6, out_channels=3, depth=4)
        ckpt = "recurrent-convunet-iso3200_net_Denoise.pth"
    net.load_state_dict(torch.load(os.path.join(REF, "trained-nets", ckpt), map_location="cpu"))
    net.eval()

    def flow_of(tgt, src):
        # data/base_dataset.py:174-178 / :232-236 -> util/flow_utils.py:144-149: flow = TVL1_flow(target, source)
        flow = np.ascontiguousarray(bridge.TVL1_flow(seq[tgt], seq[src]))                # (H, W, 2)
        fl = torch.from_numpy(flow.transpose(2, 0, 1).copy())[None, None, None]          # [B, 1, D, 2, h, w]
        return flow, upsample_factor_2(fl, multiply_by=2)[:, 0, 0]                       # recurrent_model.py:129

    with torch.no_grad():
        # data/infer4rec_dataset.py:195-218: frames / (2^12 - 1), T = 2x - 1; models/recurrent_model.py:126 demosaic
        n = [ha(T(seq[t] / np.float32(4095.0))[None]) for t in range(NFR)]
        # ground truth: the clean full-resolution image (gray texture, replicated over R, G, B)
        gt = [(2.0 * ha.pack_in_one(clean[t].permute(2, 0, 1)[None]) / 4095.0 - 1.0)[:, None].repeat(1, 3, 1, 1)
              for t in range(NFR)]
        flows, fflows, dens, psnrs = [], [], [], []
        lastden = n[0]                                                                   # recurrent_model.py:236-238
        lastfeat = net.get_rec_nil_features(1, 2 * H, 2 * W, device="cpu", non_blocking=False) if feat_future else None  # :240-245
        last = NFR - 1 if feat_future else NFR                                           # the last frame has no future
        for t in range(1, last):
            flow, up = flow_of(t, t - 1)
            flows.append(flow)
            warped, _ = warp(lastden, up, interp="bicubic")                              # :281-288
            netinput = torch.cat((warped, n[t]), 1)                                      # :299-311
            if feat_future:
                featinput = [f for f in lastfeat]                                        # :276-279
                featinput[0][:, 0:48] = warp(featinput[0][:, 0:48].clone(), up, interp="bicubic")[0]    # :290-297
                net.set_rec_features(featinput)                                          # :307-308
                fflow, fup = flow_of(t, t + 1)                                           # future frame t+1 -> t
                fflows.append(fflow)
                netinput = torch.cat((netinput, warp(n[t + 1], fup, interp="bicubic")[0]), 1)            # :314-324
                ex_feat = featinput[0].clone()
            den = net(netinput)                                                          # :327
            lastden = den.clone()                                                        # :335-337
            if feat_future:
                lastfeat = [net.get_current_features()[0]]                               # :339-345 (NoPF = 1)
            dens.append(den[0].numpy())
            psnrs.append(float(psnr(den, gt[t], 2.0)))
        if feat_future:
            traced_net = _trace(net, (netinput, ex_feat), wrap=True)
            assert torch.equal(traced_net(netinput, ex_feat)[0], den)
        else:
            traced_net = _trace(net, netinput)
            assert torch.equal(traced_net(netinput), den)
        ex_ha = T(seq[0] / np.float32(4095.0))[None]
        traced_ha = torch.jit.trace(ha, ex_ha)
        assert torch.equal(traced_ha(ex_ha), ha(ex_ha))
    out = dict(geometry=np.array([NFR, H, W]), frames_checksum=np.float64(seq.astype(np.float64).sum()),
               gt=np.stack([g[0, 0].numpy() for g in gt]).astype(np.float16), flows=np.stack(flows),
               denoised_last=dens[-1], denoised_mean=np.stack(dens).mean(axis=(1, 2, 3)), psnr=np.array(psnrs))
    if feat_future:
        out["future_flows"] = np.stack(fflows)
    return out, traced_net, traced_ha


if __name__ == "__main__":
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                                 # nothing is written next to the reference
        a, anet, aha = reference_pipeline(False)
        b, bnet, _ = reference_pipeline(True)
        os.chdir(cwd)
    torch.jit.save(anet, os.path.join(HERE, "pipeline_convunet_iso3200_denoiser.pt"))
    torch.jit.save(bnet, os.path.join(HERE, "pipeline_convunet_feat_future_iso12800_denoiser.pt"))
    torch.jit.save(aha, os.path.join(HERE, "pipeline_hamilton_adams_gbrg_48x80.pt"))
    np.savez_compressed(os.path.join(HERE, "pipeline_convunet_iso3200.npz"), **a)
    np.savez_compressed(os.path.join(HERE, "pipeline_convunet_feat_future_iso12800.npz"), **b)
    print("reference PSNR per frame:", np.round(a["psnr"], 3), np.round(b["psnr"], 3))
