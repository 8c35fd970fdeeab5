"""Generate the end-to-end pipeline fixtures by RUNNING THE REFERENCE ITSELF (build container only: needs
/root/reference and oracle/_ref/libref_omp.so; the committed files are what travels to the GPU box).

The north star asks that the PSNR of the shipped trained-nets checkpoints stays within 0.02 dB of the reference
pipeline.  This script runs that pipeline on a small synthetic sequence with the reference's own code on the CPU:

    packed noisy raw --(library.CPPbridge.TVL1_flow -> compiled reference C)--> flows t-1 -> t
    HamiltonAdam('gbrg') demosaic, upsample_factor_2(flow, 2), warp(lastden, flow, 'bicubic')    (util/flow_utils.py)
    netDenoise = convunet-mode=fixedfeatures + trained-nets/recurrent-convunet-iso3200            (networks/unet.py)
    recurrence exactly as models/recurrent_model.py:233-345 with D = 1, fD = 0 (scripts/test-recurrent-convunet.sh)

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


def reference_pipeline():
    _shim_missing_modules()
    sys.path.insert(0, REF)
    import library                                    # the reference's own modules
    from networks.unet import get_UNet_cls
    from util.flow_utils import upsample_factor_2, warp
    from util.Hamilton_Adam_demo import HamiltonAdam
    from util.util import psnr

    lib = os.path.join(ROOT, "oracle", "_ref", "libref_omp.so")
    assert os.path.exists(lib), "build the compiled reference first: make -C oracle"
    bridge = library.CPPbridge(lib)

    seq = synth.sequence(NFR, H, W, ISO_NAME).numpy()                                   # (NFR, H, W, 4) raw values
    clean = torch.stack([synth.clean_frame(t, H, W) for t in range(NFR)], 0)
    f0 = clean[0]
    cfg = synth.ISO[ISO_NAME]
    clean = (cfg["lo"] + (clean - f0.min()) / (f0.max() - f0.min()) * (cfg["hi"] - cfg["lo"])).clamp(0, 4095).float()

    T, _ = library.define_transforms()
    ha = HamiltonAdam("gbrg")
    net = get_UNet_cls("fixedfeatures")(in_channels=6, out_channels=3, depth=4)
    state = torch.load(os.path.join(REF, "trained-nets", "recurrent-convunet-iso3200_net_Denoise.pth"), map_location="cpu")
    net.load_state_dict(state)
    net.eval()

    with torch.no_grad():
        # data/infer4rec_dataset.py:195-218: frames / (2^12 - 1), T = 2x - 1; models/recurrent_model.py:126 demosaic
        n = [ha(T(seq[t] / np.float32(4095.0))[None]) for t in range(NFR)]
        # ground truth: the clean full-resolution image (gray texture, replicated over R, G, B)
        gt = [(2.0 * ha.pack_in_one(clean[t].permute(2, 0, 1)[None]) / 4095.0 - 1.0)[:, None].repeat(1, 3, 1, 1)
              for t in range(NFR)]
        flows, dens, psnrs = [], [], []
        lastden = n[0]                                                                   # recurrent_model.py:236-238
        for t in range(1, NFR):
            # data/base_dataset.py:174-178 -> util/flow_utils.py:144-149: flow = TVL1_flow(target, source)
            flow = bridge.TVL1_flow(seq[t], seq[t - 1])                                  # (H, W, 2)
            flows.append(np.ascontiguousarray(flow))
            fl = torch.from_numpy(flow.transpose(2, 0, 1).copy())[None, None, None]      # [B, 1, D, 2, h, w]
            fl = upsample_factor_2(fl, multiply_by=2)                                    # recurrent_model.py:129
            warped, _ = warp(lastden, fl[:, 0, 0], interp="bicubic")                     # :281-288
            den = net(torch.cat((warped, n[t]), 1))                                      # :299-327
            lastden = den.clone()                                                        # :335-337
            dens.append(den[0].numpy())
            psnrs.append(float(psnr(den, gt[t], 2.0)))
        ex_net = torch.cat((lastden, n[-1]), 1)
        # networks/unet.py:163 builds its padding buffer with torch.zeros(size).to(x.device), which a trace freezes to
        # the CPU.  At this geometry (96 x 160, divisible by 2^depth) the padding is the identity, so FOR THE EXPORTED
        # TRACE ONLY the helper is replaced by a size-checked pass-through; the golden numbers above come from the
        # unmodified network, and the trace is asserted equal to it below.
        import networks.unet as _unet

        def _same_size_pad(size, x):
            assert tuple(size) == tuple(x.size()), "trace export assumes no padding"
            return x
        _orig_pad, _unet.zero_pad_features = _unet.zero_pad_features, _same_size_pad
        traced_net = torch.jit.trace(net, ex_net)
        _unet.zero_pad_features = _orig_pad
        ex_ha = T(seq[0] / np.float32(4095.0))[None]
        traced_ha = torch.jit.trace(ha, ex_ha)
        assert torch.equal(traced_net(ex_net), net(ex_net)) and torch.equal(traced_ha(ex_ha), ha(ex_ha))
    return seq, np.stack([g[0].numpy() for g in gt]), np.stack(flows), np.stack(dens), np.array(psnrs), traced_net, traced_ha


if __name__ == "__main__":
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                                 # nothing is written next to the reference
        seq, gt, flows, dens, psnrs, tnet, tha = reference_pipeline()
        os.chdir(cwd)
    torch.jit.save(tnet, os.path.join(HERE, "pipeline_convunet_iso3200_denoiser.pt"))
    torch.jit.save(tha, os.path.join(HERE, "pipeline_hamilton_adams_gbrg_48x80.pt"))
    np.savez_compressed(os.path.join(HERE, "pipeline_convunet_iso3200.npz"), geometry=np.array([NFR, H, W]),
                        frames_checksum=np.float64(seq.astype(np.float64).sum()), gt=gt[:, 0].astype(np.float16),
                        flows=flows, denoised_last=dens[-1], denoised_mean=dens.mean(axis=(1, 2, 3)), psnr=psnrs)
    print("reference PSNR per frame:", np.round(psnrs, 3))
