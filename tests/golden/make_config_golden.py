"""PSNR fixtures AT THE CONFIGURED GEOMETRIES of BASELINE.json, made by RUNNING THE REFERENCE PIPELINE ITSELF on the CPU
(build container only: /root/reference + oracle/_ref/libref_omp.so):

    config 2  recurrent-convunet-iso3200                    30 frames, 1280x720 packed raw -> network at 2560x1440
    config 3  recurrent-convunet+feat-future-iso12800       30 frames, two flows per frame + 48-channel feature warp
    config 5  recurrent-ConvNeXtUnet+feat-future-iso3200    5 frames of 1920x1080 packed raw -> network at 3840x2160
    (+ a 6-frame 80x48 ConvNeXt case whose TorchScript trace is the exported denoiser; the trace is shape-generic)

Same loop as tests/golden/make_pipeline_golden.py (flows by the compiled reference C through library.CPPbridge,
HamiltonAdam, upsample_factor_2, warp, the shipped checkpoint, recurrence of models/recurrent_model.py:233-345), but the
frames come from synth.exact_sequence -- bit-reproducible on any machine AND on the GPU -- and the fixture holds no image
tensors: per-frame PSNR, per-flow SHA-256 + a 16x-subsampled copy, per-frame mean of the denoised output and an
8x-subsampled copy of the last denoised frame.  tests/test_gpu_pipeline_configs.py replays the loop on the GPU with OUR
flow / warp / demosaic and compares.

    python tests/golden/make_config_golden.py [c2] [c3] [c5] [cn_small]          (about 40 minutes of CPU for all)
"""
import hashlib
import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
import make_pipeline_golden as mpg  # noqa: E402
from rvdd_release_b200 import synth  # noqa: E402

CONFIGS = {
    #        checkpoint                                      network spec                   in  iso         feat+future  frames H     W
    "c2": ("recurrent-convunet-iso3200", "convunet-mode=fixedfeatures", 6, "iso3200", False, 30, 720, 1280),
    "c3": ("recurrent-convunet+feat-future-iso12800", "convunet-mode=fixedfeatures+feat", 9, "iso12800", True, 30, 720, 1280),
    "c5": ("recurrent-ConvNeXtUnet+feat-future-iso3200", "newunet-mode=feat", 9, "iso3200", True, 5, 1080, 1920),
    "cn_small": ("recurrent-ConvNeXtUnet+feat-future-iso3200", "newunet-mode=feat", 9, "iso3200", True, 6, 48, 80),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(name):
    ckpt, spec, in_ch, iso, feat_future, NFR, H, W = CONFIGS[name]
    mpg._shim_missing_modules()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import library
    from networks import define_net_arch
    from util.flow_utils import upsample_factor_2, warp
    from util.Hamilton_Adam_demo import HamiltonAdam
    from util.util import psnr

    bridge = library.CPPbridge(os.path.join(ROOT, "oracle", "_ref", "libref_omp.so"))
    t0 = time.time()
    seq_t = synth.exact_sequence(NFR, H, W, iso)
    seq = seq_t.numpy()
    cfg = synth.ISO[iso]
    T, _ = library.define_transforms()
    ha = HamiltonAdam("gbrg")
    net = define_net_arch(in_ch, 3, spec, init_gain=None)             # networks/__init__.py:120 (CPU: no DataParallel wrap)
    net.load_state_dict(torch.load(os.path.join(REF, "trained-nets", ckpt + "_net_Denoise.pth"), map_location="cpu"))
    net.eval()

    def flow_of(tgt, src):
        flow = np.ascontiguousarray(bridge.TVL1_flow(seq[tgt], seq[src]))                # (H, W, 2)
        fl = torch.from_numpy(flow.transpose(2, 0, 1).copy())[None, None, None]
        return flow, upsample_factor_2(fl, multiply_by=2)[:, 0, 0]                       # recurrent_model.py:129

    def gt_of(t):
        clean = (cfg["lo"] + synth.exact_clean_frame(t, H, W) * (cfg["hi"] - cfg["lo"])).float()
        return (2.0 * ha.pack_in_one(clean.permute(2, 0, 1)[None]) / 4095.0 - 1.0)[:, None].repeat(1, 3, 1, 1)

    flows, fflows, psnrs, den_means = [], [], [], []
    with torch.no_grad():
        n = [ha(T(seq[t] / np.float32(4095.0))[None]) for t in range(NFR)]
        lastden = n[0]
        lastfeat = net.get_rec_nil_features(1, 2 * H, 2 * W, device="cpu", non_blocking=False) if feat_future else None
        last = NFR - 1 if feat_future else NFR
        den = ex = None
        for t in range(1, last):
            flow, up = flow_of(t, t - 1)
            flows.append(flow)
            warped, _ = warp(lastden, up, interp="bicubic")
            netinput = torch.cat((warped, n[t]), 1)
            if feat_future:
                featinput = [f for f in lastfeat]
                featinput[0][:, 0:48] = warp(featinput[0][:, 0:48].clone(), up, interp="bicubic")[0]
                net.set_rec_features(featinput)
                fflow, fup = flow_of(t, t + 1)
                fflows.append(fflow)
                netinput = torch.cat((netinput, warp(n[t + 1], fup, interp="bicubic")[0]), 1)
                ex = (netinput, featinput[0].clone())
            den = net(netinput)
            lastden = den.clone()
            if feat_future:
                lastfeat = [net.get_current_features()[0]]
            psnrs.append(float(psnr(den, gt_of(t), 2.0)))
            den_means.append(float(den.double().mean()))
            print("  %s frame %d/%d  psnr %.4f  (%.0f s)" % (name, t, last - 1, psnrs[-1], time.time() - t0), flush=True)
        traced = None
        if name == "cn_small":
            import networks.new_unet as _nu

            def _same_size_pad(size, x):
                assert tuple(size) == tuple(x.size()), "trace export assumes no padding"
                return x
            orig, _nu.zero_pad_features = _nu.zero_pad_features, _same_size_pad
            try:
                traced = torch.jit.trace(mpg._FeatWrap(net), ex)
            finally:
                _nu.zero_pad_features = orig
            assert torch.equal(traced(*ex)[0], den)
            x2, f2 = torch.randn(1, 9, 64, 96), torch.randn(1, 48, 64, 96)               # the trace must be shape-generic
            net.set_rec_features([f2.clone()])
            assert torch.equal(traced(x2, f2.clone())[0], net(x2))
    out = dict(geometry=np.array([NFR, H, W]), iso=iso, checkpoint=ckpt, sha_frames=sha(seq), psnr=np.array(psnrs),
               flow_sha=np.array([sha(f) for f in flows]), flow_sub=np.stack([f[::16, ::16] for f in flows]),
               denoised_mean=np.array(den_means), denoised_last_sub=den[0, :, ::8, ::8].numpy())
    if feat_future:
        out["future_flow_sha"] = np.array([sha(f) for f in fflows])
        out["future_flow_sub"] = np.stack([f[::16, ::16] for f in fflows])
    return out, traced


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in CONFIGS] or list(CONFIGS)
    torch.set_num_threads(os.cpu_count() or 1)
    cwd = os.getcwd()
    for name in names:
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            out, traced = run(name)
            os.chdir(cwd)
        np.savez_compressed(os.path.join(HERE, "config_%s.npz" % name), **out)
        if traced is not None:
            torch.jit.save(traced, os.path.join(HERE, "pipeline_convnext_feat_future_iso3200_denoiser.pt"))
        print(name, "reference PSNR per frame:", np.round(out["psnr"], 3).tolist(), flush=True)
