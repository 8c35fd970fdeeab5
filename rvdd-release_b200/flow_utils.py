"""Mirror of the reference's ``util/flow_utils.py`` alignment API (flow_utils.py:70-174) on the CUDA bridge.

Same function names and argument meaning: ``warp``, ``single_warp``, ``compute_flow``, ``compute_flow_and_warp``,
``upsample_factor_2``.  Differences, all deliberate:

* tensors must live on the GPU (numpy entry points move them there); there is no CPU fallback;
* ``warp`` returns the mask on ``x.device`` -- the reference forces it to a CPU FloatTensor
  (flow_utils.py:102), a hidden device->host sync per call that every caller then throws away;
* ``warp`` takes two optional keywords, ``flow_mul`` and a half-resolution ``flow``, that fuse
  ``upsample_factor_2(flow, multiply_by=2)`` (recurrent_model.py:129) into the gather;
* inference only: the warp is not differentiable.
"""
import numpy as np
import torch

from . import bridge as _bridge


def torch_flow(flow):
    """(h, w, 2) numpy flow -> [1, 2, h, w] tensor (flow_utils.py:10-14)."""
    return torch.from_numpy(np.ascontiguousarray(flow, dtype=np.float32)).unsqueeze(0).permute(0, 3, 1, 2)


def torch_image(iio_img_like):
    """(h, w, c) numpy image -> [1, c, h, w] tensor (flow_utils.py:16-17)."""
    return torch.from_numpy(np.ascontiguousarray(iio_img_like, dtype=np.float32)).permute(2, 0, 1).unsqueeze(0)


def warp(x, flow, interp, flow_mul=1.0, want_mask=True):
    """
    Backward-warp a tensor according to the given optical flow (flow_utils.py:70-102).

    Args:
        x    : CUDA tensor [B, C, H, W], image / feature map to be warped (any strides, e.g. a channel slice).
        flow : CUDA tensor [B, 2, H, W] (or [B, 2, H/2, W/2] to fuse the x2 bilinear upsampling), optical flow
        interp: 'bilinear' or 'bicubic'

    Returns:
        y   : [B, C, H, W], x warped according to flow
        mask: [B, 1, H, W] float mask of defined pixels (on x.device), or None with want_mask=False
    """
    if interp not in ("bilinear", "bicubic"):
        raise ValueError("warp: unsupported interpolation %r" % (interp,))
    return _bridge.default_bridge().warp(x, flow.to(x.device), interp, flow_mul=flow_mul, want_mask=want_mask)


def single_warp(iio_img_like, np_flow, interpolation="bicubic", givemask=False):
    """Warp one (H, W, C) numpy image by an (H, W, 2) numpy flow and return it in the same layout
    (flow_utils.py:105-122)."""
    dev = _bridge.default_bridge().device
    img = torch_image(iio_img_like).to(dev)
    flow = torch_flow(np_flow).to(dev)
    warped, mask = warp(img, flow, interpolation)
    out = warped.cpu().numpy().squeeze(0).transpose(1, 2, 0)
    if givemask:
        return out, mask.cpu()
    return out


def compute_flow(iio_img1, iio_img2, flow_type='tvl1'):
    """Flow from img2 (target) to img1 (source) (flow_utils.py:126-134)."""
    from .library import CPPbridge
    if flow_type != 'tvl1':
        raise TypeError(f"Unknown flow type {flow_type}")
    return CPPbridge().TVL1_flow(iio_img2, iio_img1)


def compute_flow_and_warp(iio_img1, iio_img2, flow_type='tvl1', interpolation='bicubic', iio_flow_img1=None):
    """flow = TVL1(target=img2, source=flow_img1); warped = single_warp(img1, flow) (flow_utils.py:138-156).
    Returns (warped, undef_mask, flow)."""
    if iio_flow_img1 is None:
        iio_flow_img1 = iio_img1
    from .library import CPPbridge
    if flow_type != 'tvl1':
        raise TypeError(f"Unknown flow type {flow_type}")
    flow = CPPbridge().TVL1_flow(iio_img2, iio_flow_img1)
    warped, undef_mask = single_warp(iio_img1, flow, interpolation, givemask=True)
    return warped, undef_mask, flow


def upsample_factor_2(downsampled_batch, multiply_by=1.):
    """[..., C, H, W] -> [..., C, 2H, 2W], bilinear, align_corners=True, times multiply_by (flow_utils.py:159-174)."""
    return _bridge.default_bridge().upsample2(downsampled_batch, multiply_by)


def remosaick(x):
    """RGB [B, 3, 2H, 2W] -> packed 'gbrg' raw [B, 4, H, W] (util/Hamilton_Adam_demo.py:237-246)."""
    return torch.stack((x[:, 1, 0::2, 0::2], x[:, 2, 0::2, 1::2], x[:, 0, 1::2, 0::2], x[:, 1, 1::2, 1::2]), dim=1)


def compute_flows_from_denoised(denoised, noisy_packed, predemosaic=True):
    """Online flow from the previous DENOISED frame to the current noisy frame, entirely on the GPU -- the
    ``--val_flow_from_denoised`` path of the reference (validate.py:16-38), which moves the denoised frame to the
    CPU, remosaicks it, calls the C TV-L1 and moves the flow back, every frame.

    denoised     : CUDA tensor [1, 3, 2H, 2W] (or [1, 4, H, W] with predemosaic=False) in the network's [-1, 1] range
    noisy_packed : CUDA tensor [1, 4, H, W], the current noisy packed-raw frame (data['n'][0, -4:]) in [-1, 1]
    returns      : [1, 1, 2, H, W] flow (source = denoised frame, target = noisy frame), the layout of data['flow']
    """
    b = _bridge.default_bridge()
    if predemosaic:
        # remosaick + singleiT of library.py:67 ((x + 1) / 2) + mean of the 4 packed channels (library.py:165-167), fused
        g_src = b.remosaick_gray(denoised[:1], "gbrg", add=1.0, mul=0.5)
        g_tgt = b.gray(((noisy_packed[:1] + 1.) / 2.).permute(0, 2, 3, 1).contiguous())
        gray = torch.cat((g_tgt, g_src), 0)
    else:
        pair = torch.cat((noisy_packed[:1], denoised[:1]), 0)
        gray = b.gray(((pair + 1.) / 2.).permute(0, 2, 3, 1).contiguous())
    flow = b.tvl1_flow(gray, src=[1], tgt=[0])                   # target = noisy frame, source = denoised frame
    return flow.unsqueeze(0)
