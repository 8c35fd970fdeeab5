"""Tiny driver for ncu captures of the stage kernels: python tools/prof_stage.py warp48|warp4|hwc4|gauss|demosaic"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import bridge
br = bridge.default_bridge()
what = sys.argv[1]
if what.startswith("warp") or what == "hwc4":
    B, C, H, W = (1, 48, 1440, 2560) if what == "warp48" else (29, 4, 720, 1280)
    x = torch.randn(B, C, H, W, device="cuda")
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
    flow = torch.stack((5.0 + 3.0 * torch.sin(yy / 97.0), -3.0 + 2.0 * torch.cos(xx / 131.0)), 0)[None].repeat(B, 1, 1, 1).contiguous()
    flow += 0.05 * torch.randn_like(flow)
    out = torch.empty_like(x)
    if what == "hwc4":      # frames in their on-disk layout, channel innermost: the warp of the headline step
        x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        out = torch.empty(B, H, W, C, device="cuda").permute(0, 3, 1, 2)
    for _ in range(3):
        br.warp(x, flow, "bicubic", want_mask=False, out=out)
elif what == "gauss":       # the pyramid kernels of a 29-pair batch (presmoothing + fused zoom-out levels)
    import numpy as np
    from rvdd_release_b200 import synth
    frames = synth.sequence(30, 720, 1280, "iso3200", device="cuda")
    gray = br.gray(frames)
    for _ in range(2):
        br.tvl1_flow(gray, np.arange(29), np.arange(1, 30))
else:
    x = torch.rand(8, 4, 720, 1280, device="cuda") * 2 - 1
    for _ in range(3):
        br.demosaic(x, "gbrg")
torch.cuda.synchronize()
