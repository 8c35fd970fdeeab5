"""CPU restatement of the reference's torch warp path.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

``warp`` and ``upsample_factor_2`` restate util/flow_utils.py:70-102 and :159-174 of the reference with the same
torch primitives (``grid_sample`` bicubic / border / align_corners=True, ``interpolate`` bilinear x2), on the CPU.
``grid_sample_manual`` is an independent numpy restatement of what ATen's grid_sampler_2d computes for that mode
(Keys A=-0.75, centre unclipped, taps clamped), used to pin the semantics the CUDA kernel implements.

Pin: checked against the reference's own ``util/flow_utils.py`` imported from /root/reference when the golden
vectors under tests/golden/ were generated (tests/golden/make_golden.py); the arithmetic below it is torch's
(third-party; the reference pins torch==1.8.0, this image has 2.11) -- "parity unpinned" beyond "equals this torch".
"""
import numpy as np
import torch
import torch.nn.functional as F


def warp(x, flow, interp="bicubic"):
    """x [B,C,H,W], flow [B,2,H,W] (ch0 = x-displacement) -> (warped, mask [B,1,H,W] float)."""
    B, C, H, W = x.shape
    ys, xs = torch.meshgrid(torch.arange(H, device=x.device), torch.arange(W, device=x.device), indexing="ij")
    base = torch.stack((xs, ys), 0).unsqueeze(0).float()                 # flow_utils.py:84-89
    v = base + flow                                                       # :90
    gx = 2.0 * v[:, 0] / (W - 1) - 1.0                                    # :93
    gy = 2.0 * v[:, 1] / (H - 1) - 1.0                                    # :94
    mask = (gx >= -1) & (gx <= 1) & (gy >= -1) & (gy <= 1)                # :95-96
    grid = torch.stack((gx, gy), dim=-1)                                  # :97
    out = F.grid_sample(x, grid, padding_mode="border", mode=interp, align_corners=True)   # :98-99
    return out, mask.unsqueeze(1).float()


def upsample_factor_2(t, multiply_by=1.0):
    *rem, C, H, W = t.shape                                               # :168
    up = F.interpolate(t.reshape(-1, C, H, W), scale_factor=2, mode="bilinear", align_corners=True)   # :170-172
    return up.reshape(*rem, C, 2 * H, 2 * W) * multiply_by                # :174


def _cubic_coeffs(t, A=-0.75):
    def c1(x): return ((A + 2) * x - (A + 3)) * x * x + 1
    def c2(x): return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
    return [c2(t + 1), c1(t), c1(1 - t), c2(2 - t)]


def grid_sample_manual(x, flow):
    """numpy float64 restatement of warp(..., 'bicubic') for x [C,H,W], flow [2,H,W]."""
    C, H, W = x.shape
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    ix = xs + flow[0].astype(np.float64)
    iy = ys + flow[1].astype(np.float64)
    fx, fy = np.floor(ix), np.floor(iy)
    cx, cy = _cubic_coeffs(ix - fx), _cubic_coeffs(iy - fy)
    out = np.zeros((C, H, W))
    for r in range(4):
        yy = np.clip(fy - 1 + r, 0, H - 1).astype(int)
        row = np.zeros((C, H, W))
        for c in range(4):
            xx = np.clip(fx - 1 + c, 0, W - 1).astype(int)
            row += x[:, yy, xx] * cx[c]
        out += row * cy[r]
    return out
