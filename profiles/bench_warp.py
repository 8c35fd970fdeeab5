"""Throughput of the backward-warp kernel (util/flow_utils.py::warp replacement) at the model's tensor shapes:
frames 3 x 2H x 2W, feature maps 48 x 2H x 2W (configs 2/3: H x W = 720 x 1280 packed raw -> 1440 x 2560 RGB), with the
flow given at full resolution or at half resolution (fused upsample_factor_2).  Prints GB/s of algorithmic traffic
(read C planes + flow, write C planes) against the measured HBM peak, and torch's grid_sample path for comparison."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import bridge  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
br = bridge.default_bridge()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def torch_warp(x, flow):
    B, C, H, W = x.shape
    yy, xx = torch.meshgrid(torch.arange(H, device=x.device), torch.arange(W, device=x.device), indexing="ij")
    grid = torch.stack((xx, yy), 0)[None].float() + flow
    gx = 2.0 * grid[:, 0] / (W - 1) - 1.0
    gy = 2.0 * grid[:, 1] / (H - 1) - 1.0
    return F.grid_sample(x, torch.stack((gx, gy), -1), padding_mode="border", mode="bicubic", align_corners=True)


rows = []
ONLY_DM = "--demosaic-only" in sys.argv
ONLY_HWC = "--hwc-only" in sys.argv                       # python profiles/bench_warp.py --hwc-only: the HWC section alone
for (B, C, H, W) in ([] if ONLY_HWC or ONLY_DM else [(1, 3, 1440, 2560), (1, 48, 1440, 2560), (4, 48, 720, 1280), (29, 4, 720, 1280)]):
    x = torch.randn(B, C, H, W, device="cuda")
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
    flow = torch.stack((5.0 + 3.0 * torch.sin(yy / 97.0), -3.0 + 2.0 * torch.cos(xx / 131.0)), 0)[None].repeat(B, 1, 1, 1).contiguous()
    flow += 0.05 * torch.randn_like(flow)                     # smooth motion + a little roughness, like a TV-L1 flow
    half = (0.5 * flow[:, :, ::2, ::2]).contiguous()
    rough = 3.0 * torch.randn(B, 2, H, W, device="cuda")      # white-noise flow: tiles fall back to direct gathers
    out = torch.empty_like(x)
    by = 4.0 * B * H * W * (2 * C + 2)
    t_full = timeit(lambda: br.warp(x, flow, "bicubic", want_mask=False, out=out))
    t_half = timeit(lambda: br.warp(x, half, "bicubic", flow_mul=2.0, want_mask=False, out=out))
    t_ref = timeit(lambda: torch_warp(x, flow), n=5)
    t_rough = timeit(lambda: br.warp(x, rough, "bicubic", want_mask=False, out=out))
    rows.append(dict(shape=[B, C, H, W], ms=t_full, gbs=by / t_full / 1e6, frac=by / t_full / 1e6 / peak,
                     ms_fused_up2=t_half, ms_white_noise_flow=t_rough, ms_torch_grid_sample=t_ref))
    print(json.dumps(rows[-1]))

# (H, W, 4) frames in their on-disk layout (channel innermost): the warp of the headline step (compute_flow_and_warp)
for (B, C, H, W) in ([] if ONLY_DM else [(29, 4, 720, 1280), (8, 4, 1080, 1920)]):
    x = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2)
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
    flow = torch.stack((5.0 + 3.0 * torch.sin(yy / 97.0), -3.0 + 2.0 * torch.cos(xx / 131.0)), 0)[None].repeat(B, 1, 1, 1).contiguous()
    flow += 0.05 * torch.randn_like(flow)
    rough = 3.0 * torch.randn(B, 2, H, W, device="cuda")
    out = torch.empty(B, H, W, C, device="cuda").permute(0, 3, 1, 2)
    by = 4.0 * B * H * W * (2 * C + 2)
    t_full = timeit(lambda: br.warp(x, flow, "bicubic", want_mask=False, out=out))
    t_rough = timeit(lambda: br.warp(x, rough, "bicubic", want_mask=False, out=out))
    t_bil = timeit(lambda: br.warp(x, flow, "bilinear", want_mask=False, out=out))
    smooth = torch.stack((5.0 + 3.0 * torch.sin(yy / 97.0), -3.0 + 2.0 * torch.cos(xx / 131.0)), 0)[None].repeat(B, 1, 1, 1).contiguous()
    t_smooth = timeit(lambda: br.warp(x, smooth, "bicubic", want_mask=False, out=out))
    rows.append(dict(layout="HWC", shape=[B, C, H, W], ms=t_full, gbs=by / t_full / 1e6, frac=by / t_full / 1e6 / peak,
                     ms_white_noise_flow=t_rough, ms_noise_free_flow=t_smooth, ms_bilinear=t_bil, tma=bool(os.environ.get("RVDD_WARP_TMA")),
                     one_px_per_thread=bool(os.environ.get("RVDD_WARP_HWC_1PX"))))
    print(json.dumps(rows[-1]))

# Hamilton-Adams demosaic (csrc/demosaic.cu): 4 B read + 12 B written per full-resolution pixel
for (B, H, W) in ([] if ONLY_HWC else [(1, 720, 1280), (8, 720, 1280), (1, 1080, 1920)]):
    x = torch.rand(B, 4, H, W, device="cuda") * 2 - 1
    by = 16.0 * B * 4 * H * W
    t = timeit(lambda: br.demosaic(x, "gbrg"))
    rgb = br.demosaic(x, "gbrg")
    t2 = timeit(lambda: br.remosaick_gray(rgb, "gbrg"))
    rows.append(dict(kernel="demosaic_ha", packed_shape=[B, 4, H, W], ms=t, gbs=by / t / 1e6, frac=by / t / 1e6 / peak,
                     general_path_only=bool(os.environ.get("RVDD_DEMOSAIC_GENERAL")), remosaick_gray_ms=t2, remosaick_gray_gbs=(4.0 * B * H * W * 5) / t2 / 1e6))
    print(json.dumps(rows[-1]))
