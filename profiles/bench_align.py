"""Per-frame cost of the alignment work around the denoiser at the model's full resolution (packed 720 x 1280 ->
1440 x 2560), configs 2 and 3 of BASELINE.json, WITHOUT the denoiser (out of scope): Hamilton-Adams demosaic of the new
frame(s) + the backward warps that build the network input (rvdd_release_b200.recurrent_align.FrameAligner), and the
same with the flow computed online from the previous denoised frame (validate.py:16-38 path).  For comparison: the
reference's formulation of the same warps in torch on the same GPU (meshgrid + grid_sample + F.interpolate per call).
Prints one JSON line per configuration."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import flow_utils, synth  # noqa: E402
from rvdd_release_b200.recurrent_align import FrameAligner  # noqa: E402

H, W = 720, 1280


def torch_warp(x, flow):
    """The reference's formulation (util/flow_utils.py:70-102) in torch on the GPU: the comparison arm."""
    B, C, Hh, Ww = x.shape
    ys, xs = torch.meshgrid(torch.arange(Hh, device=x.device), torch.arange(Ww, device=x.device), indexing="ij")
    grid = torch.stack((xs, ys), 0)[None].float() + flow
    gx = 2.0 * grid[:, 0] / (Ww - 1) - 1.0
    gy = 2.0 * grid[:, 1] / (Hh - 1) - 1.0
    return F.grid_sample(x, torch.stack((gx, gy), -1), padding_mode="border", mode="bicubic", align_corners=True)


def torch_up2(flow):
    return F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0


def timeit(fn, n=30):
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


seq = synth.sequence(3, H, W, "iso3200", device="cuda")
packed = (2.0 * (seq / 4095.0) - 1.0).permute(0, 3, 1, 2).contiguous()                 # [3, 4, H, W]
yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
flow = torch.stack((2.5 + 1.5 * torch.sin(yy / 97.0), -1.5 + torch.cos(xx / 131.0)), 0)[None].contiguous()
flow += 0.03 * torch.randn_like(flow)

for name, fD, Cf in (("config 2: recurrent-convunet (1 flow, 3-ch warp)", 0, 0),
                     ("config 3: recurrent-convunet+feat-future (2 flows, 3 + 48 + 3 channel warps)", 1, 48)):
    al = FrameAligner(depth=1, future_depth=fD, feature_channels=Cf)
    n = al.demosaic(packed)
    al.reset(n[0:1])
    den = n[0:1].clone()
    feat = torch.randn(1, 48, 2 * H, 2 * W, device="cuda") if Cf else None
    al.update(den, feat)

    def ours():
        nn = al.demosaic(packed[1:2 + fD])                       # the frames that are new at this step
        al.step(nn[0:1], flow, [nn[1:2]] if fD else [], [flow] if fD else [])

    def torch_ref():                                            # util/flow_utils.py on the GPU, as the reference runs it
        up = torch_up2(flow)
        outs = [torch_warp(den, up), n[1:2]]
        if fD:
            outs.append(torch_warp(n[2:3], up))
        torch.cat(outs, 1)
        if Cf:
            torch_warp(feat.clone(), up)

    t_ours, t_ref = timeit(ours), timeit(torch_ref, n=10)
    t_online = timeit(lambda: flow_utils.compute_flows_from_denoised(den, packed[1:2]), n=3)
    print(json.dumps({"workload": name, "frame": [2 * H, 2 * W], "align_ms_per_frame": t_ours,
                      "frames_per_s_alignment_only": 1e3 / t_ours, "torch_grid_sample_warps_only_ms": t_ref,
                      "online_flow_from_denoised_ms_single_pair": t_online}))
