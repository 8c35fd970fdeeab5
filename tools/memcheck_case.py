"""Small case for compute-sanitizer (one tool per gpurun call): both solver instantiations incl. the fused pass on every level,
the (H, W, 4) warp in both variants, the NCHW tile warp, the demosaic.  python tools/memcheck_case.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import bridge, synth
br = bridge.default_bridge()
frames = synth.sequence(4, 90, 160, "iso3200", device="cuda")
gray = br.gray(frames)
src, tgt = np.arange(3, dtype=np.int32), np.arange(1, 4, dtype=np.int32)
ref = None
for mode, px in (("never", -1), ("always", 0)):
    br.set_fuse(mode, px)
    flow = br.tvl1_flow(gray, src, tgt, check=True)
    ref = flow if ref is None else ref
    assert torch.equal(flow, ref)
x = frames[:3].permute(0, 3, 1, 2)
for tma in ("", "1"):
    if tma:
        os.environ["RVDD_WARP_TMA"] = "1"
    w, _ = br.warp(x, ref, "bicubic")
feat = torch.randn(1, 48, 90, 160, device="cuda")
br.warp(feat, ref[:1], "bicubic")
br.demosaic((frames / 4095.0 * 2 - 1).permute(0, 3, 1, 2).contiguous())
torch.cuda.synchronize()
print("memcheck case ok", float(ref.abs().mean()), float(w.abs().mean()))
