"""Flow throughput at the other pipeline geometries (config 5: 3840x2160; 1920x1080; the packed REDS geometry 640x360):
python tools/time_sizes.py"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import bridge, synth
br = bridge.default_bridge()
for (h, w, npairs) in [(360, 640, 148), (360, 640, 116), (1080, 1920, 14), (2160, 3840, 8)]:
    frames = synth.sequence(npairs + 1, h, w, "iso3200", device="cuda")
    src, tgt = np.arange(npairs, dtype=np.int32), np.arange(1, npairs + 1, dtype=np.int32)
    def step():
        return br.tvl1_flow(br.gray(frames), src, tgt)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    br.check()
    print(json.dumps({"fused_kernel": br.last_solver_fused(), "fuse_min_px": os.environ.get("RVDD_FUSE_MIN_PX", "default"), "size": [w, h], "pairs_per_batch": npairs, "ms_per_batch": ms, "pairs_per_s": npairs / ms * 1e3,
                      "mpix_per_s": npairs * h * w / ms / 1e3}))
