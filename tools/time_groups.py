import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import bridge, synth
br = bridge.default_bridge()
h, w, npairs = 360, 640, 148
frames = synth.sequence(npairs + 1, h, w, "iso3200", device="cuda")
src, tgt = np.arange(npairs, dtype=np.int32), np.arange(1, npairs + 1, dtype=np.int32)
gray = br.gray(frames)
for g in (0, 29, 37, 49, 59, 74, 98, 148):
    br.set_groups(g)
    for _ in range(2):
        br.tvl1_flow(gray, src, tgt)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        br.tvl1_flow(gray, src, tgt)
    b.record(); torch.cuda.synchronize()
    print(g, round(npairs / (a.elapsed_time(b) / 3) * 1e3, 1), "pairs/s")
br.check()
