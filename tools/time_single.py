"""Where a single 1280x720 pair spends its time (whole GPU on one pair): python tools/time_single.py [never|always|auto]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import bridge, synth
br = bridge.default_bridge()
if len(sys.argv) > 1:
    br.set_fuse(sys.argv[1])
for K in (1, 2, 4, 8):
    frames = synth.sequence(K + 1, 720, 1280, "iso3200", device="cuda")
    gray = br.gray(frames)
    src, tgt = np.arange(K, dtype=np.int32), np.arange(1, K + 1, dtype=np.int32)
    for _ in range(2):
        br.tvl1_flow(gray, src, tgt)
    br.profile(True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        br.tvl1_flow(gray, src, tgt)
    b.record(); torch.cuda.synchronize()
    print("K=%d  %.2f ms per call; per level" % (K, a.elapsed_time(b) / 3), [round(x, 2) for x in br.profile_scales()],
          "wc", [round(x, 2) for x, _ in br.profile_phases()], "it", [round(y, 2) for _, y in br.profile_phases()])
    br.profile(False)
