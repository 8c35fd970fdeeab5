"""Time split of the solver on the bench workload (29 pairs 1280x720 ISO 3200): warp-constants phase vs iterations per level.
RVDD_BRIDGE_LIB selects a variant library (what-if builds give wrong flows; only their phase times mean anything)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import bridge, synth
br = bridge.default_bridge()
K = 29
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
ph = br.profile_phases()
print(os.environ.get("RVDD_BRIDGE_LIB", "default"), "%.2f ms per call" % (a.elapsed_time(b) / 3), "wc", [round(x, 2) for x, _ in ph],
      "sum %.2f" % sum(x for x, _ in ph), "it sum %.2f" % sum(y for _, y in ph), "fused", br.last_solver_fused())
