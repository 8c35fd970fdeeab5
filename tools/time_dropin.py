"""Latency of the reference's own call path through the drop-in symbol: library.CPPbridge(...).TVL1_flow(Im1, Im2)
(ctypes -> tvl1flow(I0, I1, u, nx, ny) with host buffers), one 1280x720 pair at a time, as data/base_dataset.py does."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import synth
from rvdd_release_b200.library import CPPbridge
seq = synth.sequence(3, 720, 1280, "iso3200").numpy()
b = CPPbridge()
ts = []
for i in range(6):
    t0 = time.perf_counter()
    f = b.TVL1_flow(seq[1 + i % 2], seq[i % 2])
    ts.append(time.perf_counter() - t0)
print("TVL1_flow per call (s):", [round(t, 4) for t in ts], "flow mean", float(np.abs(f).mean()))
