"""Tiny driver for ncu captures of the persistent solver: python tools/prof_solver.py [pairs] [iso] [never|always|auto]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvdd_release_b200 import bridge, synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 29
iso = sys.argv[2] if len(sys.argv) > 2 else "iso3200"
br = bridge.default_bridge()
br.set_fuse(sys.argv[3] if len(sys.argv) > 3 else "auto")
frames = synth.sequence(K + 1, 720, 1280, iso, device="cuda")
gray = br.gray(frames)
src, tgt = np.arange(K, dtype=np.int32), np.arange(1, K + 1, dtype=np.int32)
for _ in range(3):
    flow = br.tvl1_flow(gray, src, tgt)
br.check()
print("ok", float(flow.abs().mean()), "fused" if br.last_solver_fused() else "single")
