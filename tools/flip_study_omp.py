"""Companion of tools/flip_study.py: how often do two builds / thread counts OF THE REFERENCE ITSELF disagree?  The OpenMP
reduction of the residual (tvl1flow_lib.c:211, float, `reduction(+:error)`) depends on the number of threads, so the unmodified
reference compiled with -fopenmp takes its stopping decisions from a different float sum than its serial build.  For the same
pairs as flip_study.py: flows of oracle/_ref/libref_serial.so vs libref_omp.so with OMP_NUM_THREADS = argv[2].

    python tools/flip_study_omp.py [pairs=120] [threads=8] >> profiles/flip_study_r02.txt
"""
import os
import sys

import numpy as np

N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
thr = sys.argv[2] if len(sys.argv) > 2 else "8"
os.environ["OMP_NUM_THREADS"] = thr
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import RefLib  # noqa: E402
from rvdd_release_b200 import synth  # noqa: E402

ser, omp = RefLib("serial"), RefLib("omp")
diff, worst = [], 0.0
for k in range(N):
    iso = ("clean", "iso3200", "iso12800")[k % 3]
    I0, I1 = synth.gray_pair(720, 1280, iso, t=1 + (k // 3) % 7, noise_seed=31 * k)
    a, b = ser.tvl1flow(I0, I1), omp.tvl1flow(I0, I1)
    if not np.array_equal(a, b):
        e = float(np.sqrt(((a - b) ** 2).sum(0)).mean())
        diff.append((k, iso, e))
        worst = max(worst, e)
print("# the reference against itself: serial build vs OpenMP build with %s threads, same %d pairs: %d pairs with different flows "
      "(a flipped stopping decision), max mean-EPE %.2e px: %s" % (thr, N, len(diff), worst, [(k, i, round(e, 5)) for k, i, e in diff]))
