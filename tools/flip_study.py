"""Stop-rule study: the reference sums the per-pixel residuals of an iteration in FLOAT, in pixel order
(tvl1flow_lib.c:210-223; thread-count dependent under OpenMP); the CUDA solver sums the same float terms in DOUBLE in a fixed
order.  Does the decision `error > eps^2` ever differ?  For N synthetic 1280x720 pairs (clean / ISO 3200 / ISO 12800, different
frames and noise seeds) the oracle port is run with both sums (err_mode 0 = float pixel order, 1 = double) and compared:
iteration counts per (scale, warp), flows, and the relative difference of the two sums over all iterations.

    python tools/flip_study.py [pairs=120] [procs=8] > profiles/flip_study_r02.txt
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(k):
    import torch
    torch.set_num_threads(1)
    from oracle.oracle import PortLib
    from rvdd_release_b200 import synth
    iso = ("clean", "iso3200", "iso12800")[k % 3]
    t = 1 + (k // 3) % 7
    I0, I1 = synth.gray_pair(720, 1280, iso, t=t, noise_seed=31 * k)
    P = PortLib()
    f0, it0, _, ef, ed = P.tvl1flow_traced(I0, I1, err_mode=0, err_cap=4096)
    f1, it1, _, _, _ = P.tvl1flow_traced(I0, I1, err_mode=1)
    n = min(len(ef), len(ed))
    rel = np.abs(ef[:n].astype(np.float64) - ed[:n]) / np.maximum(ed[:n], 1e-300)
    near = np.abs(ed[:n] - 1e-4) / 1e-4                       # distance of the decision variable from eps^2
    epe = float(np.sqrt(((f0 - f1) ** 2).sum(0)).mean())
    return dict(k=k, iso=iso, t=t, flips=int((it0 != it1).sum()), iters=int(it0.sum()), equal=bool(np.array_equal(f0, f1)), epe=epe,
                rel_med=float(np.median(rel)), rel_max=float(rel.max()), closest=float(near.min()))


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    with mp.Pool(procs) as pool:
        rows = pool.map(one, range(N))
    print("# stop-rule study: float pixel-order sum (reference) vs fixed-order double sum (CUDA solver), %d pairs of 1280x720" % N)
    print("# pair iso frame  iterations  flipped(scale,warp)  flows_equal  meanEPE  rel.diff of sums: median max   closest |err-eps^2|/eps^2")
    for r in rows:
        print("%4d %-8s t=%d %6d %3d %-5s %.2e  %.1e %.1e  %.1e" % (r["k"], r["iso"], r["t"], r["iters"], r["flips"], r["equal"], r["epe"],
                                                                  r["rel_med"], r["rel_max"], r["closest"]))
    nf = sum(1 for r in rows if r["flips"])
    print("# pairs with at least one flipped decision: %d of %d; total inner-loop decisions: %d; max mean-EPE between the two: %.2e px"
          % (nf, N, sum(r["iters"] for r in rows), max(r["epe"] for r in rows)))
    print("# relative difference of the two sums over all iterations: median %.1e, max %.1e; the decision variable came within %.1e "
          "(relative) of eps^2 at the closest" % (float(np.median([r["rel_med"] for r in rows])), max(r["rel_max"] for r in rows),
                                                  min(r["closest"] for r in rows)))
