"""Golden vector of the largest configured geometry (3840x2160, 9 scales) made by RUNNING THE COMPILED REFERENCE
(oracle/_ref/libref_omp.so, the unmodified /root/reference sources) in the build container -- about a minute of CPU there,
too long to sit next to the GPU test.  The inputs come from synth.exact_gray_pair (bit-reproducible on any IEEE machine) and
are NOT stored: the test regenerates them and checks their SHA-256 first.  Stored: hashes of inputs and flow, the
per-(scale, warp) iteration counts of the oracle port (which must equal the reference flow bit for bit, asserted here), a
24x-subsampled copy of the flow and float64 checksums.

    python tests/golden/make_large_golden.py
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import PortLib, RefLib  # noqa: E402
from rvdd_release_b200 import synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    for name, (h, w) in {"large_tvl1_2160x3840_exact": (2160, 3840)}.items():
        I0, I1 = synth.exact_gray_pair(h, w)
        t0 = time.time()
        flow = RefLib("omp").tvl1flow(I0, I1)
        t1 = time.time()
        pflow, iters, _, _, _ = PortLib().tvl1flow_traced(I0, I1, err_mode=0)
        t2 = time.time()
        assert np.array_equal(flow, pflow), "port and compiled reference disagree"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), h=h, w=w, sha_I0=sha(I0), sha_I1=sha(I1), sha_flow=sha(flow),
                            iters=iters, flow_sub=flow[:, ::24, ::24].copy(), sum_u=float(flow[0].sum(dtype=np.float64)),
                            sum_v=float(flow[1].sum(dtype=np.float64)), abs_mean=float(np.abs(flow).mean(dtype=np.float64)))
        print(name, "reference %.1f s, port %.1f s, iterations per scale %s" % (t1 - t0, t2 - t1, iters.sum(1).tolist()))


if __name__ == "__main__":
    main()
