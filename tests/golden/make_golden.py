"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF (only works in the build
container, where /root/reference exists; the committed .npz files are what travels).

* tvl1_*.npz   : inputs + flow of the unmodified reference C (oracle/_ref/libref_serial.so, symbol tvl1flow of
                 libBridge.cpp:44) and the per-stage outputs of its exported functions, on small seeded inputs.
* demosaic_*.npz: inputs + outputs of the reference's util/Hamilton_Adam_demo.py (HamiltonAdam.forward / remosaick).
* warp_*.npz   : inputs + outputs of the reference's own util/flow_utils.py (warp, upsample_factor_2) imported from
                 /root/reference and run on the CPU with this image's torch.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import RefLib  # noqa: E402
from rvdd_release_b200 import synth  # noqa: E402


def tvl1_golden():
    R = RefLib("serial")
    for name, (h, w, iso) in {"tvl1_64x96_iso3200": (64, 96, "iso3200"), "tvl1_45x80_clean": (45, 80, "clean"),
                              "tvl1_50x67_iso12800": (50, 67, "iso12800")}.items():
        I0, I1 = synth.gray_pair(h, w, iso)
        flow = R.tvl1flow(I0, I1)
        a, b = R.normalize(I0, I1)
        g = R.gaussian(a, 0.8)
        z = R.zoom_out(g)
        zi = R.zoom_in(z, w, h)
        dx, dy = R.centered_gradient(g)
        bw = R.bicubic_warp(g, flow[0], flow[1], True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), I0=I0, I1=I1, flow=flow, norm0=a, norm1=b, gauss08=g,
                            zoom_out=z, zoom_in=zi, cgx=dx, cgy=dy, bicubic_warp=bw)
        print(name, flow.reshape(2, -1).mean(1))


def warp_golden():
    sys.path.insert(0, "/root/reference")
    from util import flow_utils as ref  # the reference's own module
    g = torch.Generator().manual_seed(0)
    for name, (B, C, H, W, amp) in {"warp_c3": (2, 3, 24, 40, 3.0), "warp_c48": (1, 48, 16, 24, 6.0),
                                    "warp_c4_border": (1, 4, 20, 20, 30.0)}.items():
        x = torch.randn(B, C, H, W, generator=g)
        flow = amp * torch.randn(B, 2, H, W, generator=g)
        yb, mb = ref.warp(x, flow, "bicubic")
        yl, ml = ref.warp(x, flow, "bilinear")
        half = amp * torch.randn(B, 2, H // 2, W // 2, generator=g)
        up = ref.upsample_factor_2(half, multiply_by=2)
        yh, _ = ref.warp(x, up, "bicubic")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), flow=flow.numpy(), bicubic=yb.numpy(),
                            bilinear=yl.numpy(), mask=mb.numpy(), half_flow=half.numpy(), up2x2=up.numpy(),
                            bicubic_half=yh.numpy())
        print(name, float(yb.abs().mean()))


def demosaic_golden():
    """util/Hamilton_Adam_demo.py of the reference, imported and run on the CPU."""
    sys.path.insert(0, "/root/reference")
    from util.Hamilton_Adam_demo import HamiltonAdam
    g = torch.Generator().manual_seed(3)
    for name, (B, k, H, W, pattern) in {"demosaic_gbrg": (2, 1, 18, 26, "gbrg"), "demosaic_gbrg_2frames": (1, 2, 9, 35, "gbrg"),
                                        "demosaic_rggb": (1, 1, 11, 8, "rggb"), "demosaic_grbg": (1, 1, 6, 7, "grbg"),
                                        "demosaic_bggr": (1, 1, 5, 16, "bggr")}.items():
        # smooth image + noise in the network's [-1, 1] range, sampled through the CFA
        yy, xx = torch.meshgrid(torch.arange(2 * H, dtype=torch.float32), torch.arange(2 * W, dtype=torch.float32), indexing="ij")
        base = 0.6 * torch.sin(0.21 * xx + 0.13 * yy) * torch.cos(0.17 * yy - 0.05 * xx)
        x = torch.empty(B, 4 * k, H, W)
        for b in range(B):
            for j in range(4 * k):
                x[b, j] = base[(j % 4) // 2::2, (j % 4) % 2::2] + 0.1 * torch.randn(H, W, generator=g) + 0.05 * (j // 4)
        ha = HamiltonAdam(pattern)
        with torch.no_grad():
            y = ha(x)
            r = ha.remosaick(y[:, :3])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), y=y.numpy(), remosaick=r.numpy(),
                            pattern=np.array(pattern))
        print(name, float(y.abs().mean()))


if __name__ == "__main__":
    tvl1_golden()
    warp_golden()
    demosaic_golden()
