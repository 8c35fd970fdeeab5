"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float32) of the reference's Hamilton-Adams demosaicking and
remosaicking (util/Hamilton_Adam_demo.py), used by tests/ as the checker for the CUDA kernel in csrc/demosaic.cu.
Nothing under rvdd-release_b200/ may import this module.

Pinned: equal (to float32 rounding of the convolution sums, <= 1e-6) to the reference's own HamiltonAdam module
imported from /root/reference and run on the CPU -- golden vectors tests/golden/demosaic_*.npz made by
tests/golden/make_golden.py.  The reference has no test or known-answer vector for this code.
"""
import numpy as np

_CH = {"r": 0, "g": 1, "b": 2}


def pack_in_one(x):
    """[B, 4, H, W] -> [B, 2H, 2W] (Hamilton_Adam_demo.py:226-234)."""
    B, _, H, W = x.shape
    y = np.zeros((B, 2 * H, 2 * W), dtype=x.dtype)
    y[:, 0::2, 0::2] = x[:, 0]
    y[:, 0::2, 1::2] = x[:, 1]
    y[:, 1::2, 0::2] = x[:, 2]
    y[:, 1::2, 1::2] = x[:, 3]
    return y


def bayer_mask(pattern, H, W):
    """[3, H, W] colour masks (:201-224)."""
    m = np.zeros((3, H, W), dtype=np.float32)
    m[_CH[pattern[0]], 0::2, 0::2] = 1
    m[_CH[pattern[1]], 0::2, 1::2] = 1
    m[_CH[pattern[2]], 1::2, 0::2] = 1
    m[_CH[pattern[3]], 1::2, 1::2] = 1
    return m


def algo2_mask(pattern, H, W):
    """(maskGr, maskGb): green samples on the red / blue rows (:175-199)."""
    gr, gb = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    if pattern == "grbg":
        gr[0::2, 0::2] = 1; gb[1::2, 1::2] = 1
    elif pattern == "rggb":
        gr[0::2, 1::2] = 1; gb[1::2, 0::2] = 1
    elif pattern == "gbrg":
        gb[0::2, 0::2] = 1; gr[1::2, 1::2] = 1
    elif pattern == "bggr":
        gb[0::2, 1::2] = 1; gr[1::2, 0::2] = 1
    else:
        raise ValueError(pattern)
    return gr, gb


def _sh(p, pad, dy, dx):
    """view of the replication-padded image shifted by (dy, dx)"""
    H, W = p.shape[-2] - 2 * pad, p.shape[-1] - 2 * pad
    return p[..., pad + dy:pad + dy + H, pad + dx:pad + dx + W]


def _pad(a, n):
    return np.pad(a, [(0, 0)] * (a.ndim - 2) + [(n, n), (n, n)], mode="edge")          # nn.ReplicationPad2d


def algo1(raw, green_mask):
    """green plane (:123-142); raw [B, 2H, 2W]."""
    f = np.float32
    p = _pad(raw, 2)
    kh = f(.5) * _sh(p, 2, 0, -1) + f(.5) * _sh(p, 2, 0, 1)
    kv = f(.5) * _sh(p, 2, -1, 0) + f(.5) * _sh(p, 2, 1, 0)
    dh = (_sh(p, 2, 0, -2) + f(-2.) * raw) + _sh(p, 2, 0, 2)
    dv = (_sh(p, 2, -2, 0) + f(-2.) * raw) + _sh(p, 2, 2, 0)
    fh = _sh(p, 2, 0, -1) - _sh(p, 2, 0, 1)
    fv = _sh(p, 2, -1, 0) - _sh(p, 2, 1, 0)
    rawh, rawv = kh - dh / f(4), kv - dv / f(4)
    clh, clv = np.abs(fh) + np.abs(dh), np.abs(fv) + np.abs(dv)
    s = np.sign(clh - clv)
    green = (1 + s) * rawv / f(2) + (1 - s) * rawh / f(2)
    return green * (1 - green_mask) + raw * green_mask


def algo2(green, chan, mask_ochan, mask_gr, mask_gb):
    """red or blue plane (:145-172); chan = CFA masked to that colour."""
    f = np.float32
    c, g = _pad(chan, 1), _pad(green, 1)
    kh = f(.5) * _sh(c, 1, 0, -1) + f(.5) * _sh(c, 1, 0, 1)
    kv = f(.5) * _sh(c, 1, -1, 0) + f(.5) * _sh(c, 1, 1, 0)
    kp = f(.5) * _sh(c, 1, -1, -1) + f(.5) * _sh(c, 1, 1, 1)
    kn = f(.5) * _sh(c, 1, -1, 1) + f(.5) * _sh(c, 1, 1, -1)
    fp = -_sh(c, 1, -1, -1) + _sh(c, 1, 1, 1)
    fn = -_sh(c, 1, -1, 1) + _sh(c, 1, 1, -1)
    gdh = (f(.25) * _sh(g, 1, 0, -1) + f(-.5) * green) + f(.25) * _sh(g, 1, 0, 1)
    gdv = (f(.25) * _sh(g, 1, -1, 0) + f(-.5) * green) + f(.25) * _sh(g, 1, 1, 0)
    gdp = (_sh(g, 1, -1, -1) + f(-2.) * green) + _sh(g, 1, 1, 1)
    gdn = (_sh(g, 1, -1, 1) + f(-2.) * green) + _sh(g, 1, 1, -1)
    ch = mask_gr * (kh - gdh)
    cv = mask_gb * (kv - gdv)
    cp = mask_ochan * (kp - gdp / f(4))
    cn = mask_ochan * (kn - gdn / f(4))
    clp = mask_ochan * (np.abs(fp) + np.abs(gdp))
    cln = mask_ochan * (np.abs(fn) + np.abs(gdn))
    s = np.sign(clp - cln)
    out = (1 + s) * cn / f(2) + (1 - s) * cp / f(2)
    return (out + ch + cv) + chan


def hamilton_adam(x, pattern="gbrg"):
    """HamiltonAdam(pattern).forward (:249-289): [B, 4k, H, W] float32 -> [B, 3k, 2H, 2W]."""
    x = np.asarray(x, dtype=np.float32)
    B0, c4, H, W = x.shape
    raw = pack_in_one(x.reshape(-1, 4, H, W))
    mask = bayer_mask(pattern, 2 * H, 2 * W)
    xm = raw[:, None] * mask[None]
    green = algo1(xm.sum(1, dtype=np.float32), mask[1])
    gr, gb = algo2_mask(pattern, 2 * H, 2 * W)
    red = algo2(green, xm[:, 0], mask[2], gr, gb)
    blue = algo2(green, xm[:, 2], mask[0], gb, gr)
    y = np.stack((red, green, blue), 1).astype(np.float32)
    return y.reshape(B0, -1, 2 * H, 2 * W)


def remosaick(x, pattern="gbrg"):
    """[B, 3, 2H, 2W] -> [B, 4, H, W] (:237-246; the reference hard-codes the gbrg order)."""
    c = [_CH[k] for k in pattern]
    return np.stack((x[:, c[0], 0::2, 0::2], x[:, c[1], 0::2, 1::2], x[:, c[2], 1::2, 0::2], x[:, c[3], 1::2, 1::2]), 1)
