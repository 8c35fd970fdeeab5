"""Seeded synthetic packed-raw frames for parity tests and the bench (no dataset, no network).

A band-limited texture (24 sinusoids, seed 1) is sampled at the four Bayer sub-pixel phases to give
``(H, W, 4)`` packed-raw-like frames, advected from frame to frame by a smooth, bounded motion of about 3 px,
scaled to the 12-bit raw range of the reference's unprocessing script and corrupted with its heteroscedastic
Gaussian noise model (dataset/generate_raw_from_RGB.py:174-189 in the reference):

    ISO 3200 : range [266, 3610], sigma^2 = 8.0034 x - 2043.51144
    ISO 12800: range [268, 4075], sigma^2 = 28.3015 x - 6307.62081      (variance clipped at 0)

Everything is torch so the same code runs on the CPU (tests) and on the GPU (bench input synthesis); it is input
generation only and is never part of a timed region.
"""
import math

import numpy as np
import torch

ISO = {
    "clean": dict(lo=266.0, hi=3610.0, a=0.0, b=0.0),
    "iso3200": dict(lo=266.0, hi=3610.0, a=8.0034, b=-2043.51144),
    "iso12800": dict(lo=268.0, hi=4075.0, a=28.3015, b=-6307.62081),
}
_PHASES = ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (0.5, 0.5))   # (dx, dy) of the 4 packed channels


def _texture_params(seed=1, n=24):
    r = np.random.RandomState(seed)
    return (r.uniform(0.2, 1.0, n), r.uniform(-0.15, 0.15, n), r.uniform(-0.15, 0.15, n),
            r.uniform(0.0, 2 * math.pi, n))


def _texture(x, y, params):
    amp, fx, fy, ph = params
    out = torch.zeros_like(x)
    for a, kx, ky, p in zip(amp, fx, fy, ph):
        out += a * torch.sin(kx * x + ky * y + p)
    return out


def _positions(t, h, w, device):
    """Texture coordinates seen by packed pixel (x, y) of frame t (float64 grids of shape (h, w))."""
    y, x = torch.meshgrid(torch.arange(h, dtype=torch.float64, device=device),
                          torch.arange(w, dtype=torch.float64, device=device), indexing="ij")
    px = x - 2.5 * t - 4.0 * torch.sin(y / 97.0 + 0.35 * t)
    py = y + 1.5 * t - 3.0 * torch.cos(x / 131.0 + 0.25 * t)
    return px, py


def clean_frame(t, h, w, device="cpu", seed=1):
    """Noise-free texture of frame t at the 4 sub-pixel phases, (h, w, 4) float64, arbitrary units."""
    params = _texture_params(seed)
    px, py = _positions(float(t), h, w, device)
    return torch.stack([_texture(px + dx, py + dy, params) for dx, dy in _PHASES], dim=-1)


def sequence(n_frames, h, w, iso="iso3200", device="cpu", seed=1, noise_seed=0):
    """``(n_frames, h, w, 4)`` float32 packed-raw-like frames.

    The raw scaling is fixed by frame 0 (so brightness is constant over the sequence); noise is fresh per frame
    (generator seeded with ``noise_seed + frame``)."""
    cfg = ISO[iso]
    f0 = clean_frame(0, h, w, device, seed)
    lo, hi = float(f0.min()), float(f0.max())
    frames = []
    for t in range(n_frames):
        f = f0 if t == 0 else clean_frame(t, h, w, device, seed)
        raw = cfg["lo"] + (f - lo) / (hi - lo) * (cfg["hi"] - cfg["lo"])
        raw = raw.clamp_(0.0, 4095.0)
        if cfg["a"] != 0.0:
            g = torch.Generator(device=device)
            g.manual_seed(noise_seed + t)
            var = (cfg["a"] * raw + cfg["b"]).clamp_(min=0.0)
            raw = raw + var.sqrt() * torch.randn(raw.shape, dtype=torch.float64, device=device, generator=g)
        frames.append(raw.to(torch.float32))
    return torch.stack(frames, 0)


def gray_pair(h, w, iso="iso3200", t=1, device="cpu", seed=1, noise_seed=0):
    """(I0, I1) = mean-of-4 gray of frame t (target) and frame t-1 (source), float32 (h, w) numpy arrays --
    the two images library.py:165-167 hands to tvl1flow."""
    seq = sequence(t + 1, h, w, iso, device, seed, noise_seed)
    g = seq.cpu().numpy().mean(axis=3, dtype=np.float32)
    return np.ascontiguousarray(g[t]), np.ascontiguousarray(g[t - 1])


def exact_gray_pair(h, w, seed=7, noise=True):
    """(I0, I1) float32 gray images of a noisy textured pair with ~2-4 px smooth motion, built ONLY from operations that
    are bit-reproducible on every IEEE-754 machine (PCG64 integers, float64 + - * /, floor, abs, sqrt -- no sin / exp /
    log, whose last bit depends on the libm / SIMD path of the host).  Used for golden vectors of geometries too large to
    run the oracle next to the GPU test (tests/golden/large_tvl1_2160x3840_exact.npz): the inputs are regenerated on the GPU
    box and must hash to the committed value."""
    rng = np.random.Generator(np.random.PCG64(seed))
    y, x = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")

    def tri(v):                                         # triangle wave in [0, 1], period 2
        return np.abs(v - 2.0 * np.floor(v / 2.0) - 1.0)

    def value_noise(cy, cx, cell, grid):
        gy, gx = cy / cell, cx / cell
        iy, ix = np.floor(gy), np.floor(gx)
        fy, fx = gy - iy, gx - ix
        iy = np.clip(iy.astype(np.int64), 0, grid.shape[0] - 2)
        ix = np.clip(ix.astype(np.int64), 0, grid.shape[1] - 2)
        a, b, c, d = grid[iy, ix], grid[iy, ix + 1], grid[iy + 1, ix], grid[iy + 1, ix + 1]
        top = a + fx * (b - a)
        bot = c + fx * (d - c)
        return top + fy * (bot - top)

    octaves = [(64.0, 1.0), (16.0, 0.5), (6.0, 0.25)]
    grids = [rng.integers(0, 1024, size=(int(h // c) + 12, int(w // c) + 12)).astype(np.float64) for c, _ in octaves]
    amp_sum = sum(a for _, a in octaves)

    def frame(moved):
        cy, cx = y + 16.0, x + 16.0
        if moved:                                       # the source frame: texture seen through a smooth displacement
            cx = cx - (2.25 + 1.5 * tri(y / 97.0))
            cy = cy - (-1.5 + 1.0 * tri(x / 131.0))
        t = np.zeros((h, w), np.float64)
        for (cell, a), g in zip(octaves, grids):
            t = t + a * value_noise(cy, cx, cell, g)
        return 266.0 + t / (1023.0 * amp_sum) * (3610.0 - 266.0)

    out = []
    for moved in (False, True):
        raw = frame(moved)
        if noise:                                       # ISO-3200-like heteroscedastic noise after the mean of 4 channels
            var = np.maximum(8.0034 * raw - 2043.51144, 0.0) / 4.0
            s = rng.integers(0, 65536, size=(4, h, w)).astype(np.float64).sum(axis=0) - 2.0 * 65535.0
            raw = raw + np.sqrt(var) * (s / 37837.0)    # sum of 4 uniforms: standard deviation 65536 / sqrt(3) = 37837
        out.append(np.ascontiguousarray(raw.astype(np.float32)))
    return out[0], out[1]


# ------------------------------------------------------------------------------------------------ exact sequences
#
# Sequences for fixtures that must be regenerated bit for bit on another machine AND on another device (golden vectors of
# the full-size pipeline configurations are made on the CPU of the build container and replayed on the GPU box): only
# int64 hashing and IEEE float64 + - * / floor sqrt, each as its own torch op (no fused multiply-add, no libm).

_M64 = (1 << 64) - 1


def _wrap(v):
    """Python int -> the int64 value with the same 64-bit pattern."""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, k):
    """logical right shift of an int64 tensor"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def _hash64(x):
    """splitmix64 finaliser on an int64 tensor (wrap-around arithmetic)."""
    x = x + _wrap(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _wrap(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _wrap(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _div(a, c):
    """a / c as a true IEEE division on every device.  (On CUDA, torch turns tensor / python-scalar into a multiplication
    by the reciprocal, which is not the same number; dividing by a 0-dim tensor ON THE SAME DEVICE takes the ordinary
    element-wise division kernel.)"""
    return a / torch.tensor(c, dtype=a.dtype, device=a.device)


def _tri(v):
    """triangle wave in [0, 1] with period 2 (exact operations only; v / 2 is exact either way)"""
    return torch.abs(v - 2.0 * torch.floor(v / 2.0) - 1.0)


def _value_noise(cy, cx, cell, salt):
    """bilinear interpolation of hashed lattice values in [0, 1023] at float64 coordinates (cy, cx)"""
    gy, gx = _div(cy, cell), _div(cx, cell)
    iy, ix = torch.floor(gy), torch.floor(gx)
    fy, fx = gy - iy, gx - ix
    iy, ix = iy.to(torch.int64), ix.to(torch.int64)

    def lat(dy, dx):
        key = (iy + dy) * 1000003 + (ix + dx) * 7919 + salt * 104729
        return (_hash64(key) & 1023).to(torch.float64)

    a, b, c, d = lat(0, 0), lat(0, 1), lat(1, 0), lat(1, 1)
    top = a + fx * (b - a)
    bot = c + fx * (d - c)
    return top + fy * (bot - top)


_OCTAVES = ((64.0, 1.0), (16.0, 0.5), (6.0, 0.25))


def exact_clean_frame(t, h, w, device="cpu"):
    """Noise-free frame t at the 4 Bayer sub-pixel phases, (h, w, 4) float64 in [0, 1]."""
    y, x = torch.meshgrid(torch.arange(h, dtype=torch.float64, device=device),
                          torch.arange(w, dtype=torch.float64, device=device), indexing="ij")
    tt = float(t)
    px = x - 2.5 * tt - 4.0 * _tri(_div(y, 97.0) + 0.35 * tt) + 4096.0
    py = y + 1.5 * tt - 3.0 * _tri(_div(x, 131.0) + 0.25 * tt) + 4096.0
    out = []
    for dx, dy in _PHASES:
        v = torch.zeros_like(x)
        for k, (cell, amp) in enumerate(_OCTAVES):
            v = v + amp * _value_noise(py + dy, px + dx, cell, k)
        out.append(_div(v, 1023.0 * sum(a for _, a in _OCTAVES)))
    return torch.stack(out, dim=-1)


def exact_sequence(n_frames, h, w, iso="iso3200", device="cpu", noise_seed=0):
    """``(n_frames, h, w, 4)`` float32 packed-raw-like frames, bit-identical on every machine and device: hashed value-noise
    texture advected by a smooth ~3 px motion, raw range and heteroscedastic noise model of ``ISO[iso]`` (the noise is a
    sum of four 16-bit uniforms per sample instead of a Gaussian)."""
    cfg = ISO[iso]
    frames = []
    idx = torch.arange(h * w * 4, dtype=torch.int64, device=device).reshape(h, w, 4)
    for t in range(n_frames):
        raw = cfg["lo"] + exact_clean_frame(t, h, w, device) * (cfg["hi"] - cfg["lo"])
        if cfg["a"] != 0.0:
            r = _hash64(idx + (noise_seed * 1000 + t) * 1000000007)
            s = ((r & 65535) + (_lsr(r, 16) & 65535) + (_lsr(r, 32) & 65535) + _lsr(r, 48)).to(torch.float64) - 2.0 * 65535.0
            var = torch.clamp(cfg["a"] * raw + cfg["b"], min=0.0)
            raw = raw + torch.sqrt(var) * _div(s, 37837.0)
        frames.append(torch.clamp(raw, 0.0, 4095.0).to(torch.float32))
    return torch.stack(frames, 0)
