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
