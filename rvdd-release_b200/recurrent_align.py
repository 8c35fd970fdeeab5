"""The alignment half of the reference's recurrent inference step (models/recurrent_model.py:126-129, :233-324) as
one object: demosaic of the incoming packed frames, and, per frame, the backward warps that build the denoiser's input --

    warped previous denoised frame | noisy frame t | warped future frame(s)          -> netinput  [B, 3 (D + 1 + fD), 2H, 2W]
    warped recurrent feature maps                                                    -> featinput [B, 48, 2H, 2W]

-- written STRAIGHT into preallocated network input buffers (no torch.cat, no .clone(), no mask round trip to the
host, flow_utils.py:102), with the x2 bilinear upsampling of the half-resolution flows (recurrent_model.py:129) fused
into the gathers.  The denoiser itself is not part of this package: `step` returns the two buffers, the caller runs
its network and hands the results back through `update`.

Flows follow the dataset's layout (data/infer4rec_dataset.py:198-202): per frame [past flows ..., future flows ...],
each [2, H, W] at the packed-raw resolution, source -> target t.
"""
import torch

from . import bridge as _bridge
from .hamilton_adam import HamiltonAdam


class FrameAligner:
    def __init__(self, depth=1, future_depth=0, feature_channels=0, pattern="gbrg", predemosaic=True):
        """depth = model_patch_depth - 1 previous frames (D), future_depth = fD, feature_channels = channels ONE previous
        frame contributes to the recurrent feature map (48 with --feature_rec; the network sees D * 48)."""
        if depth < 1:
            raise ValueError("depth must be >= 1 (model_patch_depth >= 2)")
        self.D, self.fD, self.Cf = depth, future_depth, feature_channels
        self.ha = HamiltonAdam(pattern) if predemosaic else None
        self.br = _bridge.default_bridge()
        self.lastden = self.lastfeat = self.netinput = self.featinput = None

    def demosaic(self, packed):
        """[B, 4k, H, W] packed noisy frames in [-1, 1] -> [B, 3k, 2H, 2W] (recurrent_model.py:126)."""
        return self.ha(packed) if self.ha is not None else packed

    def reset(self, first_frames):
        """Start of a video (recurrent_model.py:233-245): the previous 'denoised' frames are the first D noisy frames
        ([B, D * C, H, W], oldest first), the recurrent features are zero."""
        B, DC, H, W = first_frames.shape
        if DC % self.D:
            raise ValueError("reset: expected D * C channels")
        C = DC // self.D
        self.lastden = [first_frames[:, b * C:(b + 1) * C] for b in range(self.D)]
        self.netinput = torch.empty((B, C * (self.D + 1 + self.fD), H, W), dtype=torch.float32, device=first_frames.device)
        if self.Cf:
            self.lastfeat = [torch.zeros((B, self.Cf, H, W), dtype=torch.float32, device=first_frames.device)
                             for _ in range(self.D)]
            self.featinput = torch.empty((B, self.D * self.Cf, H, W), dtype=torch.float32, device=first_frames.device)

    def step(self, noisy_t, flow_past, future=(), flow_future=()):
        """noisy_t [B, C, H, W]; flow_past: the D flows (previous frame b -> t, oldest first) as a list / a tensor
        [B, D, 2, h, w], or a single [B, 2, h, w] tensor when D == 1, at H x W or at half resolution;
        future: list of fD frames [B, C, H, W], flow_future: their flows (t+1+b -> t).
        Returns (netinput, featinput or None), views of buffers that the next call overwrites."""
        if self.lastden is None:
            raise RuntimeError("call reset(first_frames) at the start of a video")
        if torch.is_tensor(flow_past):
            flow_past = [flow_past] if flow_past.dim() == 4 else [flow_past[:, b] for b in range(flow_past.shape[1])]
        if len(flow_past) != self.D:
            raise ValueError("expected %d past flow(s)" % self.D)
        if len(future) != self.fD or len(flow_future) != self.fD:
            raise ValueError("expected %d future frame(s) and flow(s)" % self.fD)
        C, D = noisy_t.shape[1], self.D
        mul = 2.0 if flow_past[0].shape[-1] * 2 == noisy_t.shape[-1] else 1.0    # upsample_factor_2(flow, multiply_by=2)
        w = self.br.warp
        for b in range(D):                                                                                      # :281-304
            w(self.lastden[b], flow_past[b], "bicubic", flow_mul=mul, want_mask=False, out=self.netinput[:, b * C:(b + 1) * C])
            if self.Cf:                                                                                         # :290-297
                w(self.lastfeat[b], flow_past[b], "bicubic", flow_mul=mul, want_mask=False,
                  out=self.featinput[:, b * self.Cf:(b + 1) * self.Cf])
        self.netinput[:, D * C:(D + 1) * C].copy_(noisy_t)                                                      # :311
        for b, (fr, fl) in enumerate(zip(future, flow_future)):                                                 # :314-324
            w(fr, fl, "bicubic", flow_mul=mul, want_mask=False, out=self.netinput[:, (D + 1 + b) * C:(D + 2 + b) * C])
        return self.netinput, self.featinput

    def update(self, denoised, features=None):
        """Feed the network's outputs back (recurrent_model.py:335-345): the oldest previous frame / feature map drops out."""
        self.lastden = self.lastden[1:] + [denoised]
        if self.Cf:
            if features is None:
                raise ValueError("feature recurrence needs this frame's features")
            self.lastfeat = self.lastfeat[1:] + [features]
