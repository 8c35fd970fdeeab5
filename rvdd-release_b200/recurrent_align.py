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
        """depth = model_patch_depth - 1 previous frames (D), future_depth = fD, feature_channels = 48 with --feature_rec."""
        if depth != 1:
            raise NotImplementedError("the shipped checkpoints use one previous frame (model_patch_depth 2)")
        self.D, self.fD, self.Cf = depth, future_depth, feature_channels
        self.ha = HamiltonAdam(pattern) if predemosaic else None
        self.br = _bridge.default_bridge()
        self.lastden = self.lastfeat = self.netinput = self.featinput = None

    def demosaic(self, packed):
        """[B, 4k, H, W] packed noisy frames in [-1, 1] -> [B, 3k, 2H, 2W] (recurrent_model.py:126)."""
        return self.ha(packed) if self.ha is not None else packed

    def reset(self, first_frame):
        """Start of a video (recurrent_model.py:233-245): the previous 'denoised' frame is the first noisy frame, the
        recurrent features are zero."""
        B, C, H, W = first_frame.shape
        self.lastden = first_frame
        self.netinput = torch.empty((B, C * (self.D + 1 + self.fD), H, W), dtype=torch.float32, device=first_frame.device)
        if self.Cf:
            self.lastfeat = torch.zeros((B, self.Cf, H, W), dtype=torch.float32, device=first_frame.device)
            self.featinput = torch.empty_like(self.lastfeat)

    def step(self, noisy_t, flow_past, future=(), flow_future=()):
        """noisy_t [B, C, H, W]; flow_past [B, 2, h, w] (previous frame -> t), at H x W or at half resolution;
        future: list of fD frames [B, C, H, W], flow_future: their flows (t+1+b -> t).
        Returns (netinput, featinput or None), views of buffers that the next call overwrites."""
        if self.lastden is None:
            raise RuntimeError("call reset(first_frame) at the start of a video")
        if len(future) != self.fD or len(flow_future) != self.fD:
            raise ValueError("expected %d future frame(s) and flow(s)" % self.fD)
        C = noisy_t.shape[1]
        mul = 2.0 if flow_past.shape[-1] * 2 == noisy_t.shape[-1] else 1.0       # upsample_factor_2(flow, multiply_by=2)
        w = self.br.warp
        w(self.lastden, flow_past, "bicubic", flow_mul=mul, want_mask=False, out=self.netinput[:, 0:C])        # :281-288
        self.netinput[:, C:2 * C].copy_(noisy_t)                                                                # :311
        for b, (fr, fl) in enumerate(zip(future, flow_future)):                                                 # :314-324
            w(fr, fl, "bicubic", flow_mul=mul, want_mask=False, out=self.netinput[:, (2 + b) * C:(3 + b) * C])
        if self.Cf:                                                                                             # :290-297
            w(self.lastfeat, flow_past, "bicubic", flow_mul=mul, want_mask=False, out=self.featinput)
        return self.netinput, self.featinput

    def update(self, denoised, features=None):
        """Feed the network's outputs back (recurrent_model.py:335-345)."""
        self.lastden = denoised
        if self.Cf:
            if features is None:
                raise ValueError("feature recurrence needs this frame's features")
            self.lastfeat = features
