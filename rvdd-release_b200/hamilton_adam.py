"""Mirror of the reference's ``util/Hamilton_Adam_demo.py::HamiltonAdam`` on the CUDA bridge: same constructor,
``forward`` / ``__call__`` (packed Bayer raw ``[B, 4k, H, W]`` -> RGB ``[B, 3k, 2H, 2W]``), ``remosaick`` and
``pack_in_one``, so ``recurrentModel`` can use it in place of the module (models/recurrent_model.py:98,126,151).

The reference runs three fixed-weight convolutions and ~40 element-wise tensor ops per call; here ``forward`` is one
kernel (csrc/demosaic.cu).  CUDA tensors only, no CPU fallback, inference only.
"""
import torch

from . import bridge as _bridge

_PATTERNS = ("grbg", "rggb", "gbrg", "bggr")


class HamiltonAdam:
    def __init__(self, pattern):
        if pattern not in _PATTERNS:
            raise ValueError("pattern can be: 'grbg', 'rggb', 'gbrg', 'bggr' (got %r)" % (pattern,))
        self.pattern = pattern

    def to(self, *args, **kwargs):          # nn.Module-style plumbing used by recurrentModel.to_device
        return self

    def forward(self, x):
        """Hamilton-Adams demosaicing (Hamilton_Adam_demo.py:249-289)."""
        return _bridge.default_bridge().demosaic(x, self.pattern)

    __call__ = forward

    def pack_in_one(self, x):
        """[B, 4, H, W] -> [B, 2H, 2W] CFA image (:226-234)."""
        B, _, H, W = x.shape
        y = torch.empty((B, 2 * H, 2 * W), dtype=x.dtype, device=x.device)
        y[:, 0::2, 0::2] = x[:, 0]
        y[:, 0::2, 1::2] = x[:, 1]
        y[:, 1::2, 0::2] = x[:, 2]
        y[:, 1::2, 1::2] = x[:, 3]
        return y

    def remosaick(self, x):
        """RGB [B, 3, 2H, 2W] -> packed raw [B, 4, H, W] (:237-246; like the reference, always in 'gbrg' order)."""
        return torch.stack((x[:, 1, 0::2, 0::2], x[:, 2, 0::2, 1::2], x[:, 0, 1::2, 0::2], x[:, 1, 1::2, 1::2]), dim=1)
