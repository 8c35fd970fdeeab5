"""Mirror of the reference's ``library.CPPbridge`` (library.py:143-175) on top of the CUDA libBridge.so.

Same class name, constructor argument and ``TVL1_flow(Im1, Im2)`` contract, so code written against the reference
(util/flow_utils.py:128-132, data/axel4rec_dataset.py:65) runs unchanged: host numpy images in, ``(h, w, 2)`` float32
flow out, ``[..., 0]`` = x-displacement, ``[..., 1]`` = y-displacement, with Im2(x + flow(x)) ~ Im1(x).
``TVL1_flow_cuda`` is the tensors-in/tensors-out sibling used when the frames already live on the GPU.
"""
import ctypes
import os

import numpy as np

from . import bridge as _bridge


def warpedimagefile(wfolder, fromCode, toCode):
    """Flow / warped file naming of the reference cache (library.py:140-141)."""
    return os.path.join(wfolder, fromCode + '_' + toCode + '.tif')


def _rgb2gray(img):
    # skimage.color.rgb2gray weights (library.py:163-164); skimage itself is not a dependency here
    w = np.array([0.2125, 0.7154, 0.0721], dtype=img.dtype if img.dtype == np.float32 else np.float64)
    return img[..., :3] @ w


class CPPbridge(object):
    def __init__(self, libpath=None):
        # library.py:145-148: dlopen + the tvl1flow prototype.  All prototypes are declared by load_library.
        self.libBridge = _bridge.load_library(libpath)

    def TVL1_flow(self, Im1, Im2):
        """Im1 (target) and Im2 (source) are H x W x C images, C in {1, 3, 4}; gray conversion as library.py:162-170."""
        h, w = Im1.shape[:2]
        h1, w1 = Im2.shape[:2]
        assert h1 == h and w1 == w, "Both images Im1 and Im2 are supposed to share same size"

        if (Im1.ndim == 3 and Im1.shape[2] == 4 and Im2.shape[2] == 4 and Im1.dtype == np.float32
                and Im2.dtype == np.float32):
            # packed raw frames (the pipeline's case, data/base_dataset.py:159-178): upload the frames and take the mean of
            # the 4 channels on the GPU -- the same float32 ((a+b)+c)+d)/4 numpy computes (library.py:165-167), without the
            # ~25 ms the two np.mean calls cost at 1280x720
            import torch
            b = _bridge.default_bridge()
            pair = torch.stack((torch.from_numpy(np.ascontiguousarray(Im1)).to(b.device),
                                torch.from_numpy(np.ascontiguousarray(Im2)).to(b.device)))
            flow = b.tvl1_flow(b.gray(pair), src=[1], tgt=[0], check=True)
            return flow[0].cpu().numpy().transpose(1, 2, 0)

        I1 = np.zeros(h * w, dtype=ctypes.c_float)
        I2 = np.zeros(h * w, dtype=ctypes.c_float)
        flow = np.zeros(2 * h * w, dtype=ctypes.c_float)

        if Im1.shape[2] == 3:
            I1[:] = _rgb2gray(Im1).flatten()[:]
            I2[:] = _rgb2gray(Im2).flatten()[:]
        elif Im1.shape[2] == 4:
            I1[:] = np.mean(Im1, axis=2).flatten()[:]
            I2[:] = np.mean(Im2, axis=2).flatten()[:]
        elif Im1.shape[2] == 1:
            I1[:] = Im1.flatten()[:]
            I2[:] = Im2.flatten()[:]

        self.libBridge.tvl1flow(I1.ctypes.data, I2.ctypes.data, flow.ctypes.data, ctypes.c_int(w), ctypes.c_int(h))
        return flow.reshape(2, h, w).transpose(1, 2, 0)

    @staticmethod
    def TVL1_flow_cuda(frames, src, tgt, trace=False):
        """frames: CUDA float32 [n, h, w, c] packed frames -> flow [npairs, 2, h, w] on the device
        (pair k: target frames[tgt[k]], source frames[src[k]])."""
        b = _bridge.default_bridge()
        return b.tvl1_flow(b.gray(frames), src, tgt, trace=trace)
