"""B200-native frame alignment for RVDD: dual TV-L1 optical flow + flow-based backward warp.

Mirrors the reference's Python surface for this path (library.CPPbridge, util.flow_utils) on top of the
C-ABI CUDA library ``libBridge.so`` built from ``csrc/`` (see include/rvdd_bridge.h).
"""
__version__ = "0.1.0"
