"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/rvdd_bridge.h declares.
No compute call is made here (there is no GPU); the product path must refuse to run instead of falling back."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "rvdd_bridge.h")).read()
    return sorted(set(re.findall(r"RVDD_API\s+[\w\s\*]+?\b(\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(libpath):
    from rvdd_release_b200 import bridge
    names = _declared()
    assert "tvl1flow" in names and len(names) >= 14
    lib = bridge.load_library(libpath)
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(bridge.EXPORTED_SYMBOLS) == names          # the Python binding covers the whole header
    assert lib.rvdd_abi_version() == 1


def test_reference_binding_is_accepted(libpath):
    """The exact ctypes binding library.py:145-148 performs works on our library."""
    import ctypes
    lib = ctypes.cdll.LoadLibrary(libpath)
    lib.tvl1flow.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.tvl1flow.restype = None


def test_pyramid_geometry_matches_oracle(libpath, port):
    from rvdd_release_b200 import bridge
    lib = bridge.load_library(libpath)
    import ctypes as C
    for nx, ny in [(1280, 720), (640, 360), (1920, 1080), (3840, 2160), (131, 97), (45, 23), (64, 48)]:
        nxs, nys = (C.c_int * 16)(), (C.c_int * 16)()
        S = lib.rvdd_pyramid(nx, ny, None, nxs, nys)
        assert [(nxs[s], nys[s]) for s in range(S)] == port.pyramid_sizes(nx, ny)
    nxs, nys = (C.c_int * 16)(), (C.c_int * 16)()
    assert lib.rvdd_pyramid(1280, 720, None, nxs, nys) == 7 and (nxs[6], nys[6]) == (20, 12)


def test_no_cpu_fallback(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rvdd_release_b200 import bridge, flow_utils
    with pytest.raises(bridge.BridgeError):
        bridge.Bridge(libpath)
    with pytest.raises(bridge.BridgeError):
        flow_utils.warp(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8), "bicubic")
    import ctypes as C
    lib = bridge.load_library(libpath)
    h = C.c_void_p()
    assert lib.rvdd_create(C.byref(h)) != 0 and b"CUDA" in lib.rvdd_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rvdd-release_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "oracle/", "oracle.", "libtvl1_port", "libref_", "_ref/",
                               "hostsim"):
                    if needle == "hostsim" and f.endswith(".h"):
                        continue            # the headers only mention the host-compiled unit tests in a comment
                    assert needle not in src, (os.path.join(dirpath, f), needle)


def test_hostbind_is_best_effort():
    """NUMA pinning helper: parses cpulists, and never fails where there is no GPU / no sysfs entry."""
    from rvdd_release_b200 import hostbind
    assert hostbind._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostbind._parse_cpulist("") == set()
    info = hostbind.bind_to_gpu(0)
    assert info["gpu"] == 0 and isinstance(info["bound"], bool)
