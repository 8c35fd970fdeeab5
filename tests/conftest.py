import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import PortLib, build
    build()
    return PortLib()


@pytest.fixture(scope="session")
def reflib():
    """The unmodified reference compiled into oracle/_ref (skips where neither it nor /root/reference exist)."""
    from oracle.oracle import RefLib
    try:
        return RefLib("serial")
    except (FileNotFoundError, OSError, subprocess.CalledProcessError) as e:
        pytest.skip("reference build unavailable: %s" % e)


@pytest.fixture(scope="session")
def libpath():
    sys.path.insert(0, os.path.join(ROOT, "rvdd-release_b200"))
    import build as _build
    return _build.build_lib()


@pytest.fixture(scope="session")
def bridge(libpath):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rvdd_release_b200 import bridge as B
    return B.default_bridge()
