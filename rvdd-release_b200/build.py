"""Build libBridge.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

    python rvdd-release_b200/build.py          # builds rvdd-release_b200/lib/libBridge.so and ./build/libBridge.so

The second copy sits where the reference code looks for it ('./build/libBridge.so' relative to the working
directory: util/flow_utils.py:129,145, data/axel4rec_dataset.py:65 in the reference).  Both are git-ignored and
travel to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SOURCES = ["bridge.cu", "prep.cu", "solver.cu", "warp.cu", "demosaic.cu", "selftest.cu"]
LIB = os.path.join(PKG, "lib", "libBridge.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build(lib=None):
    lib = lib or LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "rvdd_bridge.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False, variant=None, defines=()):
    """Compile csrc/*.cu -> lib/libBridge.so (sm_100a).  Returns the library path.

    variant/defines build an experimental lib/libBridge_<variant>.so with extra -D macros (tuning runs only; select
    it at run time with RVDD_BRIDGE_LIB)."""
    if variant:
        vlib = os.path.join(PKG, "lib", "libBridge_%s.so" % variant)
        if not force and os.path.exists(vlib) and not needs_build(vlib):
            return vlib
        return _build(vlib, os.path.join(PKG, "lib", "obj_" + variant), verbose, ["-D" + d for d in defines], install=False)
    if not force and not needs_build():
        return LIB
    return _build(LIB, os.path.join(PKG, "lib", "obj"), verbose, [], install=True)


def _build(LIB, objdir, verbose, extra, install):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
        objs.append(obj)
    tmp = LIB + ".tmp"
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs, check=True)
    os.replace(tmp, LIB)
    if install:
        os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
        shutil.copy2(LIB, os.path.join(ROOT, "build", "libBridge.so"))
    return LIB


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    _defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=_variant, defines=_defs))
