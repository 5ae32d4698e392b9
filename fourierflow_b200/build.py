"""Build libffno_b200.so in-tree with nvcc for sm_100a (no torch, no pybind: a plain C-ABI library).

    python -m fourierflow_b200.build          # rebuild if any source is newer than the .so
    python -m fourierflow_b200.build --force

The .so lands in fourierflow_b200/lib/ (git-ignored, but it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libffno_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "--expt-relaxed-constexpr", "-cudart", "static"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libffno_b200.so")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(HERE, "..", "include", "ffno_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:      # ranks of one torchrun build once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not stale():
            return LIB
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    timeline = os.environ.get("FFNO_TIMELINE") == "1"
    tag = os.environ.get("FFNO_BUILD_TAG", "")              # experiment builds: FFNO_BUILD_TAG=name FFNO_BUILD_DEFS="-DX=1 ..."
    extra = os.environ.get("FFNO_BUILD_DEFS", "").split() if tag else []
    suffix = ("_timeline" if timeline else "") + (f"_{tag}" if tag else "")
    lib = LIB[:-3] + suffix + ".so" if suffix else LIB      # diagnostics builds live beside the product, never replace it
    objdir = os.path.join(LIBDIR, "obj" + suffix)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc(), *ARCH, *FLAGS, "-c", src, "-o", obj]
        if timeline:                                    # in-kernel clock64 stamps (tools/*_timeline.py, FFNO_B200_LIB=...)
            cmd.insert(1, "-DFFNO_TIMELINE")
        for d in extra:
            cmd.insert(1, d)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(src)} ---\n{out}\n")
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see log above)")
    cmd = [nvcc(), *ARCH, "-shared", "-cudart", "static", "-o", lib, *objs]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
