"""Builds libunitspeech_b200.so (sm_100a only) in-tree with nvcc.

The shared library is the product; there is no other backend.  ``build_library()`` is what
``__graft_entry__.build()`` runs (nvcc cross-compiles without a GPU).
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libunitspeech_b200.so")
SOURCES = ["conv_igemm.cu", "elementwise.cu", "attention.cu", "engine.cu", "vocoder.cu", "train_kernels.cu", "wgrad_tc.cu", "frontend.cu"]
HEADERS = ["ptx.cuh", "pdl.h", "conv_igemm.h", "kernels.h", "train.h", os.path.join("..", "..", "include", "unitspeech_b200.h"),
           os.path.join("..", "..", "include", "unitspeech_b200_train.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(build_dir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True)
        objs.append(o)
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        _write_build_info(nvcc)
    return LIB_PATH


BUILD_INFO = os.path.join(LIB_DIR, "build_info.json")


def _write_build_info(nvcc: str) -> None:
    """Record of what the shipped .so was built from (bench.py quotes it as `build`): compiler, flags, source digests."""
    import hashlib
    import json
    import time
    ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    files = SOURCES + [h for h in HEADERS]
    digest = hashlib.sha256()
    for f in sorted(set(files)):
        path = os.path.normpath(os.path.join(CSRC, f))
        if os.path.exists(path):
            with open(path, "rb") as fh:
                digest.update(fh.read())
    info = {"nvcc": ver, "flags": NVCC_FLAGS, "sources": sorted(set(SOURCES)), "sources_sha256": digest.hexdigest(),
            "built_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "host": os.uname().nodename}
    with open(BUILD_INFO, "w") as f:
        json.dump(info, f, indent=1)


def build_info() -> dict:
    import json
    try:
        with open(BUILD_INFO) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return {}


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
