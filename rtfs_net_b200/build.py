"""Builds rtfs_net_b200/lib/librtfs_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m rtfs_net_b200.build [--force]

nvcc cross-compiles without a GPU; the .so travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librtfs_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    out = [os.path.join(INCLUDE, "rtfs_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def up_to_date():
    if not os.path.exists(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(s) <= t for s in sources())


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def refresh_abi():
    """Freeze the header's enum tables into rtfs_net_b200/_abi.py (so the package does not need include/ at run time)."""
    import importlib

    from . import _abi, _lib

    if os.path.exists(_lib.HEADER):
        snap = lambda: {k: v for k, v in vars(_abi).items() if k.isupper()}
        before = snap()
        _lib.write_abi()
        importlib.reload(_abi)
        after = snap()
        if before != after:
            importlib.reload(_lib)


def build(force=False, verbose=False):
    """Compile the library if any source is newer than it.  Returns the library path."""
    refresh_abi()
    if not force and up_to_date():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # GPU box without a toolkit: use the prebuilt library that travelled
        raise RuntimeError("nvcc not found and no prebuilt librtfs_b200.so present")
    os.makedirs(LIB_DIR, exist_ok=True)
    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    tmp = LIB_PATH + ".tmp"
    extra = os.environ.get("RTFS_NVCC_EXTRA", "").split()  # e.g. -DRTFS_TCP_TRACE (epilogue phase stamps), -DRTFS_MBAR_HINT_NS=0
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + units
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
