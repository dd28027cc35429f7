"""In-tree build of the CUDA library (sm_100a) -- used by ``__graft_entry__.build()``.

    python -m protocols.secure_comparison_b200.build [--force] [-v]

nvcc cross-compiles without a GPU.  Every ``csrc/*.cu`` is compiled to an object file in parallel (one nvcc per
translation unit), then linked into ``libscb200.so`` next to this file so that it travels with the repository snapshot to
the GPU box (it is git-ignored, not gpurun-ignored).  The library carries a BUILD ID = SHA-256 over its sources, the C-ABI
header and the compiler flags (``scb_build_id()``); ``build()`` recompiles whenever the id of the tree differs from the id
inside the existing ``.so`` -- content, not mtimes, decides."""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "scb200.h")
SRC = [os.path.join(CSRC, f) for f in ("scb200.cu", "scb200_protocol.cu", "scb200_rng.cu", "scb200_peaks.cu")]
LIB = os.path.join(HERE, "libscb200.so")
OBJ_DIR = os.path.join(HERE, "_build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def source_id() -> str:
    """SHA-256 over every file of csrc/, the header and the flags (sorted by name, so it is reproducible)."""
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp"))) + [HEADER]
    for path in files:
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:32]


def built_id() -> str:
    """Build id embedded in the existing library ("" when there is none or it predates build ids).  No GPU needed."""
    if not os.path.exists(LIB):
        return ""
    try:
        lib = ctypes.CDLL(LIB)
        fn = lib.scb_build_id
        fn.restype = ctypes.c_char_p
        return fn().decode()
    except (OSError, AttributeError):
        return ""


def needs_build() -> bool:
    return built_id() != source_id()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    bid = source_id()

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + [f'-DSCB_BUILD_ID="{bid}"'] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return obj, cmd, proc

    with ThreadPoolExecutor(max_workers=len(SRC)) as pool:
        results = list(pool.map(compile_one, SRC))
    for obj, cmd, proc in results:
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if verbose:
            print(proc.stderr)
    tmp = LIB + ".tmp"
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + [obj for obj, _, _ in results]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
