"""In-tree build of libnam_b200.so with nvcc for sm_100a (no JIT cache, no torch extension machinery).

The shared library is written to neuralampmodelercore_b200/lib/ so it travels with the repo snapshot
to the GPU box; it is git-ignored (*.so) so history stays source-only.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libnam_b200.so"
INCLUDE = PKG.parent / "include"

SOURCES = ["nam_b200.cu", "nam_model_spec.cpp", "json_lite.cpp", "wavenet_pack.cpp", "generic_pack.cpp", "nam_dsp_shim.cpp"]

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libnam_b200.so cannot be built")


def sources() -> list[Path]:
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*")) + list(INCLUDE.rglob("*.h"))
    return any(p.stat().st_mtime > t for p in deps if p.is_file())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA/C++ source of the product into lib/libnam_b200.so (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    tmp = LIB_DIR / "libnam_b200.so.tmp"
    cmd = [_nvcc(), *NVCC_FLAGS, f"-I{INCLUDE}", "-o", str(tmp), *[str(s) for s in sources()]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(proc.stderr)
    return LIB_PATH
