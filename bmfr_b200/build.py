"""In-tree native builds: the CUDA library (sm_100a) and, for the test side, the CPU oracle.

Nothing here needs a GPU: nvcc cross-compiles sm_100a on the CPU box and the resulting
``bmfr_b200/libbmfr_b200.so`` travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "bmfr_b200" / "csrc"
LIB = ROOT / "bmfr_b200" / "libbmfr_b200.so"
BUILD_DIR = ROOT / "bmfr_b200" / "_build"

CUDA_SOURCES = ["bmfr_kernels.cu", "bmfr_reforder.cu", "bmfr_fit.cu", "bmfr_post.cu", "bmfr_pipeline.cu", "synth.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-O2",
    "-Xptxas", "-v",
]
# Contraction only where fmaf() is written (csrc/bmfr_device.cuh): everything that is compared bit for
# bit with the oracle.  bmfr_post.cu is tolerance-only arithmetic and is compiled with contraction.
PER_SOURCE_FLAGS = {"bmfr_post.cu": []}
DEFAULT_SOURCE_FLAGS = ["--fmad=false"]


def _src_flags(src: str):
    return PER_SOURCE_FLAGS.get(src, DEFAULT_SOURCE_FLAGS)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _digest(paths, extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(str(p.name).encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), out: Path | None = None) -> Path:
    """Compile csrc/*.cu into bmfr_b200/libbmfr_b200.so (skipped when sources are unchanged).
    `out` + `extra_flags` build a tuning variant next to it (see scripts/)."""
    if out is not None:
        return _build_variant(out, list(extra_flags), verbose)
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "bmfr_b200.h"]
    stamp = BUILD_DIR / "lib.sha256"
    digest = _digest(deps, " ".join(NVCC_FLAGS + list(extra_flags)) + repr(PER_SOURCE_FLAGS))
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    if not (CSRC / CUDA_SOURCES[0]).exists():
        raise RuntimeError(f"sources missing under {CSRC}")
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    for src in CUDA_SOURCES:
        obj = BUILD_DIR / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *_src_flags(src), *extra_flags, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD_DIR / (src + ".log")).write_text(r.stdout + r.stderr)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(str(obj))
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-Xcompiler", "-fopenmp", "-lgomp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libbmfr_b200.so failed")
    stamp.write_text(digest)
    return LIB


IO_LIB = ROOT / "bmfr_b200" / "libbmfr_io.so"


def build_io(force: bool = False) -> Path:
    """libbmfr_io.so: EXR reader, camera_matrices.h parser, PNG writer (include/bmfr_io.h); host only, no CUDA."""
    src, hdr = CSRC / "bmfr_io.cpp", ROOT / "include" / "bmfr_io.h"
    if not force and IO_LIB.exists() and IO_LIB.stat().st_mtime > max(src.stat().st_mtime, hdr.stat().st_mtime):
        return IO_LIB
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-shared", "-fPIC", str(src), "-o", str(IO_LIB), "-lz"], check=True)
    return IO_LIB


def build_driver(force: bool = False) -> Path:
    """bmfr_run: the reference's driver program (tasks(), bmfr.cpp:179-556) on top of the C ABI."""
    exe, srcs = ROOT / "bmfr_b200" / "bmfr_run", [CSRC / "bmfr_main.cpp", CSRC / "bmfr_io.cpp"]
    build_library()
    newest = max([s.stat().st_mtime for s in srcs] + [LIB.stat().st_mtime, (ROOT / "include" / "bmfr_io.h").stat().st_mtime])
    if not force and exe.exists() and exe.stat().st_mtime > newest:
        return exe
    subprocess.run(["g++", "-std=c++17", "-O2", "-fopenmp", *map(str, srcs), "-o", str(exe), f"-L{LIB.parent}", "-lbmfr_b200", "-lz",
                    f"-Wl,-rpath,{LIB.parent}", "-Wl,-rpath,$ORIGIN"], check=True)
    return exe


def _build_variant(out: Path, extra_flags, verbose):
    nvcc = _nvcc()
    vdir = BUILD_DIR / out.stem
    vdir.mkdir(parents=True, exist_ok=True)
    objs = []
    for src in CUDA_SOURCES:
        obj = vdir / (src + ".o")
        r = subprocess.run([nvcc, *NVCC_FLAGS, *_src_flags(src), *extra_flags, "-c", str(CSRC / src), "-o", str(obj)], capture_output=True, text=True)
        (vdir / (src + ".log")).write_text(r.stdout + r.stderr)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(str(obj))
    subprocess.run([nvcc, "-shared", "-o", str(out), *objs, "-Xcompiler", "-fopenmp", "-lgomp"], check=True)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_io(force="--force" in sys.argv))
    print(build_driver(force="--force" in sys.argv))
