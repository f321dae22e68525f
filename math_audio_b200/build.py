"""Build libbemb200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m math_audio_b200.build [--force]

What build.rs of the `bem-b200-sys` crate does on the Rust side;
kept in Python here because that is the host toolchain of this image.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIBDIR = HERE / "lib"
LIB = LIBDIR / "libbemb200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off"]

# translation unit -> extra flags
UNITS = {
    # decision-taking kernels: IEEE ops in the reference's order, no FMA contraction
    "assembly_exact.cu": ["-fmad=false"],
    "assembly_far.cu": [],
    "linalg.cu": [],
    "gmres.cu": [],
    "gmres_fused.cu": [],
    "block_gmres.cu": [],
    "schwarz.cu": [],
    "block_matvec.cu": [],
    "postprocess.cu": ["-fmad=false"],
    "room.cu": [],
    "direct.cu": [],
    "sweep.cu": [],
    "api.cu": [],
}


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: libbemb200 cannot be built (there is no CPU fallback)")
    return cand


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "bemb200.h"]
    objs = []
    rebuilt = False
    for unit, extra in UNITS.items():
        src = CSRC / unit
        obj = objdir / (unit + ".o")
        objs.append(str(obj))
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
            rebuilt = True
    if rebuilt or force or not LIB.exists():
        cmd = [nvcc, *ARCH, "-shared", "-o", str(LIB), *objs, "-Xcompiler", "-fPIC", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
