"""Build the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""

from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libchimeralm_b200.so"
SOURCES = [CSRC / "api.cu", CSRC / "bam_ingest.cpp"]
HEADERS = sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "chimeralm_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB), *map(str, SOURCES), "-lz"]
    if os.environ.get("CLM_EXPERIMENTS") == "1":
        # recorded-slower variants (block_mlp2 / block_mlp16 / block_mlp_pp, the first long-convolution kernel): not part of
        # the product library; `clm_set_option` refuses their switches unless they were compiled in
        cmd.insert(1, "-DCLM_EXPERIMENTS")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({r.returncode}):\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
