"""Builds libugp_b200.so in-tree with nvcc for sm_100a.

The shared library is the product's only native artefact; it is git-ignored
but travels to the GPU box with the repo snapshot.  `python -m
ug_pg_unet_b200.build` (or `__graft_entry__.build()`) rebuilds it when any
source is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libugp_b200.so"
SOURCES = ["tmap.cu", "conv_gemm.cu", "conv_halo.cu", "conv_pair.cu", "conv_wgrad.cu", "conv_wgrad_halo.cu", "conv_transpose.cu", "elementwise.cu", "bn_bwd.cu", "first_layer.cu", "conv_first.cu",
           "weights.cu", "head_loss.cu", "mlp_head.cu", "eval_metrics.cu", "optim.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libugp_b200.so cannot be built")


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    lib_m = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "ugp_common.cuh", CSRC / "ugp_vec.cuh", CSRC / "conv_gemm.cuh", CSRC / "conv_epilogue.cuh", CSRC / "first_rows.cuh",
                                         PKG_DIR.parent / "include" / "ugp_b200.h"]
    return any(d.stat().st_mtime > lib_m for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


MICROBENCH_SRC = PKG_DIR.parent / "tools" / "microbench" / "bench_mma.cu"
MICROBENCH_LIB = MICROBENCH_SRC.parent / "libugp_microbench.so"


def build_microbench(force: bool = False) -> Path:
    """tools/microbench/libugp_microbench.so: the tcgen05 issue-cost microbenchmark (tools/t_mma*.py).
    A diagnostic library of its own - nothing of it is in the product library or its header."""
    if not force and MICROBENCH_LIB.exists() and MICROBENCH_LIB.stat().st_mtime >= max(
            MICROBENCH_SRC.stat().st_mtime, (CSRC / "ugp_common.cuh").stat().st_mtime):
        return MICROBENCH_LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", str(MICROBENCH_SRC), "-o", str(MICROBENCH_LIB), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {MICROBENCH_SRC}:\n{r.stdout}\n{r.stderr}")
    return MICROBENCH_LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
    print(build_microbench(force="--force" in sys.argv))
