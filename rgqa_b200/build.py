"""Build the CUDA extension in-tree with nvcc for sm_100a (no torch involvement, plain C ABI).

Output: ``rgqa_b200/_C/librgqa_b200.so``.  The library is git-ignored but travels to the GPU box
with the repo snapshot.  ``python -m rgqa_b200.build`` rebuilds; ``build(force=False)`` is a no-op
when the library is newer than every source.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_C"
LIB = OUT_DIR / "librgqa_b200.so"
SOURCES = ["api.cu", "prof.cu", "gemm_tcgen05.cu", "gemm_ln_tcgen05.cu", "visn_embed_tcgen05.cu", "qkv_attn_tcgen05.cu", "gemm_simt.cu", "attention.cu", "rowops.cu", "scorers.cu", "estimator.cu", "optim.cu", "backward.cu", "attention_bwd_mma.cu"]
HEADERS = ["common.cuh", "kernels.h", "../../include/rgqa_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; the rgqa_b200 CUDA extension cannot be built")
    return cand


def have_nvcc() -> bool:
    return bool(shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"))


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not stale():
        return LIB
    OUT_DIR.mkdir(exist_ok=True)
    obj_dir = OUT_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for s in SOURCES:
        o = obj_dir / (Path(s).stem + ".o")
        objs.append(str(o))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out:
            print(out)
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", *objs, "-o", str(LIB)]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
