"""Build libb200nerf.so (the C-ABI library) in-tree with nvcc for sm_100a only.

    python -m nerfstudio_b200.build          # or: from nerfstudio_b200.build import build; build()

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "libb200nerf.so")
STAMP = os.path.join(PKG, "csrc", ".build_stamp")

SOURCES = ["runtime.cu", "hashgrid.cu", "mlp.cu", "encodings.cu", "sampling.cu", "render.cu", "raygen.cu",
           "packed.cu", "adam.cu", "fuse.cu", "tc_test.cu", "mlp_tc.cu", "density_fused.cu", "splat.cu", "wide_mlp.cu", "step_glue.cu", "ray_tail.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v" if os.environ.get("B2N_PTXAS_V") else "-O3"] + os.environ.get("B2N_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 core cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(INCLUDE, "b200nerf.h"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu into one shared library; returns its path.  Rebuilds only when sources changed."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, "csrc", "_obj"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(PKG, "csrc", "_obj", src.replace(".cu", ".o"))
        cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj, "-I", INCLUDE, "-I", CSRC] + NVCC_FLAGS
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[b200nerf build] {src} FAILED\n{' '.join(cmd)}\n{out}\n")
        elif verbose and out.strip():
            sys.stderr.write(f"[b200nerf build] {src}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed building libb200nerf.so")
    link = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(link, check=True)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
