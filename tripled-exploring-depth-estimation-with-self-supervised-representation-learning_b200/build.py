"""Builds libtdl.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m <package>.build        or        __graft_entry__.build()

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to
the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtdl.so")
SOURCES = ["tdl_api.cu", "tdl_photo.cu", "tdl_photo2.cu", "tdl_smooth.cu", "tdl_feat.cu", "tdl_feat2.cu", "tdl_pose.cu", "tdl_input.cu"]
HEADERS = [os.path.join(CSRC, "tdl_common.cuh"), os.path.join(CSRC, "tdl_ssim.cuh"), os.path.join(CSRC, "tdl_internal.h"), os.path.join(CSRC, "tdl_tma.cuh"),
           os.path.join(ROOT, "include", "tdl.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libtdl.so cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    logs = []
    with ThreadPoolExecutor(max_workers=4) as ex:
        for s, r in ex.map(compile_one, jobs):
            logs.append(r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}:\n{r.stdout}\n{r.stderr}")
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or force or _stale(OUT, objs):
        r = subprocess.run([nvcc, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("\n".join(logs))
    with open(os.path.join(objdir, "ptxas.log"), "a" if not force else "w") as fh:
        fh.write("\n".join(logs))
    return OUT


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
