"""Build the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).

The exact kernel is instantiated once per decoder in its own translation unit so the objects
compile in parallel; everything is linked into stringsext_b200/libstringsext_b200.so.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libstringsext_b200.so")
DEPS = ["sx_mb_tables.inc", "sx_core.cuh", "sx_fast_utf8.cuh", "sx_fast_generic.cuh", "sx_mask_utf8.cuh", "sx_sparse_utf8.cuh", "sx_exact.cuh", "sx_scan.cu", "sx_exact_inst.cu", os.path.join("..", "..", "include", "stringsext_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
N_INST = 9


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or is_stale()):
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    jobs = [([nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, "sx_scan.cu"), "-o", os.path.join(OBJ, "sx_scan.o")])]
    for i in range(N_INST):
        jobs.append([nvcc, *NVCC_FLAGS, *extra, f"-DSX_INST={i}", "-c", os.path.join(CSRC, "sx_exact_inst.cu"), "-o",
                     os.path.join(OBJ, f"sx_exact_{i}.o")])

    def run(cmd):
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        print("\n".join(logs))
    objs = [os.path.join(OBJ, "sx_scan.o")] + [os.path.join(OBJ, f"sx_exact_{i}.o") for i in range(N_INST)]
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB, *objs], cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
