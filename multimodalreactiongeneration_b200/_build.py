"""Builds csrc/*.cu into the in-tree C-ABI shared library (sm_100a only).

    python -m multimodalreactiongeneration_b200._build [--force]

nvcc cross-compiles without a GPU; the resulting ``csrc/libmrg_b200.so`` is git-ignored but travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmrg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(
        os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("MRG_EXTRA_NVCC_FLAGS", "").split()  # e.g. -DMRG_REC_TRACE (developer builds)
    objs = []
    procs = []
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    live = {os.path.basename(s)[:-3] + ".o" for s in _sources()}
    for stale in glob.glob(os.path.join(CSRC, "build", "*.o")):   # objects of deleted sources must not be linked
        if os.path.basename(stale) not in live:
            os.remove(stale)
    for src in _sources():
        obj = os.path.join(CSRC, "build", os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *flags, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed; see stderr")
    # -cudart shared: torch already ships libcudart, so the library does not embed a second copy of the runtime
    cmd = [NVCC, "-shared", "-cudart", "shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
