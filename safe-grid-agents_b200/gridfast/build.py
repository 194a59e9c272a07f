"""Build the CUDA library in-tree (sm_100a only).

    python -m gridfast.build        (or __graft_entry__.build())

nvcc cross-compiles without a GPU.  The resulting libsgk.so sits next to this
file: git-ignored, but it travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB = os.path.join(HERE, "libsgk.so")
SOURCES = ["sgk.cu", "sgk_dqn.cu"]
HEADERS = ["sgk_common.cuh", "sgk_envs.cuh", "sgk_table.cuh", "sgk_internal.cuh", "sgk_mlp_tc.cuh", "../../include/sgk.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-shared",
    "--threads", "2",           # the two translation units compile side by side
]
# SGK_FAST_BUILD=1: optimise each translation unit's kernels on all cores (3m40 -> 1m20).  Development
# only: measured -2.8 % on the headline rollout kernel (4.61 -> 4.74 ms per launch, same box A/B,
# DESIGN.md section 4), so release builds -- build() below, what the driver runs -- do without.
FAST_FLAGS = ["--split-compile", "0"]


OBJ_DIR = os.path.join(HERE, "_build")
# what each translation unit includes (sgk.cu does not see the tensor-core header)
DEPS = {"sgk.cu": [h for h in HEADERS if h != "sgk_mlp_tc.cuh"], "sgk_dqn.cu": HEADERS}


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    built = os.path.getmtime(target)
    return any(os.path.getmtime(s) > built for s in sources)


def stale():
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return _newer(LIB, deps)


def build(force=False, verbose=False):
    """One object per translation unit (compiled side by side, recompiled only when its own sources
    changed), then one link.  No relocatable device code: the units share host symbols only."""
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    fast = FAST_FLAGS if os.environ.get("SGK_FAST_BUILD") else []
    tag = "fast" if fast else "release"
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ("-shared", "--threads", "2")]
    jobs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, "%s.%s.o" % (os.path.splitext(src)[0], tag))
        objs.append(obj)
        if force or _newer(obj, [os.path.join(CSRC, f) for f in [src] + DEPS[src]]):
            cmd = [nvcc] + flags + fast + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
            jobs.append((cmd, subprocess.Popen(cmd)))
    for cmd, job in jobs:
        if job.wait() != 0:
            raise subprocess.CalledProcessError(job.returncode, cmd)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
