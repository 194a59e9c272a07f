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
# profiles/r02_notes.md), so release builds -- build() below, what the driver runs -- do without.
FAST_FLAGS = ["--split-compile", "0"]


def stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    fast = FAST_FLAGS if os.environ.get("SGK_FAST_BUILD") else []
    cmd = [nvcc] + NVCC_FLAGS + fast + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
