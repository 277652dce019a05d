"""Builds libpdg_b200.so (the CUDA kernels + C-ABI) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpdg_b200.so")
SOURCES = ["pdg_b200.cu"]
HEADERS = ["pdg_device.cuh", "pdg_kernels.cuh", "pdg_run.cuh", "pdg_randomize.cuh", "pdg_loglin.cuh",
           "pdg_tally.cuh", os.path.join("..", "..", "include", "pdg_b200.h")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        fp = os.path.join(CSRC, f)
        if os.path.exists(fp) and os.path.getmtime(fp) > t:
            return True
    return False


def build_native(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if os.environ.get("PDG_LANE_SMEM"):
        cmd.insert(1, "-DPDG_LANE_SMEM")
    if os.environ.get("PDG_NO_K4"):
        cmd.insert(1, "-DPDG_NO_K4=1")
    if os.environ.get("PDG_PHASE_TIMERS"):
        cmd.insert(1, "-DPDG_PHASE_TIMERS")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build_native(force="--force" in sys.argv, verbose=True)
    print(LIB)
