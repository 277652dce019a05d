"""Builds libpdg_b200.so (the CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

The C-ABI translation unit and one translation unit per likelihood model (the template
instantiations of the MH kernel) are compiled in parallel, then linked into one shared library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libpdg_b200.so")
SOURCES = ["pdg_b200.cu", "pdg_model_ggm.cu", "pdg_model_loglin.cu", "pdg_model_uniform.cu", "pdg_probe.cu"]
HEADERS = ["pdg_device.cuh", "pdg_kernels.cuh", "pdg_run.cuh", "pdg_randomize.cuh", "pdg_loglin.cuh",
           "pdg_tally.cuh", "pdg_ctx.h", "pdg_model.inl", "pdg_ggm_big.cuh", "pdg_post.cuh",
           os.path.join("..", "..", "include", "pdg_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        fp = os.path.join(CSRC, f)
        if os.path.exists(fp) and os.path.getmtime(fp) > t:
            return True
    return False


def _flags():
    fl = []
    for name in ("PDG_LANE_SMEM", "PDG_PHASE_TIMERS"):
        if os.environ.get(name):
            fl.append("-D" + name)
    if os.environ.get("PDG_NO_K4"):
        fl.append("-DPDG_NO_K4=1")
    if os.environ.get("PDG_EXTRA_NVCC"):
        fl += os.environ["PDG_EXTRA_NVCC"].split()
    return fl


def build_native(force=False, verbose=False, variant=None, extra=()):
    """variant: developer build with extra nvcc flags into libpdg_b200_<variant>.so (select it with PDG_LIB)"""
    if variant is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    obj_dir = OBJ if variant is None else OBJ + "_" + variant
    lib = LIB if variant is None else os.path.join(HERE, "libpdg_b200_%s.so" % variant)
    os.makedirs(obj_dir, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(s):
        o = os.path.join(obj_dir, s[:-3] + ".o")
        cmd = [nvcc] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"] + _flags() + list(extra)
        if verbose:
            cmd.append("-Xptxas=-v")
        cmd += ["-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", lib] + objs)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_native(force=True, verbose="-v" in sys.argv, variant=sys.argv[i + 1], extra=sys.argv[i + 2:]))
    else:
        print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
