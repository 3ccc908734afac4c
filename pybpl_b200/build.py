"""Builds pybpl_b200/libpybpl_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m pybpl_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the working tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libpybpl_b200.so")
SOURCES = ["bpl_render.cu", "bpl_splines.cu", "bpl_token_prior.cu", "bpl_token_sample.cu", "bpl_type_sample.cu", "bpl_type_prior.cu"]
HEADERS = [os.path.join(CSRC, "bpl_common.cuh"), os.path.join(CSRC, "bpl_forward.cuh"), os.path.join(CSRC, "bpl_forward1s.cuh"), os.path.join(CSRC, "bpl_backward_teams.cuh"), os.path.join(CSRC, "bpl_anysize.cuh"), os.path.join(CSRC, "bpl_points.cuh"), os.path.join(CSRC, "bpl_token_common.cuh"), os.path.join(os.path.dirname(HERE), "include", "pybpl_b200.h")]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", SO]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpybpl_b200.so")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
