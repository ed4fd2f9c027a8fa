"""Builds iskra_b200/libiskra_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m iskra_b200.build [--force]

-fmad=false: the parity-critical arithmetic must not be contracted into FMAs;
kernels that want FMAs (GEMM, FFT) call fma() explicitly.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libiskra_b200.so")
SOURCES = ["api.cu", "particles.cu", "advance_fused.cu", "advance_tile.cu", "sort.cu", "poisson.cu", "mcc.cu", "see.cu", "comm.cu", "surfaces.cu", "dsmc.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "iskra_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    env = dict(os.environ)
    env.pop("CC", None)   # the image's $CC wrapper is not a usable nvcc host compiler
    env.pop("CXX", None)
    r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed, see %s" % log)
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", SO)
