"""
Builds libmoonb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m moonrtx_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmoonb200.so")
OBJ = os.path.join(HERE, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=true",      # bit-exact paths use __f*_rn intrinsics, which are never contracted
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            extra = os.environ.get("MRTX_NVCC_EXTRA", "").split()      # development only (tuning sweeps)
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose and out:
            print(out)
    if failed:
        raise RuntimeError("libmoonb200 build failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs + ["-ldl", "-Xcompiler", "-fPIC"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
