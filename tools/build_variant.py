"""Development helper: build moonrtx_b200/_variants/libmoonb200_<tag>.so with extra nvcc flags on trace.cu (tuning sweeps).

    python tools/build_variant.py <tag> [-DNAME=VALUE ...]

tools/bench_trace.py loads it when MRTX_LIB points at it.  Never used by the package itself.
"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moonrtx_b200 import build as b
tag, extra = sys.argv[1], sys.argv[2:]
b.build()
vdir = os.path.join(b.HERE, "_variants")
os.makedirs(vdir, exist_ok=True)
obj = os.path.join(vdir, f"trace_{tag}.o")
subprocess.run(["nvcc"] + b.NVCC_FLAGS + extra + ["-c", os.path.join(b.CSRC, "trace.cu"), "-o", obj], check=True)
objs = [o for o in glob.glob(os.path.join(b.OBJ, "*.o")) if os.path.basename(o) != "trace.o"] + [obj]
lib = os.path.join(vdir, f"libmoonb200_{tag}.so")
subprocess.run(["nvcc", "-shared", "-o", lib] + objs + ["-ldl", "-Xcompiler", "-fPIC"], check=True)
os.remove(obj)
print(lib)
