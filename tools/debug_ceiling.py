"""DEVELOPMENT TOOL: the ceiling test of rising (shadow) rays on the CPU - same decisions as the plain walk, fewer nodes."""
import ctypes as C, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from debug_parity import build, camera_rays
from debug_fast import run_fast, shadow_rays
from debug_beam import SCALE, dm
from oracle import downscale_oracle as dorc
from helpers import sun_at_phase


def trace_from(l, counts, rs, rays, level, ceil_level):
    H, W = counts.shape
    dmax, dmin = dm(counts, rs)
    l.dbg_set_ceiling.argtypes = [C.c_int, C.c_float]
    l.dbg_set_ceiling(ceil_level, dmin)
    l.dbg_trace_fast_from.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    out = np.zeros((len(rays), 8))
    s_min = np.zeros(len(rays)); lvl = np.full(len(rays), level, np.int32)
    l.dbg_trace_fast_from(counts.ctypes.data, 1, W, H, SCALE, rs, dmax, rays.ctypes.data, len(rays), s_min.ctypes.data, lvl.ctypes.data, 10.0, out.ctypes.data)
    l.dbg_set_ceiling(0, dmin)
    return out


if __name__ == "__main__":
    l = build()
    path = sys.argv[1]
    counts = np.load(path, mmap_mode="r")
    counts = np.ascontiguousarray(counts)
    H, W = counts.shape
    rs = float(np.float32(np.float32(np.float32(counts.max()) * np.float32(SCALE)) + np.float32(1)))
    iw, ih = (int(v) for v in (sys.argv[2], sys.argv[3])) if len(sys.argv) > 3 else (480, 270)
    rays, _ = camera_rays(iw, ih, (0, -300, 0), (0, 0, 0), (0, 0, 1), 4.242192793)
    t0 = time.time()
    prim = run_fast(l, counts, rays, scale=SCALE, rs=rs, start_level=-3)
    print(f"primary: nodes/ray {prim[prim[:, 6] > 0, 6].mean():.2f} ({time.time() - t0:.0f} s incl. pyramid)")
    for ph in (90.0, 45.0):
        sr = shadow_rays(rays, np.stack([prim[:, 0].astype(int) & 3 == 1, prim[:, 1]], axis=1).astype(float), sun_at_phase(ph))
        base = trace_from(l, counts, rs, sr, 2, 0)
        for cl in (2, 3, 4, 5):
            out = trace_from(l, counts, rs, sr, 2, cl)
            sb, so = base[:, 0].astype(int) & 3, out[:, 0].astype(int) & 3
            wrong = ((sb == 1) != (so == 1)) & (sb != 2) & (so != 2)
            print(f"phase {ph}: shadow rays {len(sr)}  occluded {int((sb == 1).sum())}  ceil@{cl}: WRONG {int(wrong.sum())}  defer {int((sb == 2).sum())}->{int((so == 2).sum())}  "
                  f"nodes/ray {base[:, 6].mean():.2f} -> {out[:, 6].mean():.2f}  (unoccluded only: {base[sb == 0, 6].mean():.2f} -> {out[sb == 0, 6].mean():.2f})")
