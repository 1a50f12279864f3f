"""DEVELOPMENT TOOL: the beam pre-pass (trace_fast.cuh, BeamCtl) on the CPU - every jittered sample of a pixel must be
decided exactly as without it (same status, same patch, same ray parameter), and the node counts say what it saves."""
import ctypes as C, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from debug_parity import build
from debug_fast import run_fast
from oracle import downscale_oracle as dorc
from moonrtx_b200.synth import synth_ldem

SCALE = float(np.float32(0.5 / 1737400.0))


def pixel_rays(iw, ih, eye, target, up, fov, jit, u=(0, 0, 1), v=(0, -1, 0)):
    """rays of every pixel at sub-pixel positions jit[(n, 2)] -> (ih*iw, n, 6) in the body frame"""
    eye = np.array(eye, float); w = np.array(target, float) - eye; w /= np.linalg.norm(w)
    right = np.cross(w, np.array(up, float)); right /= np.linalg.norm(right); up2 = np.cross(right, w)
    t = math.tan(math.radians(fov) / 2); asp = iw / ih
    ys, xs = np.meshgrid(np.arange(ih), np.arange(iw), indexing="ij")
    sx = ((xs[..., None] + jit[None, None, :, 0]) / iw * 2 - 1) * t * asp
    sy = (1 - (ys[..., None] + jit[None, None, :, 1]) / ih * 2) * t
    d = w + sx[..., None] * right + sy[..., None] * up2
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    ez = np.array(u, float); ez /= np.linalg.norm(ez); vv = np.array(v, float); vv -= vv.dot(ez) * ez; vv /= np.linalg.norm(vv)
    ex = np.cross(ez, vv); Rm = np.stack([ex, -vv, ez])
    ob = Rm @ eye; db = d @ Rm.T
    rays = np.concatenate([np.broadcast_to(ob, db.shape), db], axis=-1)
    return np.ascontiguousarray(rays.reshape(ih * iw, -1, 6)), t / ih * math.sqrt(2.0)


def dm(counts, rs):
    f = lambda m: float(np.float32(np.float32(np.float32(np.float32(m) * np.float32(SCALE)) + np.float32(1)) / np.float32(rs)))
    return f(counts.max()), f(counts.min())


def run_case(l, counts, rs, name, iw, ih, eye, target, fov, nj=8, drop=1, u=(0, 0, 1), v=(0, -1, 0), seed=1):
    H, W = counts.shape
    rng = np.random.default_rng(seed)
    jit = np.concatenate([[[0.5, 0.5]], rng.uniform(0, 1, (nj - 5, 2)), [[0, 0], [1 - 1e-9, 0], [0, 1 - 1e-9], [1 - 1e-9, 1 - 1e-9]]])
    rays, delta = pixel_rays(iw, ih, eye, target, (0, 0, 1), fov, jit, u, v)
    npx, n = rays.shape[:2]
    dmax, dmin = dm(counts, rs)
    centre = np.ascontiguousarray(rays[:, 0])
    bo = np.zeros((npx, 4))
    l.dbg_beam.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                           C.c_double, C.c_double, C.c_void_p]
    l.dbg_beam(counts.ctypes.data, 1, W, H, SCALE, rs, dmax, dmin, centre.ctypes.data, npx, delta, 10.0, bo.ctypes.data)
    flat = np.ascontiguousarray(rays.reshape(-1, 6))
    base = run_fast(l, counts, flat, scale=SCALE, rs=rs, start_level=-3)
    alive = np.repeat(bo[:, 0] > 0, n)
    s_min = np.ascontiguousarray(np.repeat(np.where(bo[:, 0] > 0, bo[:, 1], 1e30), n))
    lvl = np.ascontiguousarray(np.repeat(np.maximum(bo[:, 2] - drop, 0), n).astype(np.int32))
    out = np.zeros((len(flat), 8))
    l.dbg_trace_fast_from.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    l.dbg_trace_fast_from(counts.ctypes.data, 1, W, H, SCALE, rs, dmax, flat.ctypes.data, len(flat), s_min.ctypes.data, lvl.ctypes.data, 10.0, out.ctypes.data)
    sb, so = base[:, 0].astype(int) & 3, out[:, 0].astype(int) & 3
    texel = 2 * math.pi * 10 / W
    decided = (sb != 2) & (so != 2)
    bad_status = decided & (sb != so)
    both = (sb == 1) & (so == 1)
    ds = np.abs(base[:, 1] - out[:, 1]) / texel
    bad_hit = both & ((base[:, 4] != out[:, 4]) | (base[:, 5] != out[:, 5]) | (ds > 1e-4))
    inside = base[:, 6] > 0
    print(f"{name}: px {npx} x {n}  hits {int((sb == 1).sum())}  beam-miss px {int((bo[:, 0] == 0).sum())}  no-info px {int(((bo[:, 0] > 0) & (bo[:, 1] == 0)).sum())}  "
          f"WRONG status {int(bad_status.sum())}  WRONG hit {int(bad_hit.sum())}  defer {int((sb == 2).sum())}->{int((so == 2).sum())}  "
          f"nodes/ray {base[inside, 6].mean():.2f} -> {out[inside, 6].mean():.2f} (+ beam {bo[:, 3].sum() / max(1, inside.sum()):.2f})  "
          f"tests/ray {base[inside, 7].mean():.2f} -> {out[inside, 7].mean():.2f}  levels {np.bincount(bo[bo[:, 0] > 0, 2].astype(int))}")
    for i in np.nonzero(bad_status | bad_hit)[0][:6]:
        print("   wrong", i, divmod(i // n, iw), "sample", i % n, "base", base[i, :6], "beam", out[i, :6], "s_b", s_min[i], "lvl", lvl[i], "ds", ds[i])
    return int(bad_status.sum() + bad_hit.sum())


if __name__ == "__main__":
    l = build()
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    drop = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t0 = time.time()
    if which == "small":
        counts = synth_ldem(1440, 720, seed=11, craters=60)
    elif which == "mid":
        counts = np.load(os.path.join(ROOT, "gpurun_out", "synth_5760x2880.npy"))
    else:
        counts = synth_ldem(11520, 5760, seed=7, craters=200)
    H, W = counts.shape
    _, rs = dorc.load_elevation(counts, 1)
    px = W / (2 * math.pi) * 2 * math.tan(math.radians(4.242192793) / 2) * 300 / 10   # texels across the frame height at fov 4.24
    bad = 0
    # whole disk, 15 / 4 / 1 texels per pixel
    for tp in (15, 4, 1):
        ih = max(int(px / tp), 8); iw = ih * 16 // 9
        if iw * ih > 400000:
            continue
        bad += run_case(l, counts, rs, f"disk {tp} texel/px", iw, ih, (0, -300, 0), (0, 0, 0), 4.242192793, drop=drop)
    # tilted: the north pole in view, close-ups of pole, limb and terminator-like grazing views
    bad += run_case(l, counts, rs, "pole tilt", 200, 150, (0, -300, 0), (0, 0, 0), 4.242192793, drop=drop, u=(0, -0.8, 0.6), v=(0, -0.6, -0.8))
    bad += run_case(l, counts, rs, "pole zoom", 160, 120, (0, -300, 0), (0, 0, 0), 0.4, drop=drop, u=(0, -1, 0.02), v=(0, -0.02, -1))
    bad += run_case(l, counts, rs, "limb zoom", 160, 120, (0, -300, 0), (9.9, 0, 1.0), 0.3, drop=drop)
    bad += run_case(l, counts, rs, "south limb", 160, 120, (0, -300, 0), (0.5, 0, -9.95), 0.5, drop=drop)
    bad += run_case(l, counts, rs, "oblique eye", 200, 150, (200, -200, 100), (0, 0, 0), 4.5, drop=drop)
    print("TOTAL WRONG", bad, f"({time.time() - t0:.0f} s)")
