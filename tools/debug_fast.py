"""DEVELOPMENT TOOL: the filtered float32 path (trace_fast.cuh) against the exact float64 path, both on the CPU."""
import ctypes as C, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from debug_parity import build, camera_rays, run_host
from oracle import downscale_oracle as dorc
from moonrtx_b200.synth import synth_ldem
from helpers import sun_at_phase, DEFAULTS


def run_fast(l, elev, rays, s_min=0.0, scale=0.0, rs=1.0, start_level=-3):
    out = np.zeros((len(rays), 8))
    if elev.dtype == np.int16:
        m = np.float32(elev.max()); dmax = float(np.float32(np.float32(np.float32(m*np.float32(scale))+np.float32(1))/np.float32(rs)))
    else:
        dmax = float(elev.max())
    l.dbg_trace_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                 C.c_double, C.c_double, C.c_int, C.c_void_p]
    l.dbg_trace_fast(elev.ctypes.data, int(elev.dtype == np.int16), elev.shape[1], elev.shape[0], scale, rs, dmax,
                     rays.ctypes.data, len(rays), s_min, 10.0, start_level, out.ctypes.data)
    return out


def report(name, W, ex, fa):
    texel = 2 * math.pi * 10 / W
    raw = fa[:, 0].astype(int)
    st = raw & 3
    reasons = np.bincount((raw >> 2)[st == 2], minlength=1)
    hit_e = ex[:, 0] > 0
    n = len(st)
    dec = st != 2
    wrong = dec & ((st == 1) != hit_e)
    both = (st == 1) & hit_e
    ds = np.abs(fa[:, 1] - ex[:, 1]) / texel
    print(f"{name}: rays {n}  exact hits {hit_e.sum()}  defer {np.mean(st == 2):.5f}  decided-wrong {wrong.sum()}  "
          f"max ds {ds[both].max() if both.any() else 0:.3g} texel  >1e-3: {(ds[both] > 1e-3).sum()}  "
          f"reasons {dict((i, int(c)) for i, c in enumerate(reasons) if c)}  nodes/ray {fa[:, 6].mean():.2f} (exact {ex[:, 5].mean():.2f})  tests/ray {fa[:, 7].mean():.2f} (exact {ex[:, 6].mean():.2f})")
    return wrong, both, ds


def shadow_rays(rays, ex, light, eps=1e-4):
    """approximate shadow rays: from primary hits, lifted along the radial direction, towards the light"""
    hit = ex[:, 0] > 0
    p = rays[hit, :3] + ex[hit, 1:2] * rays[hit, 3:]
    n = p / np.linalg.norm(p, axis=1, keepdims=True)
    o = p + eps * 10 * n
    d = np.array(light)[None] - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    lit = (n * d).sum(1) > 0
    return np.ascontiguousarray(np.concatenate([o, d], axis=1)[lit])


if __name__ == "__main__":
    l = build()
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        cases = [(720, 360, 160, 120, 4.242192793, 3), (1440, 720, 200, 150, 4.242192793, 11), (1440, 720, 256, 256, 0.6, 5)]
    elif which == "mid":
        cases = [(5760, 2880, 480, 270, 4.242192793, 7), (5760, 2880, 256, 256, 0.5, 7)]
    else:
        cases = [(23040, 11520, 480, 270, 4.242192793, 7), (23040, 11520, 256, 256, 0.3, 7)]
    for (W, H, iw, ih, fov, seed) in cases:
        t0 = time.time()
        counts = synth_ldem(W, H, seed=seed, craters=60)
        elev, rs = dorc.load_elevation(counts, 1)
        for use_i16 in (False, True):
            m = counts if use_i16 else elev
            kw = dict(scale=float(np.float32(0.5 / 1737400.0)), rs=rs) if use_i16 else {}
            eye = (0.0, -300.0, 0.0) if fov > 1 else (20.0, -298.0, 30.0)
            tgt = (0.0, 0.0, 0.0) if fov > 1 else (6.5, 0.0, 6.9)
            rays, shp = camera_rays(iw, ih, eye, tgt, DEFAULTS["up"], fov)
            ex = run_host(l, m, rays, **kw)
            fa = run_fast(l, m, rays, **kw)
            wrong, both, ds = report(f"{W}x{H} {'i16' if use_i16 else 'f32'} fov {fov} primary", W, ex, fa)
            for i in np.nonzero(wrong)[0][:5]:
                print("   wrong", i, "fast", fa[i, :6], "exact", ex[i, :3])
            for ph in (90.0, 60.0):
                sr = shadow_rays(rays, ex, sun_at_phase(ph))
                if len(sr) == 0:
                    continue
                exs = run_host(l, m, sr, any_hit=1, start_level=2, **kw)
                fas = run_fast(l, m, sr, start_level=2, **kw)
                wrong, both, ds = report(f"   shadow phase {ph}", W, exs, fas)
                for i in np.nonzero(wrong)[0][:5]:
                    print("   wrong", i, "fast", fas[i, :6], "exact", exs[i, :3])
        print(f"   ({time.time() - t0:.1f} s)")
