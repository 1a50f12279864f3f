"""DEVELOPMENT TOOL: host build of the traversal core vs the oracle on a test scene."""
import ctypes as C, math, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import downscale_oracle as dorc
from moonrtx_b200.synth import synth_ldem
from helpers import make_oracle, sun_at_phase, DEFAULTS

SO = os.path.join(ROOT, "tools", "_build", "libtrace_host.so")      # (git-ignored)
def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["nvcc", "-O2", "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off", "-shared", "-o", SO,
                    os.path.join(ROOT, "tools", "trace_host.cu")], check=True, stderr=subprocess.PIPE)
    l = C.CDLL(SO)
    l.dbg_trace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                            C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p]
    return l

def camera_rays(img_w, img_h, eye, target, up, fov, u=(0,0,1), v=(0,-1,0), stride=1):
    eye = np.array(eye, float); w = np.array(target, float) - eye; w /= np.linalg.norm(w)
    right = np.cross(w, np.array(up, float)); right /= np.linalg.norm(right); up2 = np.cross(right, w)
    t = math.tan(math.radians(fov) / 2); asp = img_w / img_h
    ys, xs = np.meshgrid(np.arange(0, img_h, stride), np.arange(0, img_w, stride), indexing="ij")
    sx = ((xs + 0.5) / img_w * 2 - 1) * t * asp; sy = (1 - (ys + 0.5) / img_h * 2) * t
    d = w[None, None] + sx[..., None] * right + sy[..., None] * up2
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    ez = np.array(u, float); ez /= np.linalg.norm(ez); vv = np.array(v, float); vv -= vv.dot(ez) * ez; vv /= np.linalg.norm(vv)
    ex = np.cross(ez, vv); Rm = np.stack([ex, -vv, ez])
    ob = Rm @ eye; db = d @ Rm.T
    rays = np.concatenate([np.broadcast_to(ob, db.shape), db], axis=-1)
    return np.ascontiguousarray(rays.reshape(-1, 6)), ys.shape

def run_host(l, elev, rays, s_min=0.0, any_hit=0, scale=0.0, rs=1.0, start_level=-3):
    out = np.zeros((len(rays), 8))
    if elev.dtype == np.int16:
        m = np.float32(elev.max()); dmax = float(np.float32(np.float32(np.float32(m*np.float32(scale))+np.float32(1))/np.float32(rs)))
    else:
        dmax = float(elev.max())
    l.dbg_trace(elev.ctypes.data, int(elev.dtype == np.int16), elev.shape[1], elev.shape[0], scale, rs, dmax,
                rays.ctypes.data, len(rays), s_min, 10.0, any_hit, start_level, out.ctypes.data)
    return out

if __name__ == "__main__":
    l = build()
    W, H, iw, ih = 720, 360, 160, 120
    elev, _ = dorc.load_elevation(synth_ldem(W, H, seed=3, craters=60), 1)
    orc = make_oracle(elev, iw, ih, light_pos=sun_at_phase(90.0))
    rays, shp = camera_rays(iw, ih, DEFAULTS["eye"], DEFAULTS["target"], DEFAULTS["up"], DEFAULTS["fov"])
    out = run_host(l, elev, rays)
    ref = np.array([np.concatenate([[h], o[:4]]) for h, o in (orc.trace_ray(r[:3], r[3:]) for r in rays)])
    texel = 2 * math.pi * 10 / W
    both = (out[:, 0] > 0) & (ref[:, 0] > 0)
    print("hit mismatch", int(((out[:, 0] > 0) != (ref[:, 0] > 0)).sum()), "hits", int(both.sum()))
    ds = np.abs(out[:, 1] - ref[:, 1]) / texel
    bad = np.nonzero(both & (ds > 1e-3))[0]
    print("bad", len(bad), "max", ds[both].max(), "nodes/ray", out[:, 5].mean(), "tests/ray", out[:, 6].mean(), "overflow", out[:, 7].sum())
    for i in bad[:10]:
        print(i, divmod(i, iw), "host s", out[i, 1], "oracle s", ref[i, 1], "ds_texel", ds[i], "lon/lat", np.degrees(ref[i, 3:5]), np.degrees(out[i, 3:5]))
    if len(sys.argv) > 1:
        i = int(sys.argv[1])
        l.dbg_set(1)
        run_host(l, elev, rays[i:i+1])
        l.dbg_set(0)
        print("oracle:", orc.trace_ray(rays[i, :3], rays[i, 3:]))
    if len(sys.argv) > 3:
        i = int(sys.argv[1]); s0 = float(sys.argv[2]); s1 = float(sys.argv[3])
        for sv in np.linspace(s0, s1, 25):
            p = rays[i, :3] + sv * rays[i, 3:]
            r = np.linalg.norm(p); lon = math.degrees(math.atan2(p[0], -p[1])); lat = math.degrees(math.atan2(p[2], math.hypot(p[0], p[1])))
            print(f"s={sv:.7f} lon={lon:.5f} lat={lat:.5f} f={r - 10*orc.displacement(lat, lon):+.3e}")
