"""DEVELOPMENT TOOL: fast vs exact path on the CPU at FULL map resolution (92160 x 46080: the GPU-synthesised 5760 x 2880 map
tiled 16 x 16, so that cells are as small against float32 as in BASELINE config 3)."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from debug_parity import build, camera_rays, run_host
from debug_fast import run_fast, report, shadow_rays
from helpers import sun_at_phase, DEFAULTS

l = build()
rep = int(os.environ.get("REP", "16"))
small = np.load(os.path.join(ROOT, "gpurun_out", "synth_5760x2880.npy"))
t0 = time.time()
counts = np.ascontiguousarray(np.tile(small, (rep, rep)))
H, W = counts.shape
scale = float(np.float32(0.5 / 1737400.0))
m = np.float32(counts.max())
rs = float(np.float32(np.float32(m * np.float32(scale)) + np.float32(1)))
print("map", W, H, "tile", time.time() - t0, flush=True)
stride = int(os.environ.get("STRIDE", "8"))
rays, shp = camera_rays(3840, 2160, DEFAULTS["eye"], DEFAULTS["target"], DEFAULTS["up"], DEFAULTS["fov"], stride=stride)
# keep rays that can touch the sphere
b2 = (rays[:, :3] ** 2).sum(1) - ((rays[:, :3] * rays[:, 3:]).sum(1)) ** 2
rays = np.ascontiguousarray(rays[b2 < 100.5])
print("rays", len(rays), flush=True)
kw = dict(scale=scale, rs=rs)
t0 = time.time(); ex = run_host(l, counts, rays, **kw); print("exact", time.time() - t0, flush=True)
t0 = time.time(); fa = run_fast(l, counts, rays, **kw); print("fast", time.time() - t0, flush=True)
wrong, both, ds = report("primary", W, ex, fa)
st = fa[:, 0].astype(int)
np.save("/tmp/full_rays.npy", rays); np.save("/tmp/full_fast.npy", fa); np.save("/tmp/full_exact.npy", ex)
sr = shadow_rays(rays, ex, sun_at_phase(90.0))
exs = run_host(l, counts, sr, any_hit=1, start_level=2, **kw)
fas = run_fast(l, counts, sr, start_level=2, **kw)
report("shadow", W, exs, fas)
np.save("/tmp/full_srays.npy", sr); np.save("/tmp/full_sfast.npy", fas); np.save("/tmp/full_sexact.npy", exs)
