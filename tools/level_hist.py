"""Development helper (CPU): node visits of the filtered walk by pyramid level, camera rays and sun rays of BASELINE config 3
(every 16th pixel), host build of the traversal.   python tools/level_hist.py /tmp/synth_92160x46080.npy"""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from debug_parity import build, camera_rays
from debug_fast import run_fast, shadow_rays
from helpers import sun_at_phase
SCALE = float(np.float32(0.5 / 1737400.0))
l = build()
counts = np.ascontiguousarray(np.load(sys.argv[1], mmap_mode="r"))
rs = float(np.float32(np.float32(np.float32(counts.max()) * np.float32(SCALE)) + np.float32(1)))
kw = dict(scale=SCALE, rs=rs)
def hist(reset=True):
    out = (C.c_ulonglong * 32)()
    l.dbg_level_hist(out, 1 if reset else 0)
    return [int(v) for v in out][:12]
rays, _ = camera_rays(3840, 2160, (0, -300, 0), (0, 0, 0), (0, 0, 1), 4.242192793, stride=16)
hist()
prim = run_fast(l, counts, rays, start_level=-3, **kw)
hp = hist()
st = prim[:, 0].astype(int) & 3
hits = np.stack([(st == 1).astype(float), prim[:, 1]], axis=1)
sr = shadow_rays(rays, hits, sun_at_phase(90.0))
shad = run_fast(l, counts, sr, start_level=2, **kw)
hs = hist()
for name, h in (("camera rays", hp), ("sun rays", hs)):
    t = sum(h)
    print(name, "node visits by level 0..11:", h, "share of levels >= 8: %.3f" % (sum(h[8:]) / max(t, 1)), "levels >= 6: %.3f" % (sum(h[6:]) / max(t, 1)))
