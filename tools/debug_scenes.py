"""DEVELOPMENT TOOL: host traversal vs oracle over several scenes (primary hits)."""
import math, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from debug_parity import build, camera_rays, run_host
from oracle import downscale_oracle as dorc
from moonrtx_b200.synth import synth_ldem
from helpers import make_oracle, DEFAULTS

l = build()
def check(name, elev, iw, ih, stride=1, **kw):
    p = {**DEFAULTS, **kw}
    orc = make_oracle(elev, iw, ih, **kw)
    rays, shp = camera_rays(iw, ih, p["eye"], p["target"], p["up"], p["fov"], p["u"], p["v"], stride)
    t = time.time(); out = run_host(l, elev, rays); th = time.time() - t
    t = time.time(); o = orc.render(stride=stride); to = time.time() - t
    ref = o["hit64"].reshape(-1, 4)
    texel = 2 * math.pi * 10 / elev.shape[1]
    gh, oh = out[:, 0] > 0, ref[:, 0] > 0
    both = gh & oh
    ds = np.abs(out[:, 1] - ref[:, 0]) / texel
    print(f"{name}: rays {len(rays)} hits {both.sum()} mismatch {(gh != oh).sum()} bad {(both & (ds > 1e-3)).sum()} "
          f"max_ds {ds[both].max() if both.any() else 0:.3g} nodes/hit {out[both, 5].mean():.1f} tests/hit {out[both, 6].mean():.2f} "
          f"overflow {out[:, 7].sum():.0f} oracle_cells_min {o['stats'].min()} host {th:.1f}s oracle {to:.1f}s")
    return out, ref

elev720, _ = dorc.load_elevation(synth_ldem(720, 360, seed=3, craters=60), 1)
check("whole disk 720", elev720, 160, 120)
elev1440, _ = dorc.load_elevation(synth_ldem(1440, 720, seed=5, craters=60), 1)
a, b = math.radians(7.0), math.radians(-5.0)
Rz = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
Rx = np.array([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]])
Rm = Rz @ Rx
check("rotated narrow", elev1440, 128, 96, u=tuple(Rm[:, 2]), v=tuple(-Rm[:, 1]), eye=(20.0, -298.0, 30.0), target=(6.5, 0.0, 6.9), fov=0.6)
check("polar", elev720, 96, 96, eye=(0.0, 0.0, 300.0), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=1.0)
check("south polar grazing", elev720, 96, 96, eye=(0.0, -300.0, -40.0), target=(0.0, 0.0, -9.9), fov=0.5)
check("limb east", elev1440, 128, 64, eye=(0.0, -300.0, 0.0), target=(9.95, 0.0, 0.0), fov=0.4)
if len(sys.argv) > 1:
    big, _ = dorc.load_elevation(synth_ldem(5760, 2880, seed=7, craters=300), 1)
    check("cfg2-like 5760", big, 1920, 1080, stride=24)
def show(name, elev, iw, ih, **kw):
    out, ref = check(name, elev, iw, ih, **kw)
    texel = 2 * math.pi * 10 / elev.shape[1]
    both = (out[:, 0] > 0) & (ref[:, 0] > 0)
    ds = np.abs(out[:, 1] - ref[:, 0]) / texel
    for i in np.nonzero(both & (ds > 1e-3))[0][:5]:
        print(" bad", i, "host s", out[i, 1], "oracle s", ref[i, 0], "host lon/lat", np.degrees(out[i, 3:5]), "oracle", np.degrees(ref[i, 2:4]))
show("polar", elev720, 96, 96, eye=(0.0, 0.0, 300.0), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=1.0)
show("south polar grazing", elev720, 96, 96, eye=(0.0, -300.0, -40.0), target=(0.0, 0.0, -9.9), fov=0.5)
