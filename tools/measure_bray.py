"""Development helper (CPU, no GPU): B_ray of SURVEY.md 8d - bytes of pyramid + base map a max-mip traversal of a ray must
read - measured with the host build of the traversal core (tools/trace_host.cu: the same walk the kernels run, compiled
for the CPU) on the every-16th-pixel sub-grid of BASELINE configs 2, 3 and 5: 32 B per node visit (one sector holds the
cell's max), 2 x 32 B per patch test (the 2 x 2 bilinear patch spans two rows).  Writes tests/golden/bray.json; the
numbers are frozen in BASELINE.md.   python tools/measure_bray.py /tmp/synth_92160x46080.npy"""
import json, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from debug_parity import build, camera_rays
from debug_fast import run_fast, shadow_rays
from helpers import sun_at_phase

SCALE = float(np.float32(0.5 / 1737400.0))


def measure(l, elev, iw, ih, fov, kw, name):
    rays, _ = camera_rays(iw, ih, (0, -300, 0), (0, 0, 0), (0, 0, 1), fov, stride=16)
    prim = run_fast(l, elev, rays, start_level=-3, **kw)
    inside = prim[:, 6] > 0
    st = prim[:, 0].astype(int) & 3
    hits = np.stack([(st == 1).astype(float), prim[:, 1]], axis=1)
    sr = shadow_rays(rays, hits, sun_at_phase(90.0))
    shad = run_fast(l, elev, sr, start_level=2, **kw)
    out = {"grid": f"every 16th pixel of {iw}x{ih}, fov {fov}", "primary_rays_in_sphere": int(inside.sum()), "shadow_rays": int(len(sr)),
           "primary_nodes_per_ray": float(prim[inside, 6].mean()), "primary_tests_per_ray": float(prim[inside, 7].mean()),
           "shadow_nodes_per_ray": float(shad[:, 6].mean()), "shadow_tests_per_ray": float(shad[:, 7].mean())}
    out["B_ray_primary"] = round(32.0 * (out["primary_nodes_per_ray"] + 2.0 * out["primary_tests_per_ray"]), 1)
    out["B_ray_shadow"] = round(32.0 * (out["shadow_nodes_per_ray"] + 2.0 * out["shadow_tests_per_ray"]), 1)
    n = out["primary_rays_in_sphere"] + out["shadow_rays"]
    out["B_ray"] = round((out["B_ray_primary"] * out["primary_rays_in_sphere"] + out["B_ray_shadow"] * out["shadow_rays"]) / n, 1)
    print(name, json.dumps(out))
    return out


if __name__ == "__main__":
    l = build()
    counts = np.load(sys.argv[1], mmap_mode="r")
    counts = np.ascontiguousarray(counts)
    H, W = counts.shape
    rs = float(np.float32(np.float32(np.float32(counts.max()) * np.float32(SCALE)) + np.float32(1)))
    kw = dict(scale=SCALE, rs=rs)
    res = {"map": f"{W}x{H} int16 synthetic LDEM (host port of the generator), sun at phase 90 deg, pixel-centre rays",
           "how": "tools/measure_bray.py: host build of the traversal core, 32 B per node visit + 64 B per patch test"}
    t0 = time.time()
    res["config3"] = measure(l, counts, 3840, 2160, 4.242192793, kw, "config 3")
    res["config5"] = measure(l, counts, 7680, 4320, 2.5, kw, "config 5")
    # config 2: the ds = 16 map (5760 x 2880 float32) of the same LDEM, as load_elevation_data produces it
    from oracle import downscale_oracle as orc
    ds = W // 5760
    elev, _ = orc.load_elevation(counts, ds)
    del counts
    res["config2"] = measure(l, elev, 1920, 1080, 4.242192793, {}, "config 2")
    res["seconds"] = round(time.time() - t0, 1)
    with open(os.path.join(ROOT, "tests", "golden", "bray.json"), "w") as f:
        json.dump(res, f, indent=1)
