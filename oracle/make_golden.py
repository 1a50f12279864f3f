"""
TEST INFRASTRUCTURE - generates tests/golden/* by running the UNMODIFIED reference
(/root/reference, imported under oracle/ref_stub.py) in the build container.

    python oracle/make_golden.py

The reference tree does not travel to the GPU box, the fixtures do.  Every fixture
stores its inputs next to the reference outputs, so the tests never need to
regenerate the inputs with a particular numpy version.
"""

import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_stub                      # noqa: E402
from moonrtx_b200.synth import synth_ldem, synth_color, synth_ephemeris   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_elevation(dl, src_i16: np.ndarray, ds: int):
    d = tempfile.mkdtemp()
    p = os.path.join(d, "ldem.tif")
    with open(p, "wb") as f:
        f.write(b"stub")
    ref_stub.set_read_image(p, src_i16.view(np.uint16).copy())
    e, rs = dl.load_elevation_data(p, ds)
    return np.array(e), float(rs)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    dl = ref_stub.import_reference("data_loader")
    mr = ref_stub.import_reference("moon_renderer")

    # ---- A1/A2 elevation downscale + normalise -----------------------------
    rng = np.random.default_rng(7)
    maps = {
        "synth": synth_ldem(360, 180, seed=11, craters=40),
        "uniform": rng.integers(-32768, 32768, size=(180, 360), dtype=np.int32).astype(np.int16),
    }
    maps["uniform"][0, 0] = -32768
    maps["uniform"][179, 359] = 32767
    out = {}
    for name, m in maps.items():
        out[f"{name}_src"] = m
        for ds in (1, 2, 3, 4, 5, 6, 9, 12):
            e, rs = ref_elevation(dl, m, ds)
            out[f"{name}_ds{ds}"] = e
            out[f"{name}_ds{ds}_rs"] = np.float64(rs)
    np.savez_compressed(os.path.join(GOLDEN, "elevation_small.npz"), **out)

    # larger map, digests only (input regenerated from the stored small seed map by tiling)
    big = np.tile(maps["synth"], (8, 8))            # 2880 x 1440
    big = (big.astype(np.int32) + (np.arange(big.shape[1])[None, :] % 97) * 3
           - (np.arange(big.shape[0])[:, None] % 89) * 5).astype(np.int16)
    digests = {"recipe": "tile(synth,(8,8)) + (col%97)*3 - (row%89)*5", "cases": {}}
    for ds in (2, 3, 4, 5, 8, 15, 16, 32):
        e, rs = ref_elevation(dl, big, ds)
        digests["cases"][str(ds)] = {"sha256": sha(e), "radius_scale": rs, "shape": list(e.shape)}
    with open(os.path.join(GOLDEN, "elevation_digests.json"), "w") as f:
        json.dump(digests, f, indent=1)

    # ---- A3/A4 colour reduce + LUT ------------------------------------------
    import cv2
    col = synth_color(256, 128, seed=5)
    col[0, 0] = (0, 0, 0)
    col[0, 1] = (255, 255, 255)
    d = tempfile.mkdtemp()
    p = os.path.join(d, "color.tif")
    cv2.imwrite(p, col)
    cout = {"src_bgr": col}
    for k in (1, 2, 4, 8):
        for g in (2.2, 1.0):
            tex = dl.load_color_data(p, g, k)
            cout[f"k{k}_g{g}"] = np.array(tex)
            for ext in (".npy", ".json"):
                c = f"{p}.ds{k}{ext}"
                if os.path.exists(c):
                    os.remove(c)
    for g in (0.5, 1.0, 1.8, 2.2, 5.0):
        cout[f"lut_g{g}"] = dl._albedo_lut(g)
    np.savez_compressed(os.path.join(GOLDEN, "color_small.npz"), **cout)

    # ---- A5 texel convention: get_elevation_m / hit_to_selenographic ----------
    nav = ref_stub.import_reference("renderer_navigation")

    class Bare(nav.NavigationMixin):
        MOON_RADIUS = mr.MoonRenderer.MOON_RADIUS
        MOON_RADIUS_KM = mr.MoonRenderer.MOON_RADIUS_KM

    b = Bare.__new__(Bare)
    elev, rs = ref_elevation(dl, maps["synth"], 3)          # 120 x 60
    b.elevation = elev
    b.elevation_radius_scale = rs
    b.moon_rotation_inv = np.eye(3)
    g = np.random.default_rng(3)
    lats = np.concatenate([g.uniform(-90, 90, 200), [90.0, -90.0, 89.9, -89.9, 0.0, 0.0]])
    lons = np.concatenate([g.uniform(-180, 180, 200), [0.0, 0.0, 179.99, -179.99, -180.0, 180.0]])
    elev_m = np.array([b.get_elevation_m(float(la), float(lo)) for la, lo in zip(lats, lons)])
    pts = g.normal(size=(100, 3))
    pts = pts / np.linalg.norm(pts, axis=1, keepdims=True) * 10.0
    sel = np.array([b.hit_to_selenographic(*map(float, q)) for q in pts], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "convention.npz"), elevation=elev, radius_scale=np.float64(rs),
                        lats=lats, lons=lons, elev_m=elev_m, hit_pts=pts, hit_latlon=sel)

    # ---- A6/A7/A10 camera + light vectors ---------------------------------
    scene = {"cases": []}
    for minutes, dist, sund, bla in [(0, 384400.0, 1.496e8, -90.0), (600, 356500.0, 1.471e8, -75.0),
                                     (-2400, 406700.0, 1.521e8, 100.0), (12000, 384400.0, 1.496e8, 35.0)]:
        eph = synth_ephemeris(minutes)._replace(distance=dist, sun_distance=sund, bright_limb_angle=bla)
        r = mr.MoonRenderer.__new__(mr.MoonRenderer)
        r.moon_ephem = eph
        cam = r.default_camera
        sd_pos, sd_r = r.calculate_sun_disk()
        scene["cases"].append({
            "distance": dist, "sun_distance": sund, "phase_angle": eph.phase_angle,
            "bright_limb_angle": bla, "elongation": eph.elongation,
            "light_pos": [float(x) for x in r.calculate_light_pos()],
            "camera_eye": [float(x) for x in cam.eye], "camera_fov": float(cam.fov),
            "camera_distance": float(r.moon_camera_distance()),
            "apparent_radius": float(r.moon_apparent_radius()),
            "sun_light_radius": float(r.SUN_LIGHT_DISTANCE * r.SUN_RADIUS_KM / sund),
            "sun_disk_pos": [float(x) for x in sd_pos], "sun_disk_radius": float(sd_r),
        })
    scene["constants"] = {k: float(getattr(mr.MoonRenderer, k)) for k in
                          ("MOON_RADIUS", "CAMERA_DISTANCE", "MOON_FILL_FRACTION", "SUN_LIGHT_DISTANCE",
                           "SUN_RADIUS", "SUN_BRIGHTNESS_SCALE", "SCENE_EPSILON", "MARCHING_STEP",
                           "MARCHING_STEP_EPS", "ACCUMULATION_FRAMES", "MOON_REFERENCE_DISTANCE")}
    with open(os.path.join(GOLDEN, "scene_vectors.json"), "w") as f:
        json.dump(scene, f, indent=1)
    print("golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    main()
