"""
The float64 render oracle against what the reference itself can state (convention fixtures
generated from renderer_navigation.py) and analytic known answers.  CPU only.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle.render_oracle import OracleScene

R = 10.0


def lonlat_to_body(lon_deg, lat_deg, r=1.0):
    lo, la = math.radians(lon_deg), math.radians(lat_deg)
    return np.array([r * math.cos(la) * math.sin(lo), -r * math.cos(la) * math.cos(lo), r * math.sin(la)])


@pytest.fixture(scope="module")
def conv(golden_dir):
    return np.load(os.path.join(golden_dir, "convention.npz"))


def test_bilinear_convention_matches_get_elevation_m(conv):
    """orc_displacement == the reference's get_elevation_m (renderer_navigation.py:558-599)."""
    sc = OracleScene(conv["elevation"])
    rs = float(conv["radius_scale"])
    for la, lo, ref in zip(conv["lats"], conv["lons"], conv["elev_m"]):
        got = (sc.displacement(la, lo) * rs - 1.0) * 1737.4 * 1000.0
        # the reference evaluates the bilinear sum on np.float32 scalars (NEP 50 keeps them float32),
        # so its answer carries float32 rounding: 6e-8 * 1 737 400 m = 0.1 m per operation
        assert got == pytest.approx(ref, abs=0.6), (la, lo)


def test_hit_lonlat_matches_hit_to_selenographic(conv):
    """lon/lat of an oracle hit == the reference's hit_to_selenographic (renderer_navigation.py:452-492)."""
    sc = OracleScene(np.ones((32, 64), dtype=np.float32))
    for p, (lat, lon) in zip(conv["hit_pts"], conv["hit_latlon"]):
        o = p * 3.0                       # outside, on the radial line through the golden point
        hit, out = sc.trace_ray(o, -p)
        assert hit
        assert out[0] == pytest.approx(20.0, abs=1e-9)        # |o| - R
        assert math.degrees(out[3]) == pytest.approx(lat, abs=1e-9)
        dlon = (math.degrees(out[2]) - lon + 180.0) % 360.0 - 180.0
        assert dlon == pytest.approx(0.0, abs=1e-9)


@pytest.mark.parametrize("d0", [1.0, 0.97])
def test_flat_sphere_analytic(d0):
    elev = np.full((45, 90), d0, dtype=np.float32)
    sc = OracleScene(elev)
    rng = np.random.default_rng(1)
    rr = R * float(np.float32(d0))
    for _ in range(200):
        o = rng.normal(size=3)
        o = o / np.linalg.norm(o) * rng.uniform(12, 300)
        tgt = rng.normal(size=3)
        tgt = tgt / np.linalg.norm(tgt) * rng.uniform(0, 11.5)
        d = (tgt - o) / np.linalg.norm(tgt - o)
        b = o.dot(d)
        disc = b * b - (o.dot(o) - rr * rr)
        hit, out = sc.trace_ray(o, d)
        if disc <= 1e-9:
            assert not hit or disc > -1e-9
            continue
        s = -b - math.sqrt(disc)
        assert hit
        assert out[0] == pytest.approx(s, abs=1e-9)
        assert out[1] == pytest.approx(rr, abs=1e-10)
        p = o + s * d
        assert np.allclose(out[4:7], p / np.linalg.norm(p), atol=1e-9)       # radial normal


def test_hit_radius_equals_surface_at_hit_lonlat():
    """Internal cross-check (SURVEY.md §8c i): the hit radius reproduces bilinear D at the hit's own lon/lat."""
    from moonrtx_b200.synth import synth_ldem
    from oracle import downscale_oracle as orc
    elev, _ = orc.load_elevation(synth_ldem(360, 180, seed=3, craters=40), 1)
    sc = OracleScene(elev, img_w=96, img_h=96)
    out = sc.render()
    h = out["hit64"].reshape(-1, 4)
    hits = h[h[:, 0] > 0]
    assert len(hits) > 3000
    for s, r, lon, lat in hits[::37]:
        assert r == pytest.approx(R * sc.displacement(math.degrees(lat), math.degrees(lon)), abs=1e-10)


def test_peak_shadow_length():
    """A single peak of height h under a sun of altitude alt casts a shadow of length h / tan(alt)
    (the reference's own check, moon_renderer.py:94-97), here with the sphere's curvature included."""
    W, H = 1440, 720
    base = 0.99
    elev = np.full((H, W), base, dtype=np.float32)
    r_eq, c0 = H // 2, W // 4 * 2            # a texel next to the equator, lon ~ 0
    elev[r_eq, c0] = 1.0
    sc = OracleScene(elev, light_radius=0.0)
    lat_p = 90.0 - (r_eq + 0.5) / H * 180.0
    lon_p = (c0 + 0.5) / W * 360.0 - 180.0
    h = R * (1.0 - base)                     # 0.1 units = 17 km
    Rb = R * base
    apex = lonlat_to_body(lon_p, lat_p)
    east = np.array([math.cos(math.radians(lon_p)), math.sin(math.radians(lon_p)), 0.0])
    for alt_deg in (2.0, 5.0, 10.0):
        # sun direction: altitude alt above the local horizon at the apex, azimuth due west
        sun = math.sin(math.radians(alt_deg)) * apex - math.cos(math.radians(alt_deg)) * east
        for theta_deg in np.linspace(0.3, 12.0, 60):
            # surface point east of the peak (downstream of the light) on the same latitude circle ~ great circle
            P = Rb * lonlat_to_body(lon_p + theta_deg, lat_p)
            n = P / np.linalg.norm(P)
            if n.dot(sun) <= 0:
                continue
            # height of the ray P + t sun above the apex direction: solve in the plane (apex, east)
            # point on the ray closest to the apex axis direction: intersect with the line {lam * apex}
            # 2-D: coordinates along apex (y) and east (x)
            px, py = P.dot(east), P.dot(apex)
            sx, sy = sun.dot(east), sun.dot(apex)
            t = -px / sx
            y0 = py + t * sy - Rb          # height over the base sphere where the ray crosses the apex axis
            o = P + 1e-7 * n
            hit, out = sc.trace_ray(o, sun)
            margin = 0.02 * h
            if y0 < h - margin and y0 > 0.2 * h:
                assert hit, (alt_deg, theta_deg, y0)
            if y0 > h + margin:
                assert not hit, (alt_deg, theta_deg, y0)


def test_lambert_terminator_and_subsolar_brightness(golden_dir):
    """Flat sphere lit from +X: L = (brightness/100) * albedo * cos(theta) * (r_light/100)^2-scaled irradiance,
    zero beyond the terminator."""
    with open(os.path.join(golden_dir, "scene_vectors.json")) as f:
        sv = json.load(f)
    case = sv["cases"][0]
    K = sv["constants"]
    brightness = 80.0
    sc = OracleScene(np.ones((45, 90), dtype=np.float32), img_w=128, img_h=128,
                     eye=case["camera_eye"], fov=case["camera_fov"], light_pos=case["light_pos"],
                     light_radius=case["sun_light_radius"], light_radiance=brightness * K["SUN_BRIGHTNESS_SCALE"])
    out = sc.render()
    acc = out["accum"]
    h64 = out["hit64"]
    lon = h64[..., 2]
    hit = h64[..., 0] > 0
    lit = acc[..., 0] > 0
    # sun at +X (bright limb -90 deg, phase 90 deg): lit exactly where the normal has a positive x
    nx = np.cos(h64[..., 3]) * np.sin(lon)
    assert np.all(lit[hit & (nx > 1e-3)])
    assert not np.any(lit[hit & (nx < -1e-3)])
    assert not np.any(lit[~hit])
    # brightness follows cos(theta): compare every lit pixel with the closed form
    # light at finite distance D = 21460 on +X: cos = (D nx - R) / dist (terminator parallax, moon_renderer.py:58-62)
    D = case["light_pos"][0]
    dist = np.sqrt((D - R * nx) ** 2 + (R ** 2 - (R * nx) ** 2))
    expect = (brightness * K["SUN_BRIGHTNESS_SCALE"]) * (case["sun_light_radius"] / dist) ** 2 * (D * nx - R) / dist
    sel = hit & (nx > 0.05)
    assert np.allclose(acc[..., 0][sel], expect[sel], rtol=1e-9)
    assert expect[sel].max() == pytest.approx(brightness / 100.0 * (case["sun_light_radius"] / 100.0) ** 2, rel=0.05)


def test_tonemap_known_values():
    sc = OracleScene(np.ones((8, 16), dtype=np.float32))
    acc = np.array([[0.0, 0.5, 2.0, 1.0], [1.0 / 0.9, 0.1, 0.01, 1.0]])
    rgba = sc.tonemap(acc)
    assert list(rgba[0]) == [0, int(math.floor((0.45 ** (1 / 2.2)) * 255 + 0.5)), 255, 255]
    assert rgba[1][0] == 255
