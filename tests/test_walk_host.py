"""
The traversal core the kernels run (csrc/trace_fast.cuh: walk_begin2 / walk_step / walk_advance / fast_test, and the
float64 exact walk of csrc/trace_core.cuh) is `__host__ __device__` source.  tools/trace_host.cu compiles the same
headers for the CPU; this test runs that build on camera and sun rays of a small scene and compares every ray with the
float64 oracle - so the CPU suite sees a broken cell step or patch test without a GPU.  (The GPU tests compare the
kernels themselves with the same oracle; this is the same arithmetic one ray per thread.)
Tolerance: north_star's hit radius / ray parameter <= 1e-3 texel; hit / miss decisions of rays the float32 filter
decides must equal the oracle's, and it may hand at most 2 % of the rays of these small maps to the float64 referee.
"""
import math
import os
import shutil
import sys

import numpy as np
import pytest

from oracle.render_oracle import OracleScene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 10.0


@pytest.fixture(scope="module")
def host():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH: the host build of the traversal core cannot be compiled")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import debug_parity
    import debug_fast
    return debug_parity.build(), debug_parity, debug_fast


def relief(W, H, seed):
    """A smooth float32 displacement map with a few km of relief (mare-like swells + crater bowls), max exactly 1."""
    rng = np.random.default_rng(seed)
    lon = (np.arange(W) + 0.5) / W * 2 * np.pi
    lat = (0.5 - (np.arange(H) + 0.5) / H) * np.pi
    lo, la = np.meshgrid(lon, lat)
    x, y, z = np.cos(la) * np.sin(lo), -np.cos(la) * np.cos(lo), np.sin(la)
    h = np.zeros((H, W))
    for _ in range(150):
        c = rng.normal(size=3); c /= np.linalg.norm(c)
        ang = np.arccos(np.clip(x * c[0] + y * c[1] + z * c[2], -1, 1))
        rad = rng.uniform(0.04, 0.2)
        t = np.clip(ang / rad, 0, 1.6)
        # (depth ~ a fifth of the radius: slopes of ten degrees and more, so that a low Sun casts shadows)
        h += rng.uniform(-1.0, 0.5) * 0.2 * rad * np.where(t < 1, 1 - t * t, -0.35 * np.exp(-8 * (t - 1)))
    d = 1.0 + h - h.max()
    return np.ascontiguousarray(d.astype(np.float32))


def oracle_trace(sc, rays):
    out = np.zeros((len(rays), 2))
    for i, q in enumerate(rays):
        hit, o = sc.trace_ray(q[:3], q[3:])
        out[i] = (1.0 if hit else 0.0, o[0])
    return out


def check(name, W, fast, orc, max_defer):
    texel = 2 * math.pi * R / W
    st = fast[:, 0].astype(int) & 3
    decided = st != 2
    hit_o = orc[:, 0] > 0
    wrong = decided & ((st == 1) != hit_o)
    assert decided.mean() >= 1.0 - max_defer, f"{name}: {1 - decided.mean():.4f} of the rays deferred"
    # a ray whose decision differs must be a grazing one: never seen on these maps
    assert int(wrong.sum()) == 0, f"{name}: {int(wrong.sum())} decided rays disagree with the oracle on hit / miss"
    both = (st == 1) & hit_o
    assert int(both.sum()) > 1000, name
    ds = np.abs(fast[both, 1] - orc[both, 1]) / texel
    assert float(ds.max()) <= 1.0e-3, f"{name}: hit parameter off by {ds.max():.3g} texel"
    return both


@pytest.mark.parametrize("W,H,seed", [(720, 360, 5), (1440, 720, 6)])
def test_host_build_of_the_float32_walk_matches_the_oracle(host, W, H, seed):
    l, dp, df = host
    elev = relief(W, H, seed)
    sc = OracleScene(elev)
    # camera rays of the default whole-disk view (moon_renderer.py:85-101), every pixel of a 128 x 96 frame
    rays, _ = dp.camera_rays(128, 96, (0, -300, 0), (0, 0, 0), (0, 0, 1), 4.242192793)
    fast = df.run_fast(l, elev, rays, start_level=-3)
    orc = oracle_trace(sc, rays)
    both = check("camera rays", W, fast, orc, 0.02)
    assert fast[fast[:, 6] > 0, 6].mean() < 40.0                  # nodes per ray: the pyramid is being used
    # sun rays from the oracle's hit points, Sun 10 degrees above the terminator horizon (long grazing shadows)
    hits = np.stack([orc[:, 0], orc[:, 1]], axis=1)
    ph = math.radians(80.0)
    sun = (21460.0 * math.sin(ph), -21460.0 * math.cos(ph), 0.0)
    srays = df.shadow_rays(rays, hits, sun)
    assert len(srays) > 1000
    sfast = df.run_fast(l, elev, srays, start_level=2)
    sorc = oracle_trace(sc, srays)
    st = sfast[:, 0].astype(int) & 3
    decided = st != 2
    assert decided.mean() >= 0.97
    assert int((decided & ((st == 1) != (sorc[:, 0] > 0))).sum()) == 0
    assert 0 < int((sorc[:, 0] > 0).sum()) < len(srays)           # some of them are in shadow, some are not


def test_host_build_of_the_float64_walk_matches_the_oracle(host):
    """The exact walk (A/B kernels 0 / 1 and the referee's patch test) on the same rays: every decision and every hit."""
    l, dp, df = host
    elev = relief(720, 360, 7)
    sc = OracleScene(elev)
    rays, _ = dp.camera_rays(96, 72, (0, -300, 0), (0, 0, 0), (0, 0, 1), 4.242192793)
    ex = dp.run_host(l, elev, rays)
    orc = oracle_trace(sc, rays)
    hit_e, hit_o = ex[:, 0] > 0, orc[:, 0] > 0
    assert int((hit_e != hit_o).sum()) == 0
    texel = 2 * math.pi * R / 720
    ds = np.abs(ex[hit_e, 1] - orc[hit_o, 1]) / texel
    assert float(ds.max()) <= 1.0e-6


def test_host_build_of_the_float32_walk_on_int16_texels_matches_the_oracle(host):
    """The production form of the map: int16 LOLA counts decoded in the walk (height = 1 + count * SCALE, radius_scale as
    data_loader.py:232-242 computes it at downscale 1), against the oracle reading the same counts."""
    l, dp, df = host
    W, H = 1440, 720
    scale = float(np.float32(0.5 / 1737400.0))
    d = relief(W, H, 9).astype(np.float64)
    counts = np.round((d - 1.0) * 0.12 / scale).astype(np.int16)            # up to ~8 km of relief: inside the int16 range
    counts -= counts.min() // 2                                              # heights on both sides of the datum
    rs = float(np.float32(np.float32(np.float32(counts.max()) * np.float32(scale)) + np.float32(1)))
    sc = OracleScene(counts, scale=scale, radius_scale=rs)
    rays, _ = dp.camera_rays(128, 96, (0, -300, 0), (0, 0, 0), (0, 0, 1), 4.242192793)
    fast = df.run_fast(l, counts, rays, scale=scale, rs=rs, start_level=-3)
    orc = oracle_trace(sc, rays)
    check("camera rays, int16 map", W, fast, orc, 0.02)
