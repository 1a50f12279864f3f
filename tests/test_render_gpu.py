"""
Parity of the CUDA render path (through the C ABI, driven with the reference's own rt.* call
sequence) against the float64 oracle.  Tolerances are BASELINE.json's: hit-radius error
<= 1e-3 texel, 8-bit image MAE <= 1 and PSNR >= 40 dB at deterministic 1 spp.
"""
import math

import numpy as np
import pytest

from helpers import image_metrics, make_gpu, make_oracle, penetration_texels, sun_at_phase

pytestmark = pytest.mark.gpu

R = 10.0


def synth_elevation(W, H, seed=3, ds=1):
    from moonrtx_b200.synth import synth_ldem
    from moonrtx_b200.data_loader import downscale_elevation
    return downscale_elevation(synth_ldem(W, H, seed=seed, craters=60), ds)


def compare(rt, orc, stride=1, texel_tol=1e-3, allow_mismatch=0):
    """Render both; returns metrics after asserting the hit-radius and image tolerances.  There is no outlier allowance: a
    pixel on which the two disagree (hit / miss, or hit radius beyond the tolerance) must be shown to be a grazing case -
    the oracle's own surface function along that ray dips below zero by less than the tolerance (a crest the ray touches) -
    and allow_mismatch only bounds how many such silhouette pixels a scene may have."""
    img = rt.render_cycle().copy()
    g = rt.get_hit_records_f64()[::stride, ::stride]
    o = orc.render(stride=stride)
    oh = o["hit64"]
    assert g.shape == oh.shape
    ghit, ohit = g[..., 0] > 0, oh[..., 0] > 0
    both = ghit & ohit
    texel = 2.0 * math.pi * R / orc.s.W                       # scene units per texel at the equator
    dr = np.abs(g[..., 1] - oh[..., 1]) / texel
    ds_ = np.abs(g[..., 0] - oh[..., 0])[both] / texel
    suspects = np.argwhere((ghit != ohit) | (both & (dr > texel_tol)))
    assert len(suspects) <= allow_mismatch, f"{len(suspects)} pixels disagree on hit / miss or hit radius"
    clean = both.copy()
    for (j, i) in suspects:
        s_root = g[j, i, 0] if ghit[j, i] else oh[j, i, 0]
        depth = penetration_texels(orc, int(i) * stride, int(j) * stride, float(s_root), texel)
        assert depth < texel_tol, (f"pixel ({i * stride}, {j * stride}): gpu hit {bool(ghit[j, i])} s={g[j, i, 0]:.9f}, oracle hit "
                                   f"{bool(ohit[j, i])} s={oh[j, i, 0]:.9f}, the ray goes {depth:.3g} texel below the surface: not a grazing case")
        clean[j, i] = False
    ref_img = orc.tonemap(o["accum"])
    mae, psnr = image_metrics(img[::stride, ::stride], ref_img)
    assert mae <= 1.0 and psnr >= 40.0, (mae, psnr)
    return {"max_dr_texel": float(dr[clean].max()) if clean.any() else 0.0, "max_ds_texel": float(ds_.max()) if len(ds_) else 0.0,
            "mae": mae, "psnr": psnr, "hits": int(both.sum()), "grazing_outliers": len(suspects), "img": img, "oracle": o}


def test_flat_sphere_is_analytic():
    elev = np.ones((45, 90), dtype=np.float32)
    rt = make_gpu(elev, 96, 96, shadows=False)
    rt.render_cycle()
    g = rt.get_hit_records_f64()
    hits = g[g[..., 0] > 0]
    assert len(hits) > 4000
    assert np.allclose(hits[:, 1], R, atol=1e-9)
    # s = distance from the eye (0,-300,0) to the sphere along the pixel ray
    yy, xx = np.nonzero(g[..., 0] > 0)
    t = math.tan(math.radians(4.242192793) / 2)
    sx = ((xx + 0.5) / 96 * 2 - 1) * t
    sy = (1 - (yy + 0.5) / 96 * 2) * t
    d = np.stack([sx, np.ones_like(sx), sy], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = -300.0 * d[:, 1]
    s = -b - np.sqrt(b * b - (300.0 ** 2 - R * R))
    assert np.allclose(hits[:, 0], s, atol=1e-8)
    rt.close()


@pytest.mark.parametrize("phase", [90.0, 30.0, 150.0])
def test_whole_disk_matches_oracle(phase):
    elev, _ = synth_elevation(720, 360)
    from moonrtx_b200.synth import synth_color
    from moonrtx_b200.data_loader import color_texture
    tex = color_texture(synth_color(256, 128), 2.2, 1)
    kw = dict(light_pos=sun_at_phase(phase))
    rt = make_gpu(elev, 160, 120, texture=tex, **kw)
    orc = make_oracle(elev, 160, 120, texture=tex, **kw)
    m = compare(rt, orc, allow_mismatch=3)
    assert m["hits"] > 4000
    c = rt.counters()
    assert c["primary_rays"] == 160 * 120 and abs(c["primary_hits"] - m["hits"]) <= 3
    assert 0 < c["shadow_rays"] <= c["primary_hits"] and c["shadow_occluded"] <= c["shadow_rays"]
    assert c["overflow"] == 0
    rt.close()


def test_int16_surface_equals_float32_surface():
    from moonrtx_b200.synth import synth_ldem
    from moonrtx_b200.data_loader import downscale_elevation
    counts = synth_ldem(720, 360, seed=9, craters=60)
    elev, rs = downscale_elevation(counts, 1)
    scale = float(np.float32(0.5 / 1737400.0))
    kw = dict(light_pos=sun_at_phase(80.0))
    rt_f = make_gpu(elev, 128, 128, **kw)
    rt_i = make_gpu(counts, 128, 128, scale=scale, radius_scale=rs, **kw)
    a = rt_f.render_cycle().copy()
    b = rt_i.render_cycle().copy()
    # same surface, same rays; the traversal bounds are decoded differently (one FMA for int16), so the
    # float32 root search starts from windows that differ in the last bits: equal to rounding, not bitwise
    hf, hi = rt_f.get_hit_records_f64(), rt_i.get_hit_records_f64()
    assert np.array_equal(hf[..., 0] > 0, hi[..., 0] > 0)
    assert np.allclose(hf, hi, rtol=0, atol=1e-7)
    assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 1
    # the exact kernel decodes both maps identically: bit-equal there
    for rt in (rt_f, rt_i):
        rt.set_uint("kernel", 1)
    a1 = rt_f.render_cycle().copy()
    b1 = rt_i.render_cycle().copy()
    assert np.array_equal(rt_f.get_hit_records_f64(), rt_i.get_hit_records_f64())
    assert np.array_equal(a1, b1)
    for rt in (rt_f, rt_i):
        rt.set_uint("kernel", 2)
    orc = make_oracle(counts, 128, 128, scale=scale, radius_scale=rs, **kw)
    compare(rt_i, orc, allow_mismatch=3)
    rt_f.close(); rt_i.close()


def test_rotated_body_offcentre_camera_narrow_fov():
    elev, _ = synth_elevation(1440, 720, seed=5)
    # libration-like rotation of the body, camera panned to the limb, narrow field
    a, b = math.radians(7.0), math.radians(-5.0)
    Rz = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]])
    Rm = Rz @ Rx
    kw = dict(u=tuple(Rm[:, 2]), v=tuple(-Rm[:, 1]), eye=(20.0, -298.0, 30.0), target=(6.5, 0.0, 6.9), fov=0.6,
              light_pos=sun_at_phase(60.0, bright_limb_deg=-70.0))
    rt = make_gpu(elev, 128, 96, **kw)
    orc = make_oracle(elev, 128, 96, **kw)
    m = compare(rt, orc, allow_mismatch=3)
    assert m["hits"] > 2000
    rt.close()


def test_polar_view():
    """Camera over the north pole: every ray crosses the converging longitude walls."""
    elev, _ = synth_elevation(720, 360, seed=6)
    kw = dict(eye=(0.0, 0.0, 300.0), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=1.0,
              light_pos=(21460.0 * math.cos(0.2), 0.0, 21460.0 * math.sin(0.2)))
    rt = make_gpu(elev, 96, 96, **kw)
    orc = make_oracle(elev, 96, 96, **kw)
    compare(rt, orc, allow_mismatch=3)
    assert rt.counters()["overflow"] == 0
    rt.close()


@pytest.mark.parametrize("pole", [+1, -1])
def test_polar_cap_rows_close_up(pole):
    """Close-up of a pole with the sun on the horizon: most pixels land in the polar-cap rows (the cells that run on
    to the pole with the row coordinate clamped and no wall on that side) and every shadow ray grazes over them.  The
    filtered kernel decides those cells itself (they used to be handed to the float64 referee one by one - a ray over
    the pole crosses thousands of them); the result must match the oracle like anywhere else."""
    elev, _ = synth_elevation(720, 360, seed=6)
    z = 12.0 * pole
    kw = dict(eye=(0.0, 0.0, z), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=8.0,
              light_pos=(21460.0 * math.cos(0.02), 0.0, pole * 21460.0 * math.sin(0.02)))
    rt = make_gpu(elev, 96, 96, **kw)
    orc = make_oracle(elev, 96, 96, **kw)
    rt.defer_stats(reset=True)
    m = compare(rt, orc, allow_mismatch=3)
    lat = m["oracle"]["hit64"][..., 3]
    hit = m["oracle"]["hit64"][..., 0] > 0
    cap = hit & (np.abs(lat) > math.radians(90.0 - 1.5 * 180.0 / 360))
    assert int(cap.sum()) > 500, int(cap.sum())                 # the view is inside the cap rows
    d = rt.defer_stats()
    assert d["deferred_samples"] <= 0.02 * 96 * 96, d
    assert rt.counters()["overflow"] == 0
    rt.close()


def test_config2_1080p_ds16_terminator():
    """BASELINE config 2: 1920x1080, 5760x2880 float32 map (a LOLA-shaped synthetic map block-meaned on the
    GPU), sun on the terminator, 1 spp primary + shadow; oracle on every 12th pixel."""
    import ctypes as C
    from moonrtx_b200 import _lib
    from moonrtx_b200.device import get_device
    from moonrtx_b200.data_loader import downscale_elevation_dev
    dev = get_device()
    W, H, ds = 23040, 11520, 4
    src = dev.alloc(W * H * 2)
    _lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, src.ptr, W, H, 20240314))
    out, rs = downscale_elevation_dev(src, W, H, ds)
    elev = out.download((H // ds, W // ds), np.float32)
    src.free(); out.free()
    kw = dict(light_pos=sun_at_phase(90.0))
    rt = make_gpu(elev, 1920, 1080, **kw)
    orc = make_oracle(elev, 1920, 1080, **kw)
    m = compare(rt, orc, stride=12, allow_mismatch=4)
    assert m["hits"] > 3000
    c = rt.counters()
    assert c["overflow"] == 0
    print({k: v for k, v in m.items() if k not in ("img", "oracle")}, c)
    rt.close()


def test_jittered_multisample_matches_oracle():
    elev, _ = synth_elevation(720, 360, seed=8)
    kw = dict(light_pos=sun_at_phase(85.0))
    rt = make_gpu(elev, 64, 64, debug_hits=False, **kw)
    rt.set_param(max_accumulation_frames=8)
    img = rt.render_cycle().copy()
    acc = rt.get_accum_buffer()
    orc = make_oracle(elev, 64, 64, jitter=True, **kw)
    o = orc.render(nsamples=8)
    assert np.all(acc[..., 3] == 8.0)
    mae, psnr = image_metrics(img, orc.tonemap(o["accum"]))
    assert mae <= 1.0 and psnr >= 35.0, (mae, psnr)     # a handful of grazing samples may flip
    rt.close()


def test_hit_buffer_and_get_hit_at():
    elev, _ = synth_elevation(720, 360, seed=3)
    rt = make_gpu(elev, 96, 96)
    rt.render_cycle()
    hb = rt.get_hit_buffer()
    orc = make_oracle(elev, 96, 96)
    o = orc.render()
    hit = o["hit32"][..., 3] > 0
    assert np.array_equal(hb[..., 3] > 0, hit)
    assert np.allclose(hb[hit], o["hit32"][hit], atol=2e-5)
    hx, hy, hz, hd = rt._get_hit_at(48, 48)
    assert hd > 0 and 0.9 * R <= math.sqrt(hx * hx + hy * hy + hz * hz) <= 1.15 * R    # renderer_navigation.py:474-476
    assert rt._get_hit_at(0, 0)[3] <= 0
    rt.close()


def test_overlay_blend_and_callbacks():
    elev = np.ones((45, 90), dtype=np.float32)
    rt = make_gpu(elev, 64, 48, light_pos=sun_at_phase(0.0), shadows=False)
    base = rt.render_cycle().copy()
    ov = np.zeros((48, 64, 4), dtype=np.uint8)
    ov[10:20, 10:20] = (0, 0, 0, 255)            # opaque black patch renders 0
    ov[24:40, 24:40] = (0, 0, 0, 128)            # ~50 % black
    ov[0:4, 0:4] = (255, 255, 255, 255)
    rt.set_texture_2d("frame_overlay", ov, filter_mode="Nearest", refresh=False)
    rt.add_postproc("Overlay")
    seen = []
    rt.set_accum_done_cb(lambda r: seen.append(r._frames_rendered))
    img = rt.render_cycle().copy()
    assert seen == [2]
    assert np.all(img[10:20, 10:20, :3] == 0)
    assert np.all(img[0:4, 0:4, :3] == 255)
    expect = (base[24:40, 24:40, :3].astype(np.int64) * 127 + 127) // 255
    assert np.array_equal(img[24:40, 24:40, :3], expect.astype(np.uint8))
    assert np.array_equal(img[44:, 50:], base[44:, 50:])
    # render thread contract: start() renders, callback runs with the padlock held (re-entrant)
    done = []
    def cb(r):
        with r._padlock:
            done.append(1)
    rt.set_accum_done_cb(cb)
    rt.start()
    import time
    t0 = time.time()
    while not done and time.time() - t0 < 10:
        time.sleep(0.01)
    rt.close()
    assert done


def test_state_errors():
    from moonrtx_b200.optix import B200OptiX
    from moonrtx_b200._lib import MoonB200Error
    rt = B200OptiX(width=32, height=32)
    with pytest.raises(MoonB200Error):
        rt.render_cycle()                          # no displacement map yet
    with pytest.raises(ValueError):
        rt.set_texture_2d("moon_color", np.zeros((4, 4, 3), dtype=np.uint8))
    with pytest.raises(ValueError):
        rt.setup_camera("cam1", cam_type="ThinLens", eye=[0, -300, 0], target=[0, 0, 0], up=[0, 0, 1], fov=4)
    rt.close()


def test_persistent_kernel_equals_per_pixel_kernel():
    """The warp-compacting production kernel and the one-thread-per-pixel kernel trace the same rays
    in the same per-pixel order: accumulators, hit buffers and counters must be identical."""
    elev, _ = synth_elevation(1440, 720, seed=11)
    kw = dict(light_pos=sun_at_phase(75.0))
    outs = []
    for kernel in (0, 1):
        rt = make_gpu(elev, 200, 150, debug_hits=True, **kw)
        rt.set_uint("kernel", kernel)
        rt.set_param(max_accumulation_frames=4, min_accumulation_step=4)
        rt.counters(reset=True)
        img = rt.render_cycle().copy()
        outs.append((img, rt.get_accum_buffer(), rt.get_hit_buffer(), rt.get_hit_records_f64(), rt.counters()))
        rt.close()
    a, b = outs
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[2], b[2])
    # the two kernels are separate instantiations: the compiler may contract a*b+c differently,
    # so float64 records agree to rounding, not bit for bit
    assert np.allclose(a[3], b[3], rtol=0, atol=1e-9)
    for k in ("primary_rays", "primary_hits", "shadow_rays", "shadow_occluded"):
        assert a[4][k] == b[4][k], k


def test_filtered_kernel_against_exact_kernel_and_defer_stats():
    """Production path (kernel 2: float32 filter + float64 referee for what it defers) against the float64
    kernel on the same frame: same hit set, same image to rounding, and the deferral statistics add up."""
    elev, _ = synth_elevation(2880, 1440, seed=12)
    kw = dict(light_pos=sun_at_phase(88.0))
    outs = {}
    for kernel in (1, 2):
        rt = make_gpu(elev, 320, 240, **kw)
        rt.set_uint("kernel", kernel)
        rt.counters(reset=True); rt.defer_stats(reset=True)
        img = rt.render_cycle().copy()
        outs[kernel] = (img, rt.get_hit_records_f64(), rt.counters(), rt.defer_stats())
        rt.close()
    (ia, ha, ca, da), (ib, hb, cb, db) = outs[1], outs[2]
    assert da["deferred_samples"] == 0
    assert db["deferred_samples"] == sum(db["primary_reasons"].values()) + sum(db["shadow_reasons"].values())
    assert db["deferred_samples"] <= 0.002 * 320 * 240
    hit_a, hit_b = ha[..., 0] > 0, hb[..., 0] > 0
    assert int((hit_a != hit_b).sum()) <= 2
    both = hit_a & hit_b
    texel = 2.0 * math.pi * R / 2880
    ds_ = np.abs(ha[..., 0] - hb[..., 0])[both] / texel
    assert int((ds_ > 1e-3).sum()) <= 2, float(ds_.max())
    for k in ("primary_rays", "primary_in_sphere"):
        assert ca[k] == cb[k], k
    for k in ("primary_hits", "shadow_rays", "shadow_occluded"):
        assert abs(ca[k] - cb[k]) <= 3, (k, ca[k], cb[k])
    d = np.abs(ia[..., :3].astype(np.int32) - ib[..., :3].astype(np.int32))
    assert d.mean() <= 0.01 and int((d.max(axis=2) > 1).sum()) <= 4


@pytest.mark.parametrize("spp", [1, 6, 40])
def test_wavefront_pipeline_equals_filtered_kernel(spp):
    """Kernel 3 (dense ray generation / streaming walk / dense shading over ray records) traces the same samples with
    the same arithmetic as kernel 2: same rays, same decisions, same deferrals; the per-pixel sum is taken in sample
    order instead of lane order, so accumulators agree to float32 rounding."""
    elev, _ = synth_elevation(2880, 1440, seed=12)
    kw = dict(light_pos=sun_at_phase(88.0))
    outs = {}
    for kernel in (2, 3):
        rt = make_gpu(elev, 320, 240, debug_hits=(spp == 1), **kw)
        rt.set_uint("kernel", kernel)
        # kernel 2 in its in-kernel form (shadow rays inside trace_kernel_fast, no ceiling test, no beam pre-pass): the
        # form whose node counts and float sums kernel 3 reproduces
        rt.set_uint("shadow_queue", 0); rt.set_uint("ceiling", 0); rt.set_uint("beam", 0)
        if spp > 1:
            rt.set_param(max_accumulation_frames=spp, min_accumulation_step=spp)
        rt.counters(reset=True); rt.defer_stats(reset=True)
        img = rt.render_cycle().copy()
        outs[kernel] = (img, rt.get_accum_buffer().copy(), rt.counters(), rt.defer_stats(),
                        rt.get_hit_records_f64().copy() if spp == 1 else None)
        rt.close()
    (ia, aa, ca, da, ha), (ib, ab, cb, db, hb) = outs[2], outs[3]
    for k in ("primary_rays", "primary_in_sphere", "primary_hits", "shadow_rays", "shadow_occluded", "node_visits", "patch_tests"):
        assert ca[k] == cb[k], (k, ca[k], cb[k])
    assert da == db
    assert np.array_equal(aa[..., 3], ab[..., 3]) and np.all(ab[..., 3] == float(spp))
    assert np.allclose(aa[..., :3], ab[..., :3], rtol=2e-6, atol=1e-6)
    d = np.abs(ia[..., :3].astype(np.int32) - ib[..., :3].astype(np.int32))
    assert int(d.max()) <= 1 and int((d.max(axis=2) > 0).sum()) <= 4
    if spp == 1:
        # (the two kernels are compiled in different translation units: the same source expression may be contracted
        #  differently, so records agree to a few ulp rather than bit for bit)
        assert np.array_equal(ha[..., 0] > 0, hb[..., 0] > 0)
        assert np.allclose(ha, hb, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("sq_mode", [2, 3, 4])
@pytest.mark.parametrize("spp,phase,shadows", [(1, 88.0, 1), (16, 90.0, 1), (24, 75.0, 1), (40, 60.0, 1), (6, 90.0, 0)])
def test_hit_queue_and_shade_kernel_equal_in_kernel_shading(spp, phase, shadows, sq_mode):
    """The production cut of kernel 2 (shadow_queue = 2: trace_kernel_fast stops at the primary hit, which goes through the
    hit queue to shade_kernel; that pushes the shadow ray; shadow_queue = 3: the same with trace_kernel_pool, whose warps
    park undecided rays in a pool and fill their hit-queue slots later; shadow_queue = 4, the default: shadow rays too go
    through a batched kernel with a straggler pool, shadow_kernel_pool) against shading inside trace_kernel_fast
    (shadow_queue = 1): the same rays, the same records, sums in fixed point either way - the frames are equal bit for bit.
    Sample counts that are no power of two leave empty slots in the hit queue (24 = 16 + 8 lanes of a second round; 40 = 32 + 8)."""
    elev, _ = synth_elevation(2880, 1440, seed=12)
    outs = []
    for sq in (1, sq_mode):
        rt = make_gpu(elev, 320, 240, debug_hits=(spp == 1), light_pos=sun_at_phase(phase))
        rt.set_uint("shadow_queue", sq); rt.set_uint("shadows", shadows)
        if spp > 1:
            rt.set_param(max_accumulation_frames=spp, min_accumulation_step=spp)
        rt.counters(reset=True); rt.defer_stats(reset=True)
        img = rt.render_cycle().copy()
        outs.append((img, rt.get_accum_buffer().copy(), rt.counters(), rt.defer_stats(), rt.get_hit_buffer().copy(),
                     rt.get_hit_records_f64().copy() if spp == 1 else None))
        rt.close()
    (ia, aa, ca, da, hia, ha), (ib, ab, cb, db, hib, hb) = outs
    assert ca == cb, (ca, cb)
    assert da == db
    assert ca["primary_hits"] > 10000 and (ca["shadow_rays"] > 5000 or not shadows)
    if shadows:
        assert np.array_equal(aa, ab)
    else:
        # (no shadow ray: shade_kernel adds the radiance in fixed point, the in-kernel form in float32 lane order)
        assert np.array_equal(aa[..., 3], ab[..., 3])
        assert np.allclose(aa[..., :3], ab[..., :3], rtol=1e-5, atol=1e-6)
    assert np.array_equal(hia, hib)
    d = np.abs(ia[..., :3].astype(np.int32) - ib[..., :3].astype(np.int32))
    assert int(d.max()) <= (0 if shadows else 1)
    if spp == 1:
        # (written by different kernels: the same source expression may be contracted differently)
        assert np.array_equal(ha[..., 0] > 0, hb[..., 0] > 0) and np.allclose(ha, hb, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("sq_mode", [2, 3, 4])
def test_deferred_list_overflow_keeps_every_sample(sq_mode):
    """With long_walk of two steps nearly every ray goes to the referee; 8 samples per pixel in ONE launch make more
    per-sample entries than the deferred list holds (one per pixel of the frame).  What does not fit waits in the pixel's
    mask word and the referee scans those: the frame equals the one made by 8 launches of one sample each, where the list
    cannot fill up (the fixed-point sums are folded into the float accumulators once per launch, so the two agree to
    float32 rounding; one lost sample of 8 would show as 12 % of a pixel)."""
    from moonrtx_b200 import _lib
    elev, _ = synth_elevation(1440, 720, seed=11)
    W, H, spp = 200, 120, 8
    outs = []
    for split in (False, True):
        rt = make_gpu(elev, W, H, debug_hits=False, light_pos=sun_at_phase(90.0))
        rt.set_uint("shadow_queue", sq_mode); rt.set_uint("long_walk", 2)
        rt._lib.mrtx_set_uint(rt._ctx, b"jitter", 1, 0)
        rt.defer_stats(reset=True)
        if split:
            for k in range(spp):
                _lib.check(rt._lib.mrtx_render(rt._ctx, 0, 0, W, H, k, 1, 1 if k == 0 else 0))
        else:
            _lib.check(rt._lib.mrtx_render(rt._ctx, 0, 0, W, H, 0, spp, 1))
        outs.append((rt.get_accum_buffer().copy(), rt.defer_stats()["deferred_samples"]))
        rt.close()
    (a, na), (b, nb) = outs
    assert na == nb and na > W * H, (na, nb, W * H)
    assert np.array_equal(a[..., 3], b[..., 3]) and float(a[..., :3].max()) > 0.1
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6), float(np.abs(a - b).max())


@pytest.mark.parametrize("spp,phase", [(1, 88.0), (16, 90.0), (40, 60.0)])
def test_shadow_queue_ceiling_and_beam_change_no_decision(spp, phase):
    """The production form of kernel 2 - shadow rays streamed through the queue kernel with lane refill, the ceiling test
    that ends rising rays early, the beam pre-pass that starts the samples of a pixel near the surface - against its
    plain in-kernel form: the same rays with the same hit / miss / occluded decisions; radiance is summed in 2^-36 fixed
    point instead of float32 lane order."""
    elev, _ = synth_elevation(2880, 1440, seed=12)
    kw = dict(light_pos=sun_at_phase(phase))
    outs = []
    for sq, ceil, beam in ((0, 0, 0), (1, 2, 1)):
        rt = make_gpu(elev, 320, 240, debug_hits=(spp == 1), **kw)
        rt.set_uint("shadow_queue", sq); rt.set_uint("ceiling", ceil); rt.set_uint("beam", beam, 2)
        if spp > 1:
            rt.set_param(max_accumulation_frames=spp, min_accumulation_step=spp)
        rt.counters(reset=True); rt.defer_stats(reset=True)
        img = rt.render_cycle().copy()
        outs.append((img, rt.get_accum_buffer().copy(), rt.counters(), rt.defer_stats(),
                     rt.get_hit_records_f64().copy() if spp == 1 else None))
        rt.close()
    (ia, aa, ca, da, ha), (ib, ab, cb, db, hb) = outs
    for k in ("primary_rays", "primary_in_sphere"):
        assert ca[k] == cb[k], (k, ca[k], cb[k])
    # (a sample whose ray starts elsewhere may be deferred where the other form decided it, and the other way round:
    #  the referee then decides it with the same exact test)
    for k in ("primary_hits", "shadow_rays", "shadow_occluded"):
        assert abs(ca[k] - cb[k]) <= 2, (k, ca[k], cb[k])
    assert cb["node_visits"] < ca["node_visits"]
    assert np.array_equal(aa[..., 3], ab[..., 3]) and np.all(ab[..., 3] == float(spp))
    bad = ~np.isclose(aa[..., :3], ab[..., :3], rtol=1e-5, atol=1e-6).all(axis=2)
    assert int(bad.sum()) <= 2, int(bad.sum())
    d = np.abs(ia[..., :3].astype(np.int32) - ib[..., :3].astype(np.int32))
    assert int((d.max(axis=2) > 1).sum()) <= 2
    if spp == 1:
        assert int(((ha[..., 0] > 0) != (hb[..., 0] > 0)).sum()) == 0
        both = ha[..., 0] > 0
        assert np.abs(ha[..., 0] - hb[..., 0])[both].max() <= 1e-9 * 300.0


@pytest.mark.parametrize("kernel", [2, 3])
def test_long_walks_and_budgeted_referee_pieces_change_nothing(kernel):
    """The referee's machinery for long chains, forced onto ordinary rays: with long_walk = 2 nearly every ray that
    enters the shell is handed to the referee (reason 15), and with referee_budget = 8 its lanes give up after a few
    cells, so pieces are re-cut again and again (nearest-hit ordering for primary rays, side-by-side intervals for
    shadow rays).  Same hits, same image as the default path."""
    elev, _ = synth_elevation(1440, 720, seed=21)
    kw = dict(light_pos=sun_at_phase(87.0))
    outs = []
    for forced in (False, True):
        rt = make_gpu(elev, 160, 120, **kw)
        rt.set_uint("kernel", kernel)
        if forced:
            rt.set_uint("long_walk", 2)
            rt.set_uint("referee_budget", 8)
        rt.defer_stats(reset=True); rt.counters(reset=True)
        img = rt.render_cycle().copy()
        outs.append((img, rt.get_hit_records_f64().copy(), rt.counters(), rt.defer_stats()))
        rt.close()
    (ia, ha, ca, da), (ib, hb, cb, db) = outs
    n15 = db["primary_reasons"].get(15, 0) + db["shadow_reasons"].get(15, 0)
    assert n15 > 0.5 * ca["primary_in_sphere"], (n15, ca["primary_in_sphere"])
    hit_a, hit_b = ha[..., 0] > 0, hb[..., 0] > 0
    assert int((hit_a != hit_b).sum()) <= 2
    both = hit_a & hit_b
    texel = 2.0 * math.pi * R / 1440
    assert int((np.abs(ha[..., 0] - hb[..., 0])[both] / texel > 1e-3).sum()) <= 2
    for k in ("primary_rays", "primary_in_sphere"):
        assert ca[k] == cb[k], k
    for k in ("primary_hits", "shadow_rays", "shadow_occluded"):
        assert abs(ca[k] - cb[k]) <= 3, (k, ca[k], cb[k])
    d = np.abs(ia[..., :3].astype(np.int32) - ib[..., :3].astype(np.int32))
    assert d.mean() <= 0.01 and int((d.max(axis=2) > 1).sum()) <= 4


def test_reference_accumulation_parameters_render_one_batched_cycle():
    """The reference asks for min_accumulation_step=1, max_accumulation_frames=64 (moon_renderer.py:578).  The
    drop-in renders the cycle's samples together (same samples, keyed on (pixel, index)); the frame is complete
    when render_cycle returns and the callbacks have fired once."""
    elev, _ = synth_elevation(720, 360, seed=4)
    kw = dict(light_pos=sun_at_phase(80.0))
    fired = []
    outs = []
    for step in (1, 24):
        rt = make_gpu(elev, 64, 48, debug_hits=False, **kw)
        rt.set_param(min_accumulation_step=step, max_accumulation_frames=24)
        rt.set_accum_done_cb(lambda r: fired.append(step))
        rt.render_cycle()
        outs.append(rt.get_accum_buffer().copy())
        assert rt.counters()["primary_rays"] == 64 * 48 * 24
        rt.close()
    assert fired == [1, 24]
    assert np.all(outs[0][..., 3] == 24.0) and np.array_equal(outs[0], outs[1])


def test_more_than_32_samples_are_chunked():
    """The filtered kernel takes <= 32 samples per launch (one mask bit each in the deferred list): 40 spp must
    equal 32 + 8 spp accumulated in two calls, and match the oracle."""
    elev, _ = synth_elevation(720, 360, seed=8)
    kw = dict(light_pos=sun_at_phase(85.0))
    rt = make_gpu(elev, 48, 48, debug_hits=False, **kw)
    rt.set_param(max_accumulation_frames=40, min_accumulation_step=40)
    img = rt.render_cycle().copy()
    acc = rt.get_accum_buffer().copy()
    assert np.all(acc[..., 3] == 40.0)
    orc = make_oracle(elev, 48, 48, jitter=True, **kw)
    o = orc.render(nsamples=40)
    mae, psnr = image_metrics(img, orc.tonemap(o["accum"]))
    assert mae <= 1.0 and psnr >= 35.0, (mae, psnr)
    # all but a handful of (grazing) samples agree: the per-pixel sums differ in few pixels
    assert float(np.mean(np.abs(acc[..., :3] - o["accum"][..., :3]).max(axis=2) > 1e-3 * 40)) <= 0.02
    rt.close()


def test_fine_cells_polar_and_limb_rays_match_oracle():
    """Cells as small against float32 as at full LOLA resolution (46080 x 23040: 1.4e-4 R), whole disk at 4K so the
    sample contains limb rays and rays that graze the poles; oracle on a sparse pixel grid."""
    import ctypes as C
    from moonrtx_b200 import _lib
    from moonrtx_b200.device import get_device
    dev = get_device()
    W, H = 46080, 23040
    src = dev.alloc(W * H * 2)
    _lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, src.ptr, W, H, 20240314))
    counts = src.download((H, W), np.int16)
    src.free()
    scale = float(np.float32(0.5 / 1737400.0))
    m = np.float32(counts.max())
    rs = float(np.float32(np.float32(m * np.float32(scale)) + np.float32(1)))
    kw = dict(light_pos=sun_at_phase(90.0))
    rt = make_gpu(counts, 3840, 2160, scale=scale, radius_scale=rs, **kw)
    orc = make_oracle(counts, 3840, 2160, scale=scale, radius_scale=rs, **kw)
    m = compare(rt, orc, stride=40, allow_mismatch=3)
    assert m["hits"] > 1500
    d = rt.defer_stats()
    assert rt.counters()["overflow"] == 0
    print({k: v for k, v in m.items() if k not in ("img", "oracle")}, d)
    rt.close()


def test_star_map_background_and_sun_disk_on_miss_rays():
    """what the 64 % of a whole-disk frame that misses the Moon shows.  The reference's own calls
    (moon_renderer.py:602-609, 643-650, 855): set_background_mode("TextureEnvironment"), set_background(star map,
    gamma, "UByte4"), a flat-shaded "sun_disk" particle placed by update_data.  The oracle reads the same 8-bit environment
    texture and the same sphere: every pixel of the frame is compared, at 1 spp and with jittered samples."""
    elev, _ = synth_elevation(720, 360, seed=6)
    rng = np.random.default_rng(5)
    stars = np.zeros((180, 360, 3), np.float32)
    stars += rng.uniform(0.0, 0.04, stars.shape).astype(np.float32)
    ys, xs = rng.integers(0, 180, 400), rng.integers(0, 360, 400)
    stars[ys, xs] = rng.uniform(0.3, 1.0, (400, 3)).astype(np.float32)
    # the Sun disk a little off the limb, as calculate_sun_disk places it near new moon (moon_renderer.py:775-777)
    disk_pos, disk_r = [14.0 * 3100 / 300, -300.0 + 3100.0, 6.0 * 3100 / 300], 3100 * math.tan(math.radians(0.45))
    kw = dict(light_pos=sun_at_phase(120.0))
    for spp in (1, 8):
        rt = make_gpu(elev, 192, 128, debug_hits=False, **kw)
        rt.set_background_mode("TextureEnvironment")
        rt.set_background(stars, gamma=2.2, rt_format="UByte4")
        rt.setup_material("flat", {"dummy": 1})
        rt.set_data("sun_disk", geom="ParticleSet", mat="flat", pos=[[0.0, 3100.0, 0.0]], r=0.01, c=2.0)
        rt.update_data("sun_disk", pos=[disk_pos], r=disk_r)
        if spp > 1:
            rt.set_param(max_accumulation_frames=spp, min_accumulation_step=spp)
        img = rt.render_cycle().copy()
        env = rt.get_background_texture()
        # the texture itself: round(255 v^gamma), data as PlotOptiX's UByte4 + gamma stores it
        want = np.rint(255.0 * np.clip(stars.astype(np.float64), 0, 1) ** 2.2).astype(np.uint8)
        assert env.shape == (180, 360, 4) and np.array_equal(env[..., :3], want) and np.all(env[..., 3] == 255)
        orc = make_oracle(elev, 192, 128, jitter=spp > 1, background=env, sun_disk=(disk_pos, disk_r, 2.0), **kw)
        o = orc.render(nsamples=spp)
        ref = orc.tonemap(o["accum"])
        miss = o["hit64"][..., 0] <= 0
        assert miss.mean() > 0.5
        d = np.abs(img[..., :3].astype(np.int32) - ref[..., :3].astype(np.int32))
        assert int(d[miss].max()) <= 1, int(d[miss].max())                  # background pixels: the same lookup
        mae, psnr = image_metrics(img, ref)
        assert mae <= 1.0 and psnr >= 40.0, (mae, psnr)
        # the disk is there (pure white where the flat colour 2.0 saturates) and stars are visible
        assert int((img[..., :3].min(axis=2) == 255).sum()) > 30
        assert int((img[miss][:, :3].max(axis=1) > 60).sum()) > 20
        # removing both gives the black background back
        rt.set_background(0)
        rt.delete_geometry("sun_disk")
        img0 = rt.render_cycle()
        if spp == 1:                                  # (with jitter a limb pixel whose last sample missed is still lit)
            assert int(img0[miss][:, :3].max()) == 0
        rt.close()
