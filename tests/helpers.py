"""Build the same scene for the CUDA drop-in (B200OptiX) and the float64 oracle."""
import math

import numpy as np

DEFAULTS = dict(
    u=(0.0, 0.0, 1.0), v=(0.0, -1.0, 0.0), pos=(0.0, 0.0, 0.0), radius=10.0,
    eye=(0.0, -300.0, 0.0), target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), fov=4.242192793,
    light_pos=(21460.0, 0.0, 0.0), light_radius=100.0, light_radiance=80.0 * (2146.0 / 100.0) ** 2,
    scene_epsilon=1.0e-4, shadows=True, exposure=0.9, gamma=2.2,
    # direct light (camera segment + light segment): the north_star path.  The reference asks for (2, 4) - two diffuse
    # interreflection bounces (moon_renderer.py:583) - which tests/test_bounce_gpu.py covers
    path_seg_range=(2, 2),
)


def make_oracle(elevation, img_w, img_h, texture=None, jitter=False, scale=None, radius_scale=None, **kw):
    from oracle.render_oracle import OracleScene
    p = {**DEFAULTS, **kw}
    return OracleScene(elevation, scale=scale, radius_scale=radius_scale, img_w=img_w, img_h=img_h,
                       texture=texture, jitter=jitter, **p)


def make_gpu(elevation, img_w, img_h, texture=None, scale=None, radius_scale=None, debug_hits=True, **kw):
    """The reference's own call sequence (moon_renderer.py:570-641) on the drop-in."""
    from moonrtx_b200.optix import B200OptiX
    p = {**DEFAULTS, **kw}
    rt = B200OptiX(width=img_w, height=img_h)
    rt.set_param(min_accumulation_step=1, max_accumulation_frames=1)
    rt.set_uint("path_seg_range", *p["path_seg_range"])
    rt.set_float("scene_epsilon", p["scene_epsilon"])
    rt.set_float("marching_step", 5.0e-3)
    rt.set_float("marching_step_eps", 3.0e-4)
    rt.set_ambient(0)
    rt.set_float("tonemap_exposure", p["exposure"])
    rt.set_float("tonemap_gamma", p["gamma"])
    rt.add_postproc("Gamma")
    rt.set_background(0)
    if texture is not None:
        rt.set_texture_2d("moon_color", texture)
    rt.update_material("diffuse", {"ColorTextures": ["moon_color"]})
    rt.set_data("moon", geom="ParticleSetTextured", geom_attr="DisplacedSurface",
                pos=list(p["pos"]), u=list(p["u"]), v=list(p["v"]), r=p["radius"])
    if isinstance(elevation, tuple) or elevation.dtype == np.int16:       # (DeviceBuffer, W, H): the map already in HBM
        rt.set_displacement_i16("moon", elevation, radius_scale=radius_scale, scale=scale)
    else:
        rt.set_displacement("moon", elevation, refresh=False)
    rt.setup_camera("cam1", cam_type="Pinhole", eye=list(p["eye"]), target=list(p["target"]), up=list(p["up"]),
                    fov=p["fov"], aperture_radius=0.01, aperture_fract=0.2, focal_scale=0.7)
    rt.setup_light("sun", color=p["light_radiance"], radius=p["light_radius"], in_geometry=False)
    rt.update_light("sun", pos=list(p["light_pos"]))
    rt.set_uint("shadows", 1 if p["shadows"] else 0)
    if debug_hits:
        rt.set_uint("debug_hits", 1)
    return rt


def image_metrics(a, b):
    a = a[..., :3].astype(np.float64)
    b = b[..., :3].astype(np.float64)
    mae = float(np.abs(a - b).mean())
    mse = float(((a - b) ** 2).mean())
    psnr = float("inf") if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)
    return mae, psnr


def sun_at_phase(phase_deg, bright_limb_deg=-90.0, dist=21460.0):
    """moon_renderer.py:723-725"""
    b, p = math.radians(bright_limb_deg), math.radians(phase_deg)
    return (-math.sin(b) * math.sin(p) * dist, -math.cos(p) * dist, math.cos(b) * math.sin(p) * dist)


def penetration_texels(orc, x, y, s_root, texel, half_window_texels=2.0, n=81):
    """How far below the oracle's surface the pixel-centre ray of (x, y) gets around parameter s_root, in texels.  A pixel on
    which the GPU and the oracle disagree is a grazing case - a crest the ray touches: which side of a tangent root a ray
    falls on is decided by the last bits of either implementation - only if this is below the tolerance."""
    import math
    s = orc.s
    sx = ((x + 0.5) / s.img_w * 2.0 - 1.0) * s.tan_half_fov * s.img_w / s.img_h
    sy = (1.0 - (y + 0.5) / s.img_h * 2.0) * s.tan_half_fov
    d = np.array(s.w) + sx * np.array(s.right) + sy * np.array(s.up)
    d /= np.linalg.norm(d)
    Rm = np.array([list(s.ex), list(s.ey), list(s.ez)])
    o_b, d_b = Rm @ (np.array(s.eye) - np.array(s.pos)), Rm @ d
    worst = 0.0
    for t in np.linspace(s_root - half_window_texels * texel, s_root + half_window_texels * texel, n):
        p = o_b + t * d_b
        r = np.linalg.norm(p)
        lat, lon = math.degrees(math.asin(p[2] / r)), math.degrees(math.atan2(p[0], -p[1]))
        worst = min(worst, r - s.radius * orc.displacement(lat, lon))
    return -worst / texel
