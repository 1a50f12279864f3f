"""
TEST INFRASTRUCTURE - ctypes wrapper of oracle/render_oracle.c (float64 CPU oracle of the
render path; see the header of that file: PARITY UNPINNED against the closed PlotOptiX
engine, pinned on the reference's own convention fixtures and analytic answers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "render_oracle.c")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "librender_oracle.so")

_D3 = C.c_double * 3


class _Scene(C.Structure):
    _fields_ = [
        ("W", C.c_int), ("H", C.c_int),
        ("map_f32", C.c_void_p), ("map_i16", C.c_void_p),
        ("scale", C.c_float), ("radius_scale", C.c_float),
        ("dmax", C.c_double),
        ("ex", _D3), ("ey", _D3), ("ez", _D3), ("pos", _D3), ("radius", C.c_double),
        ("eye", _D3), ("w", _D3), ("right", _D3), ("up", _D3), ("tan_half_fov", C.c_double),
        ("img_w", C.c_int), ("img_h", C.c_int),
        ("light_pos", _D3), ("light_radius", C.c_double), ("light_radiance", C.c_double),
        ("scene_epsilon", C.c_double),
        ("jitter", C.c_int), ("shadows", C.c_int),
        ("tex", C.c_void_p), ("tex_w", C.c_int), ("tex_h", C.c_int),
        ("exposure", C.c_float), ("inv_gamma", C.c_float),
        ("sun_disk_pos", _D3), ("sun_disk_radius", C.c_double), ("sun_disk_color", _D3),
        ("env", C.c_void_p), ("env_w", C.c_int), ("env_h", C.c_int),
        ("tubes", C.c_void_p), ("n_tubes", C.c_int),
        ("n_bounce", C.c_int),
    ]


def build(force: bool = False) -> str:
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(BUILD, exist_ok=True)
        subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"],
                       check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        l.orc_render.restype = C.c_long
        l.orc_render.argtypes = [C.POINTER(_Scene), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_trace_ray.restype = C.c_int
        l.orc_trace_ray.argtypes = [C.POINTER(_Scene), C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_displacement.restype = C.c_double
        l.orc_displacement.argtypes = [C.POINTER(_Scene), C.c_double, C.c_double]
        l.orc_tonemap.restype = None
        l.orc_tonemap.argtypes = [C.POINTER(_Scene), C.c_void_p, C.c_long, C.c_void_p]
        l.orc_scene_size.restype = C.c_int
        assert l.orc_scene_size() == C.sizeof(_Scene), "orc_scene layout mismatch"
        _lib = l
    return _lib


def _unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


class OracleScene:
    """
    Scene description with the reference's defaults (moon_renderer.py:36-136):
    displacement map (float32, or int16 + scale + radius_scale), body axes u / v as passed
    to rt.set_data / update_data, pinhole camera, sun light.
    """

    def __init__(self, elevation, *, scale=None, radius_scale=None, img_w=64, img_h=64,
                 u=(0, 0, 1), v=(0, -1, 0), pos=(0, 0, 0), radius=10.0,
                 eye=(0, -300, 0), target=(0, 0, 0), up=(0, 0, 1), fov=4.242192793,
                 light_pos=(21460.0, 0.0, 0.0), light_radius=100.0, light_radiance=80.0 * (2146.0 / 100.0) ** 2,
                 scene_epsilon=1.0e-4, jitter=False, shadows=True, texture=None,
                 exposure=0.9, gamma=2.2, background=None, sun_disk=None, tubes=None, path_seg_range=(2, 2)):
        self.s = _Scene()
        s = self.s
        self.elevation = np.ascontiguousarray(elevation)
        s.H, s.W = self.elevation.shape
        if self.elevation.dtype == np.float32:
            s.map_f32 = self.elevation.ctypes.data
            s.map_i16 = None
            s.scale, s.radius_scale = 0.0, 1.0
            s.dmax = float(self.elevation.max())
        elif self.elevation.dtype == np.int16:
            s.map_f32 = None
            s.map_i16 = self.elevation.ctypes.data
            s.scale, s.radius_scale = float(scale), float(radius_scale)
            m = np.float32(self.elevation.max())
            s.dmax = float(np.float32(np.float32(np.float32(m * np.float32(scale)) + np.float32(1)) / np.float32(radius_scale)))
        else:
            raise ValueError("elevation must be float32 or int16")
        ez = _unit(u)
        vv = np.asarray(v, dtype=np.float64)
        vv = _unit(vv - vv.dot(ez) * ez)
        ex = np.cross(ez, vv)
        s.ex, s.ey, s.ez = _D3(*ex), _D3(*(-vv)), _D3(*ez)
        s.pos = _D3(*[float(x) for x in pos])
        s.radius = float(radius)
        eye = np.asarray(eye, dtype=np.float64)
        w = _unit(np.asarray(target, dtype=np.float64) - eye)
        right = _unit(np.cross(w, np.asarray(up, dtype=np.float64)))
        up2 = np.cross(right, w)
        s.eye, s.w, s.right, s.up = _D3(*eye), _D3(*w), _D3(*right), _D3(*up2)
        s.tan_half_fov = math.tan(math.radians(fov) * 0.5)
        s.img_w, s.img_h = int(img_w), int(img_h)
        s.light_pos = _D3(*[float(x) for x in light_pos])
        s.light_radius, s.light_radiance = float(light_radius), float(light_radiance)
        s.scene_epsilon = float(scene_epsilon)
        s.jitter, s.shadows = int(bool(jitter)), int(bool(shadows))
        self.texture = None
        if texture is not None:
            self.texture = np.ascontiguousarray(texture, dtype=np.uint8)
            s.tex = self.texture.ctypes.data
            s.tex_h, s.tex_w = self.texture.shape[:2]
        s.exposure, s.inv_gamma = float(exposure), float(1.0 / gamma)
        # what rays that miss the Moon see: background = RGBA8 environment texture of linear radiance (the star map as
        # moonrtx_b200.data_loader.background_texture prepares it), sun_disk = (centre, radius, colour)
        self.background = None
        if background is not None:
            self.background = np.ascontiguousarray(background, dtype=np.uint8)
            s.env = self.background.ctypes.data
            s.env_h, s.env_w = self.background.shape[:2]
        s.sun_disk_radius = 0.0
        if sun_disk is not None:
            c, r, col = sun_disk
            s.sun_disk_pos = _D3(*[float(x) for x in c])
            s.sun_disk_radius = float(r)
            col = np.broadcast_to(np.asarray(col, dtype=np.float64).reshape(-1), (3,)) if np.size(col) in (1, 3) else np.asarray(col, dtype=np.float64)[:3]
            s.sun_disk_color = _D3(*[float(x) for x in col])
        self.set_tubes(tubes)
        # rt.set_uint("path_seg_range", min, max): camera ray + light ray are two segments, every one beyond is a bounce
        s.n_bounce = max(0, min(int(path_seg_range[1]) - 2, 4))

    def set_tubes(self, tubes):
        """overlay tubes: float32 (n, 12) = a.xyz, r, b.xyz, -, colour.rgb, - in scene space (B200OptiX._tube_segments())"""
        self.tubes = None if tubes is None or len(tubes) == 0 else np.ascontiguousarray(tubes, dtype=np.float32).reshape(-1, 12)
        self.s.tubes = None if self.tubes is None else self.tubes.ctypes.data
        self.s.n_tubes = 0 if self.tubes is None else len(self.tubes)

    def render(self, x0=0, y0=0, x1=None, y1=None, stride=1, sample0=0, nsamples=1):
        s = self.s
        x1 = s.img_w if x1 is None else x1
        y1 = s.img_h if y1 is None else y1
        nx = (x1 - x0 + stride - 1) // stride
        ny = (y1 - y0 + stride - 1) // stride
        accum = np.zeros((ny, nx, 4), dtype=np.float64)
        hit64 = np.zeros((ny, nx, 4), dtype=np.float64)
        hit32 = np.zeros((ny, nx, 4), dtype=np.float32)
        stats = np.zeros((ny, nx, 2), dtype=np.int64)
        n = lib().orc_render(C.byref(s), x0, y0, x1, y1, stride, sample0, nsamples,
                             accum.ctypes.data, hit64.ctypes.data, hit32.ctypes.data, stats.ctypes.data)
        assert n == nx * ny
        return {"accum": accum, "hit64": hit64, "hit32": hit32, "stats": stats}

    def tonemap(self, accum):
        a = np.ascontiguousarray(accum, dtype=np.float64)
        out = np.empty(a.shape[:-1] + (4,), dtype=np.uint8)
        lib().orc_tonemap(C.byref(self.s), a.ctypes.data, a.size // 4, out.ctypes.data)
        return out

    def trace_ray(self, o, d):
        o = np.ascontiguousarray(o, dtype=np.float64)
        d = np.ascontiguousarray(d, dtype=np.float64)
        out = np.zeros(8, dtype=np.float64)
        hit = lib().orc_trace_ray(C.byref(self.s), o.ctypes.data, d.ctypes.data, out.ctypes.data)
        return bool(hit), out

    def displacement(self, lat_deg, lon_deg):
        return lib().orc_displacement(C.byref(self.s), float(lat_deg), float(lon_deg))
