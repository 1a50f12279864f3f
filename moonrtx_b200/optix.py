"""
`B200OptiX`: drop-in for the subset of `plotoptix.TkOptiX` / `NpOptiX` that MoonRTX's hot
path calls (`self.rt` in moonrtx/moon_renderer.py:571-650, 852-871, renderer_video.py,
renderer_navigation.py) - same method names, argument meaning and threading contract
(SURVEY.md §8b).  The scene it accepts is the one MoonRTX builds: one textured, displaced
sphere ("moon"), one spherical light ("sun"), one pinhole camera, Gamma and Overlay
post-processing; beside them what the reference draws around the Moon: the star-map background and the
visible Sun disk on rays that miss it (set_background, set_data("sun_disk")) and the tube graphs of the
selenographic grid, the labels and the pins (set_graph / update_graph / delete_geometry), flat-shaded and
shadowless as the reference's materials make them.

Rendering runs on the B200 through libmoonb200.so; there is no CPU fallback.

Threading contract (as PlotOptiX): `start()` spawns the render thread; every accumulation
cycle renders `max_accumulation_frames` passes of 1 spp, then resolves (Gamma, Overlay),
reads the image back and fires `on_launch_finished(rt)` and the accum-done callback with
`_padlock` (an RLock) held, so a callback may call the setters and `refresh_scene()`.
Headless users can skip the thread and call `render_cycle()` synchronously.
"""

import ctypes as C
import threading
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib
from .device import Device, default_index


class _OptixShim:
    """`rt._optix.get_camera_fov(0)` / `set_camera_fov(fov)` (renderer_navigation.py:262, 521)."""

    def __init__(self, rt: "B200OptiX"):
        self._rt = rt

    def get_camera_fov(self, handle: int = 0) -> float:
        return float(self._rt._cam["fov"])

    def set_camera_fov(self, fov: float) -> None:
        self._rt.update_camera(fov=fov)


def tube_segments(geometry: dict) -> np.ndarray:
    """every visible segment of every graph: float32 (n, 12) = a.xyz, r, b.xyz, 0, colour.rgb, 0 (mrtx_set_tubes)"""
    parts = []
    for g in geometry.values():
        if g.get("geom") != "Graph" or g.get("pos") is None or g.get("edges") is None:
            continue
        pos = np.asarray(g["pos"], dtype=np.float64).reshape(-1, 3)
        edges = np.asarray(g["edges"], dtype=np.int64).reshape(-1, 2)
        if len(pos) == 0 or len(edges) == 0:
            continue
        r = np.asarray(0.05 if g.get("r") is None else g["r"], dtype=np.float64).reshape(-1)
        r = np.full(len(pos), r[0]) if r.size == 1 else r
        if r.size != len(pos):
            raise ValueError("graph radii: one value or one per vertex")
        c = np.asarray(0.94 if g.get("c") is None else g["c"], dtype=np.float64)
        if c.size == 1:
            col = np.full((len(pos), 3), float(c.reshape(-1)[0]))
        elif c.size == 3:
            col = np.broadcast_to(c.reshape(1, 3), (len(pos), 3))
        else:
            col = c.reshape(-1, 3)
        # (a segment between vertices of different radii is drawn with the smaller one: the reference only ever
        #  gives a whole label one radius, or 0 to hide it - renderer_labels.py:128-130)
        rs = np.minimum(r[edges[:, 0]], r[edges[:, 1]])
        keep = rs > 0.0
        if not keep.any():
            continue
        seg = np.zeros((int(keep.sum()), 12), np.float32)
        seg[:, 0:3] = pos[edges[keep, 0]]; seg[:, 3] = rs[keep]
        seg[:, 4:7] = pos[edges[keep, 1]]
        seg[:, 8:11] = col[edges[keep, 0]]
        parts.append(seg)
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, 12), np.float32)


class B200OptiX:
    def __init__(self, width: int = 1920, height: int = 1080,
                 on_launch_finished: Optional[Callable] = None,
                 on_rt_accum_done: Optional[Callable] = None,
                 device: Optional[Device] = None, **_ignored):
        # one C-ABI context (scene + frame buffers + stream) per renderer object
        self._own_device = device is None
        self._dev = device or Device(default_index())
        self._lib = self._dev.lib
        self._ctx = self._dev.ctx
        self._padlock = threading.RLock()
        self._width, self._height = int(width), int(height)
        self._is_started = False
        self._is_closed = False
        self._on_launch_finished = on_launch_finished
        self._accum_done_cb = on_rt_accum_done
        self._optix = _OptixShim(self)
        self._params = {"min_accumulation_step": 1, "max_accumulation_frames": 1}
        self._floats = {}
        self._uints = {}
        self._postproc = []
        self._moon = {"pos": (0.0, 0.0, 0.0), "u": (0.0, 0.0, 1.0), "v": (0.0, -1.0, 0.0), "r": 10.0}
        self._moon_name = None
        self._cam_name = None
        self._cam = {"eye": (0.0, -300.0, 0.0), "target": (0.0, 0.0, 0.0), "up": (0.0, 0.0, 1.0), "fov": 4.242192793}
        self._light = {"pos": (21460.0, 0.0, 0.0), "radius": 100.0, "color": 80.0 * (2146.0 / 100.0) ** 2}
        self._ignored_geometry = {}
        self._pinned = []
        self._img_rgba = self.pinned_empty((self._height, self._width, 4), np.uint8)
        self._img_rgba[...] = 0
        self._dirty = threading.Event()
        self._stop = threading.Event()
        self._thread = None
        self._encoder = None
        self._frames_rendered = 0
        self._frames_submitted = 0
        self.deterministic = None      # None: jitter iff max_accumulation_frames > 1
        _lib.check(self._lib.mrtx_resize(self._ctx, self._width, self._height))
        self._push_camera()
        self._push_light()
        self._push_frame()

    # ---- pinned host memory (frame read-back / texture upload at PCIe speed) ------------------
    def pinned_empty(self, shape, dtype) -> np.ndarray:
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _lib.check(self._lib.mrtx_host_alloc(n, C.byref(p)))
        self._pinned.append(p.value)
        buf = (C.c_uint8 * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def pinned_like(self, a: np.ndarray) -> np.ndarray:
        return self.pinned_empty(a.shape, a.dtype)

    # ---- pushes to the C ABI ---------------------------------------------------------
    def _push_camera(self):
        c = self._cam
        _lib.check(self._lib.mrtx_set_camera(self._ctx, _lib.vec3(c["eye"]), _lib.vec3(c["target"]),
                                             _lib.vec3(c["up"]), float(c["fov"])))

    def _push_light(self):
        l = self._light
        _lib.check(self._lib.mrtx_set_light(self._ctx, _lib.vec3(l["pos"]), float(l["radius"]), float(l["color"])))

    def _push_frame(self):
        m = self._moon
        _lib.check(self._lib.mrtx_set_frame(self._ctx, _lib.vec3(m["pos"]), _lib.vec3(m["u"]), _lib.vec3(m["v"]),
                                            float(m["r"])))

    # ---- parameters (moon_renderer.py:578-600) -----------------------------------------
    def set_param(self, **kwargs):
        with self._padlock:
            for k, v in kwargs.items():
                if k not in ("min_accumulation_step", "max_accumulation_frames", "rt_timeout",
                             "light_shading", "compute_timeout"):
                    raise ValueError(f"unknown parameter {k}")
                self._params[k] = v

    def get_param(self, name):
        return self._params.get(name)

    def set_uint(self, name: str, x: int, y: Optional[int] = None, refresh: bool = False):
        with self._padlock:
            self._uints[name] = (x, y)
            _lib.check(self._lib.mrtx_set_uint(self._ctx, name.encode(), int(x), int(y or 0)))
        if refresh:
            self.refresh_scene()

    def set_float(self, name: str, x: float, y=None, z=None, refresh: bool = False):
        with self._padlock:
            self._floats[name] = x
            _lib.check(self._lib.mrtx_set_float(self._ctx, name.encode(), float(x)))
        if refresh:
            self.refresh_scene()

    def get_float(self, name: str):
        return self._floats.get(name)

    def set_ambient(self, color, refresh: bool = False):
        # MoonRTX renders with zero ambient (moon_renderer.py:595); nothing else is supported
        if np.any(np.asarray(color, dtype=np.float64) != 0):
            raise ValueError("B200OptiX supports ambient 0 only (moon_renderer.py:595)")

    def add_postproc(self, stage: str, refresh: bool = False):
        if stage not in ("Gamma", "Overlay"):
            raise ValueError(f"unsupported post-processing stage {stage}")
        if stage not in self._postproc:
            self._postproc.append(stage)

    # ---- textures / background (moon_renderer.py:602-617, renderer_video.py:137) ----------
    def set_texture_2d(self, name: str, data, addr_mode=None, filter_mode=None, keep_on_host=False,
                       refresh: bool = False, **_):
        a = np.ascontiguousarray(data)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 4:
            raise ValueError("texture must be a (h, w, 4) uint8 array")
        slot = {"moon_color": 0, "frame_overlay": 1}.get(name)
        if slot is None:
            raise ValueError(f"unknown texture {name} (moon_color, frame_overlay)")
        if slot == 1 and a.shape[:2] != (self._height, self._width):
            raise ValueError("frame_overlay must match the frame size")
        with self._padlock:
            _lib.check(self._lib.mrtx_set_texture_rgba8(self._ctx, slot, a.ctypes.data, a.shape[1], a.shape[0]))
        if refresh:
            self.refresh_scene()

    def set_background_mode(self, mode, refresh: bool = False):
        """moon_renderer.py:605: "TextureEnvironment" = the background texture is looked up by ray direction."""
        self._background_mode = mode
        if refresh:
            self.refresh_scene()

    def set_background(self, bg, gamma: float = 1.0, rt_format=None, refresh: bool = False, **_):
        """
        moon_renderer.py:606-609.  A float32 (h, w, 3) array in [0, 1] (the star map of load_starmap) becomes the
        environment texture rays that miss the Moon see: 8-bit linear radiance v^gamma, as PlotOptiX's "UByte4"
        format with `gamma` stores it; a scalar / colour 0 = black.  A uint8 (h, w, 4) array is taken as the finished
        texture.  Its orientation (scene +Z up, -Y at longitude 0) is this engine's choice: PlotOptiX's own
        TextureEnvironment mapping is closed.
        """
        a = np.asarray(bg)
        with self._padlock:
            if a.ndim == 3 and a.shape[2] == 3 and a.dtype != np.uint8:
                a = np.ascontiguousarray(a, dtype=np.float32)
                _lib.check(self._lib.mrtx_set_background_f32(self._ctx, a.ctypes.data, a.shape[1], a.shape[0], float(gamma)))
                self._background = (a.shape[1], a.shape[0])
            elif a.ndim == 3 and a.shape[2] == 4 and a.dtype == np.uint8:
                a = np.ascontiguousarray(a)
                _lib.check(self._lib.mrtx_set_texture_rgba8(self._ctx, 2, a.ctypes.data, a.shape[1], a.shape[0]))
                self._background = (a.shape[1], a.shape[0])
            elif a.size <= 4 and not np.any(a):
                _lib.check(self._lib.mrtx_set_texture_rgba8(self._ctx, 2, None, 0, 0))
                self._background = None
            else:
                raise ValueError("background must be a float (h, w, 3) image, a uint8 (h, w, 4) texture or 0")
        if refresh:
            self.refresh_scene()

    def get_background_texture(self) -> Optional[np.ndarray]:
        """the environment texture as the device holds it (uint8 RGBA, linear radiance); None = black"""
        w, h = C.c_int(), C.c_int()
        with self._padlock:
            _lib.check(self._lib.mrtx_read_background_rgba8(self._ctx, None, C.byref(w), C.byref(h)))
            if w.value == 0:
                return None
            out = np.empty((h.value, w.value, 4), np.uint8)
            _lib.check(self._lib.mrtx_read_background_rgba8(self._ctx, out.ctypes.data, C.byref(w), C.byref(h)))
        return out

    def _push_sun_disk(self):
        g = self._ignored_geometry.get(self._sun_disk_name) if getattr(self, "_sun_disk_name", None) else None
        if g is None or g.get("pos") is None or g.get("r") is None:
            _lib.check(self._lib.mrtx_set_sun_disk(self._ctx, None, 0.0, None))
            return
        pos = np.asarray(g["pos"], dtype=np.float64).reshape(-1)[:3]
        r = float(np.asarray(g["r"], dtype=np.float64).reshape(-1)[0])
        c = g.get("c")
        col = np.asarray(1.0 if c is None else c, dtype=np.float32).reshape(-1)
        col = np.full(3, col[0], np.float32) if col.size == 1 else col[:3].astype(np.float32)
        _lib.check(self._lib.mrtx_set_sun_disk(self._ctx, (C.c_double * 3)(*pos), r, (C.c_float * 3)(*col)))

    def update_material(self, name, data, refresh: bool = False):
        self._material = (name, dict(data))

    def setup_material(self, name, data):
        self._ignored_geometry["material:" + name] = dict(data)

    # ---- geometry (moon_renderer.py:620-624, 648-650, 854-855) ------------------------------
    def set_data(self, name: str, pos=None, r=None, u=None, v=None, geom="ParticleSet", geom_attr=None,
                 mat=None, c=None, refresh: bool = False, **_):
        if geom == "ParticleSetTextured" and geom_attr == "DisplacedSurface":
            with self._padlock:
                self._moon_name = name
                if pos is not None:
                    self._moon["pos"] = tuple(np.asarray(pos, dtype=np.float64).reshape(-1)[:3])
                if u is not None:
                    self._moon["u"] = tuple(np.asarray(u, dtype=np.float64).reshape(-1)[:3])
                if v is not None:
                    self._moon["v"] = tuple(np.asarray(v, dtype=np.float64).reshape(-1)[:3])
                if r is not None:
                    self._moon["r"] = float(np.asarray(r, dtype=np.float64).reshape(-1)[0])
                self._push_frame()
        else:
            # a single flat-shaded particle is the visible Sun disk (moon_renderer.py:643-650): rendered on rays that
            # miss the Moon; other overlay geometry is recorded only
            self._ignored_geometry[name] = {"geom": geom, "pos": pos, "r": r, "c": c, "mat": mat}
            if geom == "ParticleSet" and mat == "flat" and pos is not None and np.asarray(pos).size == 3:
                with self._padlock:
                    self._sun_disk_name = name
                    self._push_sun_disk()
        if refresh:
            self.refresh_scene()

    def set_graph(self, name: str, pos=None, edges=None, r=None, c=None, mat=None, refresh: bool = False, **_):
        """Grid / label / pin tube geometry (renderer_labels.py:263-305, renderer_pins.py:18-55): a graph of thin tubes,
        vertices `pos` (n, 3) in scene space joined by `edges` (m, 2), radius `r` (one number or one per vertex; 0 hides),
        colour `c`.  Rendered flat-shaded in front of the surface, casting no shadow."""
        with self._padlock:
            self._ignored_geometry[name] = {"geom": "Graph", "pos": pos, "edges": edges, "r": r, "c": c, "mat": mat}
            self._push_tubes()
        if refresh:
            self.refresh_scene()

    def update_graph(self, name: str, pos=None, edges=None, r=None, c=None, mat=None, refresh: bool = False, **_):
        if name not in self._ignored_geometry:
            raise ValueError(f"no geometry named {name}")
        with self._padlock:
            self._ignored_geometry[name].update({k: w for k, w in (("pos", pos), ("edges", edges), ("r", r), ("c", c), ("mat", mat))
                                                 if w is not None})
            if self._ignored_geometry[name].get("geom") == "Graph":
                self._push_tubes()
        if refresh:
            self.refresh_scene()

    def _tube_segments(self) -> np.ndarray:
        return tube_segments(self._ignored_geometry)

    def _push_tubes(self):
        seg = np.ascontiguousarray(self._tube_segments())
        _lib.check(self._lib.mrtx_set_tubes(self._ctx, seg.ctypes.data if len(seg) else None, len(seg)))

    def delete_geometry(self, name: str):
        if name == self._moon_name:
            raise ValueError("the displaced surface cannot be deleted")
        g = self._ignored_geometry.pop(name, None)
        if g is not None and g.get("geom") == "Graph":
            with self._padlock:
                self._push_tubes()
        if name == getattr(self, "_sun_disk_name", None):
            with self._padlock:
                self._sun_disk_name = None
                self._push_sun_disk()

    def get_geometry_names(self):
        return [self._moon_name] + list(k for k in self._ignored_geometry if not k.startswith("material:"))

    def update_data(self, name: str, pos=None, r=None, u=None, v=None, c=None, refresh: bool = False, **_):
        if name == self._moon_name:
            with self._padlock:
                if pos is not None:
                    self._moon["pos"] = tuple(np.asarray(pos, dtype=np.float64).reshape(-1)[:3])
                if u is not None:
                    self._moon["u"] = tuple(np.asarray(u, dtype=np.float64).reshape(-1)[:3])
                if v is not None:
                    self._moon["v"] = tuple(np.asarray(v, dtype=np.float64).reshape(-1)[:3])
                if r is not None:
                    self._moon["r"] = float(np.asarray(r, dtype=np.float64).reshape(-1)[0])
                self._push_frame()
        elif name in self._ignored_geometry:
            self._ignored_geometry[name].update({k: w for k, w in (("pos", pos), ("r", r), ("c", c)) if w is not None})
            if name == getattr(self, "_sun_disk_name", None):
                with self._padlock:
                    self._push_sun_disk()
        else:
            raise ValueError(f"no geometry named {name}")
        if refresh:
            self.refresh_scene()

    def set_displacement(self, name: str, data, refresh: bool = False, **_):
        if name != self._moon_name:
            raise ValueError(f"{name} is not a displaced surface")
        a = np.ascontiguousarray(data)
        if a.ndim != 2 or a.dtype != np.float32:
            raise ValueError("displacement map must be a 2-D float32 array")
        with self._padlock:
            _lib.check(self._lib.mrtx_set_displacement_f32(self._ctx, a.ctypes.data, a.shape[1], a.shape[0]))
        if refresh:
            self.refresh_scene()

    def set_displacement_i16(self, name: str, counts, radius_scale: float,
                             scale: float = 0.5 / 1_737_400.0, refresh: bool = False):
        """
        B200 extension: the same surface straight from the int16 LDEM counts (what
        load_elevation_data(ds=1) would produce as float32, data_loader.py:215-242) at half
        the memory.  `counts` is a host int16 array or a DeviceBuffer (zero-copy).
        """
        if name != self._moon_name:
            raise ValueError(f"{name} is not a displaced surface")
        from .device import DeviceBuffer
        with self._padlock:
            if isinstance(counts, tuple) and isinstance(counts[0], DeviceBuffer):
                buf, W, H = counts
                self._displacement_keepalive = buf
                _lib.check(self._lib.mrtx_set_displacement_i16_dev(self._ctx, buf.ptr, W, H, float(np.float32(scale)),
                                                                   float(radius_scale), 0))
            else:
                a = np.ascontiguousarray(counts)
                if a.ndim != 2 or a.dtype != np.int16:
                    raise ValueError("counts must be a 2-D int16 array")
                _lib.check(self._lib.mrtx_set_displacement_i16(self._ctx, a.ctypes.data, a.shape[1], a.shape[0],
                                                               float(np.float32(scale)), float(radius_scale)))
        if refresh:
            self.refresh_scene()

    # ---- camera (moon_renderer.py:627-635, 565-568; renderer_navigation.py) -------------------
    def setup_camera(self, name: str, eye=None, target=None, up=None, cam_type: str = "Pinhole", fov: float = -1,
                     aperture_radius=None, aperture_fract=None, focal_scale=None, make_current: bool = True, **_):
        if cam_type != "Pinhole":
            raise ValueError("B200OptiX implements the Pinhole camera only (shared_types.py:56-69)")
        self._cam_name = name
        self.update_camera(name, eye=eye, target=target, up=up, fov=fov if fov and fov > 0 else None)

    def update_camera(self, name: Optional[str] = None, eye=None, target=None, up=None, fov=None, **_):
        with self._padlock:
            if eye is not None:
                self._cam["eye"] = tuple(float(x) for x in eye)
            if target is not None:
                self._cam["target"] = tuple(float(x) for x in target)
            if up is not None:
                self._cam["up"] = tuple(float(x) for x in up)
            if fov is not None:
                self._cam["fov"] = float(fov)
            self._push_camera()
        if self._is_started:
            self.refresh_scene()

    def get_camera(self, name: Optional[str] = None) -> dict:
        c = self._cam
        return {"Eye": list(c["eye"]), "Target": list(c["target"]), "Up": list(c["up"]), "FoV": c["fov"],
                "Type": "Pinhole"}

    def get_camera_name_handle(self, name=None):
        return self._cam_name, 0

    # ---- light (moon_renderer.py:640, 347, 860) ------------------------------------------------
    def setup_light(self, name: str, pos=None, color=None, radius: float = -1, in_geometry: bool = True, **_):
        self._light_name = name
        self.update_light(name, pos=pos, color=color, radius=radius if radius is not None and radius >= 0 else None)

    def update_light(self, name: str, pos=None, color=None, radius=None, **_):
        with self._padlock:
            if pos is not None:
                self._light["pos"] = tuple(float(x) for x in pos)
            if color is not None:
                col = np.asarray(color, dtype=np.float64).reshape(-1)
                self._light["color"] = float(col[0])       # MoonRTX passes a scalar radiance
            if radius is not None:
                self._light["radius"] = float(radius)
            self._push_light()

    # ---- run ----------------------------------------------------------------------------------
    def refresh_scene(self):
        self._dirty.set()

    def set_accum_done_cb(self, cb: Optional[Callable]):
        self._accum_done_cb = cb

    def set_launch_finished_cb(self, cb: Optional[Callable]):
        self._on_launch_finished = cb

    # ---- multi-GPU sharding of ONE frame; time-lapse needs none of this ------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128 opaque bytes rank 0 creates and hands to every rank (any side channel)."""
        buf = (C.c_uint8 * 128)()
        _lib.check(_lib.load().mrtx_comm_unique_id(_lib.nccl_library_path().encode(), buf))
        return bytes(buf)

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        with self._padlock:
            _lib.check(self._lib.mrtx_comm_init(self._ctx, _lib.nccl_library_path().encode(), int(world), int(rank), buf))
        self._rank, self._world = int(rank), int(world)

    def p2p_open(self, rank: int, world: int) -> bytes:
        """Peer-memory frame delivery (mrtx_p2p_open): allocate this rank's mailbox for frames of the current size and
        return its CUDA IPC handle (64 bytes) for the host to exchange between the ranks."""
        buf = (C.c_uint8 * 64)()
        with self._padlock:
            _lib.check(self._lib.mrtx_p2p_open(self._ctx, int(world), int(rank), self._width * self._height * 4, buf))
        self._rank, self._world = int(rank), int(world)
        return bytes(buf)

    def p2p_connect(self, handles: Sequence[bytes]) -> None:
        """Map every rank's mailbox (handles in rank order, as returned by p2p_open on each): from now on
        submit_frame(dst=...) / recv_frame move frames with the copy engines over NVLink, no kernel on either side."""
        blob = b"".join(handles)
        if len(blob) != 64 * len(handles):
            raise ValueError("every handle is 64 bytes")
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        with self._padlock:
            _lib.check(self._lib.mrtx_p2p_connect(self._ctx, buf))

    def p2p_close(self) -> None:
        """Give the mailboxes up (all ranks, no frame in flight): delivery goes back to ncclSend / ncclRecv."""
        with self._padlock:
            _lib.check(self._lib.mrtx_p2p_close(self._ctx))

    def render_cycle(self, read_back: bool = True, shard: Optional[str] = None,
                     tile_rows: int = 64, tile: int = 64) -> Optional[np.ndarray]:
        """
        One accumulation cycle, synchronously, on the calling thread (padlock held).

        shard=None      this GPU renders the whole frame;
        shard="samples" the cycle's samples are split across the communicator's ranks and the float4
                        accumulators are summed with one ncclAllReduce before the resolve;
        shard="tiles"   interleaved square tiles of side `tile` (a power of two) are split across the ranks in ONE
                        launch, tone-mapped straight into the send buffer and exchanged with one ncclAllGather;
        shard="rows"    interleaved bands of `tile_rows` rows are split across the ranks and the
                        resolved RGBA8 bands are exchanged with one ncclAllGather.
        """
        with self._padlock:
            n = max(1, int(self._params["max_accumulation_frames"]))
            jitter = (n > 1) if self.deterministic is None else (not self.deterministic)
            _lib.check(self._lib.mrtx_set_uint(self._ctx, b"jitter", 1 if jitter else 0, 0))
            W, H = self._width, self._height
            if shard is None:
                # PlotOptiX launches `min_accumulation_step` samples at a time so that a first noisy image is on screen
                # early (the reference asks for 1 of 64, moon_renderer.py:578).  Nothing observes the intermediate
                # states of this synchronous cycle - the callbacks fire once, after it - and the tracer is 2x more
                # efficient when the lanes of a warp trace samples of the SAME pixel (4K: 2.1 Grays/s at 1 sample per
                # launch, 4.5 at 16+), so the cycle goes down in one call; the library cuts it into launches of 32
                # samples.  Sample s of pixel p is the same ray either way (RNG keyed on (p, s)).
                _lib.check(self._lib.mrtx_render(self._ctx, 0, 0, W, H, 0, n, 1))
            elif shard == "samples":
                lo = (n * self._rank) // self._world
                hi = (n * (self._rank + 1)) // self._world
                _lib.check(self._lib.mrtx_render(self._ctx, 0, 0, W, H, lo, hi - lo, 1))
                _lib.check(self._lib.mrtx_allreduce_accum(self._ctx))
            elif shard == "tiles":
                _lib.check(self._lib.mrtx_render_tiles(self._ctx, int(tile), 0, n, 1))
            elif shard == "rows":
                first = True
                for t in range(self._rank, (H + tile_rows - 1) // tile_rows, self._world):
                    y0, y1 = t * tile_rows, min(H, (t + 1) * tile_rows)
                    _lib.check(self._lib.mrtx_render(self._ctx, 0, y0, W, y1, 0, n, 1 if first else 0))
                    first = False
                if first:                               # more ranks than bands: still clear the accumulators
                    _lib.check(self._lib.mrtx_render(self._ctx, 0, 0, W, 0, 0, 0, 1))
            else:
                raise ValueError(f"unknown shard mode {shard}")
            if shard == "tiles":
                _lib.check(self._lib.mrtx_allgather_tiles(self._ctx, int(tile)))      # resolves the owned tiles itself
            else:
                _lib.check(self._lib.mrtx_resolve(self._ctx))
            if shard == "rows":
                _lib.check(self._lib.mrtx_allgather_rows(self._ctx, int(tile_rows)))
            self._frames_rendered += 1
            if read_back:
                _lib.check(self._lib.mrtx_read_rgba8(self._ctx, self._img_rgba.ctypes.data))
                if self._encoder is not None:
                    self._encoder.grab(self._img_rgba)
            else:
                _lib.check(self._lib.mrtx_synchronize(self._ctx))
            if self._on_launch_finished is not None:
                self._on_launch_finished(self)
            if self._accum_done_cb is not None:
                self._accum_done_cb(self)
            return self._img_rgba if read_back else None

    # ---- pipelined frames (time-lapse export) --------------------------------------------------------
    def kernel_times(self, reset: bool = False) -> dict:
        """Per-kernel CUDA-event times of the mrtx_render calls since the last reset (engine switch "profile")."""
        out = (C.c_double * 8)()
        with self._padlock:
            _lib.check(self._lib.mrtx_kernel_times(self._ctx, out, 1 if reset else 0))
        # ("trace_kernel_fast" is the primary-ray kernel of the launch: trace_kernel_pool unless shadow_queue <= 2)
        names = ("cull_kernel", "beam_kernel", "trace_kernel_fast", "shadow_kernel", "trace_kernel_referee", "fold_kernel")
        d = {k: float(out[i]) for i, k in enumerate(names)}
        d["shade_kernel"] = float(out[7])
        d["launches"] = int(out[6])
        return d

    def recv_frame(self, src: int) -> int:
        """Consumer side of a frame-parallel time-lapse: post the receive of the next frame that rank `src` sends with
        submit_frame(dst=this rank).  Returns a ticket for wait_recv(); at most two receives are pending."""
        with self._padlock:
            if not hasattr(self, "_recv_out"):
                self._recv_out = [self.pinned_empty((self._height, self._width, 4), np.uint8) for _ in range(2)]
            ticket = C.c_int()
            # (the library hands out its two staging slots in turn: the pinned buffer of the same index goes with it)
            k = getattr(self, "_recv_next", 0)
            _lib.check(self._lib.mrtx_frame_recv(self._ctx, int(src), self._recv_out[k].ctypes.data, C.byref(ticket)))
            assert ticket.value == k
            self._recv_next = k ^ 1
            return k

    def wait_recv(self, ticket: int) -> np.ndarray:
        """Block until the received frame is in host memory (pinned [H, W, 4] uint8, reused by the receive after next)."""
        _lib.check(self._lib.mrtx_frame_recv_wait(self._ctx, int(ticket)))
        img = self._recv_out[int(ticket)]
        with self._padlock:
            if self._encoder is not None:
                self._encoder.grab(img)
        return img

    def submit_frame(self, overlay: Optional[np.ndarray] = None, dst: Optional[int] = None) -> int:
        """
        Queue one whole accumulation cycle of the scene as it is now and return at once: overlay upload (an RGBA8
        [H, W, 4] array, copied to pinned memory here; None = no overlay), tracing, resolve and read-back run on the
        GPU while the caller prepares the next frame.  At most two frames are in flight.  Returns a ticket for
        wait_frame().  The F11 loop of renderer_video.py (overlay, update_view, accumulate, grab) maps to
        submit_frame(i + 1) before wait_frame(i).
        """
        with self._padlock:
            if not hasattr(self, "_pipe_out"):
                shape = (self._height, self._width, 4)
                self._pipe_out = [self.pinned_empty(shape, np.uint8) for _ in range(2)]
                self._pipe_ovl = [self.pinned_empty(shape, np.uint8) for _ in range(2)]
                self._pipe_next = 0
                self._pipe_sent = [False, False]
            k = self._pipe_next
            n = max(1, int(self._params["max_accumulation_frames"]))
            jitter = (n > 1) if self.deterministic is None else (not self.deterministic)
            _lib.check(self._lib.mrtx_set_uint(self._ctx, b"jitter", 1 if jitter else 0, 0))
            ov_ptr = None
            if overlay is not None:
                if overlay.shape != self._pipe_ovl[k].shape or overlay.dtype != np.uint8:
                    raise ValueError(f"overlay must be uint8 {self._pipe_ovl[k].shape}")
                # (the library's slot k was last read two submits ago: wait_frame(k) has been called since, or is now)
                if self._frames_submitted >= 2:
                    _lib.check(self._lib.mrtx_frame_wait(self._ctx, k))      # (no-op if wait_frame(k) has been called)
                np.copyto(self._pipe_ovl[k], overlay)
                ov_ptr = self._pipe_ovl[k].ctypes.data
            ticket = C.c_int()
            if dst is None or int(dst) == getattr(self, "_rank", 0):
                _lib.check(self._lib.mrtx_frame_submit(self._ctx, ov_ptr, n, self._pipe_out[k].ctypes.data, C.byref(ticket)))
                self._pipe_sent[k] = False
            else:
                # the frame goes to rank dst over NVLink (ncclSend) instead of this rank's host memory
                _lib.check(self._lib.mrtx_frame_submit_to(self._ctx, ov_ptr, n, int(dst), C.byref(ticket)))
                self._pipe_sent[k] = True
            assert ticket.value == k
            self._pipe_next ^= 1
            self._frames_submitted += 1
            return k

    def wait_frame(self, ticket: int) -> np.ndarray:
        """Block until the frame of submit_frame() is in host memory; the returned array (pinned, [H, W, 4] uint8) is
        reused by the submit after next.  Fires the launch-finished / accumulation-done callbacks like render_cycle."""
        _lib.check(self._lib.mrtx_frame_wait(self._ctx, int(ticket)))
        img = self._pipe_out[int(ticket)]
        if self._pipe_sent[int(ticket)]:
            with self._padlock:
                self._frames_rendered += 1
            return None                                  # the pixels are on their way to the consumer rank
        with self._padlock:
            self._frames_rendered += 1
            if self._encoder is not None:
                self._encoder.grab(img)
            if self._on_launch_finished is not None:
                self._on_launch_finished(self)
            if self._accum_done_cb is not None:
                self._accum_done_cb(self)
        return img

    def _run(self):
        while not self._stop.is_set():
            if not self._dirty.wait(timeout=0.05):
                continue
            self._dirty.clear()
            try:
                self.render_cycle()
            except _lib.MoonB200Error as e:
                # a CUDA error is sticky: every later cycle would fail the same way, 20 times a second
                print(f"B200OptiX render thread stopped: {e}")
                self._render_error = e
                self._is_started = False
                return
            except Exception as e:                       # PlotOptiX logs and continues
                print(f"B200OptiX render thread: {e}")

    def start(self):
        if self._is_started:
            return
        self._is_started = True
        self._stop.clear()
        self._dirty.set()
        self._thread = threading.Thread(target=self._run, name="B200OptiX-render", daemon=True)
        self._thread.start()

    def close(self):
        self._stop.set()
        if self._thread is not None and self._thread is not threading.current_thread():
            self._thread.join(timeout=10)
        self._thread = None
        self._is_started = False
        self._is_closed = True
        if self._encoder is not None:
            self._encoder.stop()
            self._encoder = None
        if self._dev is not None:
            self._img_rgba = np.array(self._img_rgba)      # detach from the pinned allocation
            for p in self._pinned:
                self._lib.mrtx_host_free(p)
            self._pinned = []
        if self._own_device and self._dev is not None:
            self._dev.close()
            self._dev = None

    # ---- read-back (moon_renderer.py:1137-1142, renderer_dialogs.py:1222-1224) -----------------
    def get_rt_output(self) -> np.ndarray:
        return self._img_rgba.copy()

    def _get_image_xy(self, x, y):
        return int(x), int(y)

    def _get_hit_at(self, x: int, y: int):
        if not (0 <= x < self._width and 0 <= y < self._height):
            return 0.0, 0.0, 0.0, 0.0
        out = (C.c_float * 4)()
        with self._padlock:
            _lib.check(self._lib.mrtx_hit_at(self._ctx, int(x), int(y), out))
        return float(out[0]), float(out[1]), float(out[2]), float(out[3])

    def get_hit_buffer(self) -> np.ndarray:
        out = np.empty((self._height, self._width, 4), dtype=np.float32)
        with self._padlock:
            _lib.check(self._lib.mrtx_read_hit_f32(self._ctx, out.ctypes.data))
        return out

    def get_accum_buffer(self) -> np.ndarray:
        out = np.empty((self._height, self._width, 4), dtype=np.float32)
        with self._padlock:
            _lib.check(self._lib.mrtx_read_accum_f32(self._ctx, out.ctypes.data))
        return out

    def get_hit_records_f64(self) -> np.ndarray:
        """(s_hit, radius, lon, lat) per pixel of the last 1-spp launch (needs set_uint('debug_hits', 1))."""
        out = np.empty((self._height, self._width, 4), dtype=np.float64)
        with self._padlock:
            _lib.check(self._lib.mrtx_read_hit_f64(self._ctx, out.ctypes.data))
        return out

    def counters(self, reset: bool = False) -> dict:
        out = (C.c_uint64 * 16)()
        _lib.check(self._lib.mrtx_counters(self._ctx, out, 1 if reset else 0))
        names = ("primary_rays", "primary_in_sphere", "primary_hits", "shadow_rays", "shadow_occluded",
                 "node_visits", "patch_tests", "overflow", "test_phases", "test_phase_lanes", "trav_steps",
                 "trav_step_lanes", "start_phases", "start_phase_lanes", "refills", "pixels_culled")
        return {k: int(out[i]) for i, k in enumerate(names)}

    def defer_stats(self, reset: bool = False) -> dict:
        """Samples the filtered kernel handed to the exact kernel, and why (include/moonb200.h)."""
        out = (C.c_uint64 * 32)()
        _lib.check(self._lib.mrtx_defer_stats(self._ctx, out, 1 if reset else 0))
        return {"deferred_samples": int(out[0]), "primary_reasons": {r: int(out[r]) for r in range(1, 16) if out[r]},
                "shadow_reasons": {r: int(out[16 + r]) for r in range(1, 16) if out[16 + r]}}

    def save_image(self, path: str, bps: str = "Bps8"):
        import cv2
        img = self._img_rgba
        if bps == "Bps16":
            cv2.imwrite(path, (img[..., [2, 1, 0, 3]].astype(np.uint16) * 257))
        else:
            cv2.imwrite(path, img[..., [2, 1, 0, 3]])

    # ---- encoder (renderer_video.py:222-256, 290, 339-340) ----------------------------------------
    def encoder_create(self, fps: int = 25, bitrate: float = 16, idrrate=None, profile=None, preset=None):
        from .video import FrameSink
        self._encoder_cfg = (int(fps), float(bitrate))
        FrameSink.check_available()

    def encoder_start(self, out_name: str, n_frames: int = 0):
        from .video import FrameSink
        fps, _ = getattr(self, "_encoder_cfg", (25, 16.0))
        self._encoder = FrameSink(out_name, self._width, self._height, fps, n_frames)

    def encoder_is_open(self) -> bool:
        return self._encoder is not None and self._encoder.is_open()

    def encoder_stop(self):
        if self._encoder is not None:
            self._encoder.stop()
