"""
Drop-in surface against the reference's own code (no GPU): the UNMODIFIED `MoonRenderer.init_renderer`,
`update_view` and the overlay part of the F11 export are run with `TkOptiX` replaced by a proxy that accepts a
call only if `moonrtx_b200.optix.B200OptiX` defines that method and the arguments bind to its signature.  A method the
reference calls and the drop-in lacks, or a keyword it does not know, fails here - before anybody needs a B200.

Skipped where /root/reference does not exist (the GPU box).
"""

import inspect
import threading

import numpy as np
import pytest

from oracle import ref_stub

pytestmark = pytest.mark.skipif(not ref_stub.reference_available(), reason="reference tree not on this box")


def make_proxy_class(calls):
    from moonrtx_b200.optix import B200OptiX

    class SurfaceProxy:
        """Records calls; every one must bind to the B200OptiX method of the same name."""

        def __init__(self, *a, **k):
            inspect.signature(B200OptiX.__init__).bind(self, *a, **k)
            calls.append(("__init__", a, k))
            self._padlock = threading.RLock()
            self._width, self._height = k.get("width", 16), k.get("height", 16)
            self._is_started = True
            self._cams = {}
            self._optix = self

        # what the reference reads back
        def get_camera(self, name):
            inspect.signature(B200OptiX.get_camera).bind(self, name)
            return dict(self._cams[name])

        def get_camera_fov(self, handle=0):
            return 4.0

        def set_camera_fov(self, fov):
            calls.append(("_optix.set_camera_fov", (fov,), {}))

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            target = getattr(B200OptiX, name, None)
            if target is None or not callable(target):
                raise AttributeError(f"the reference calls rt.{name}, which B200OptiX does not define")
            sig = inspect.signature(target)

            def call(*a, **k):
                sig.bind(self, *a, **k)                      # TypeError: unknown keyword / wrong arity
                calls.append((name, a, k))
                if name in ("setup_camera", "update_camera"):
                    cam = self._cams.setdefault(a[0] if a else k.get("name"), {"Eye": [0, -300, 0], "Target": [0, 0, 0], "Up": [0, 0, 1]})
                    for key, field in (("eye", "Eye"), ("target", "Target"), ("up", "Up")):
                        if k.get(key) is not None:
                            cam[field] = list(k[key])
                return None
            return call

    return SurfaceProxy


def bare_renderer(mr, tmp_path):
    """A MoonRenderer with the state __init__ would have set, without Tk, Skyfield or the 9 GB of data."""
    from moonrtx_b200.synth import synth_ephemeris
    import cv2
    r = mr.MoonRenderer.__new__(mr.MoonRenderer)
    r.width, r.height = 320, 200
    r.gamma, r.brightness = 2.2, 80
    r.downscale, r.color_downscale = 4, 1
    g = np.random.default_rng(5)
    r.elevation = (1.0 - 0.01 * g.random((90, 180))).astype(np.float32)
    r.elevation_radius_scale = 1.0062
    color = tmp_path / "color.tif"
    cv2.imwrite(str(color), g.integers(0, 256, (64, 128, 3), dtype=np.uint8))
    r.color_file, r.starmap_file = str(color), None
    r.moon_ephem = synth_ephemeris(0.0)
    r.moon_rotation = np.eye(3)
    r.moon_rotation_inv = np.eye(3)
    r.initial_camera = r.default_camera            # what init_astro resolves from the ephemeris (moon_renderer.py:507-520)
    r._apparent_radius = None
    r.rt = None
    return r


def test_reference_init_renderer_and_update_view_bind_to_the_dropin(tmp_path):
    mr = ref_stub.import_reference("moon_renderer")
    calls = []
    mr.TkOptiX = make_proxy_class(calls)
    r = bare_renderer(mr, tmp_path)
    r.init_renderer()
    names = [c[0] for c in calls]
    # the calls SURVEY.md 8b lists for init_renderer (moon_renderer.py:570-650) all arrived and bound
    for must in ("__init__", "set_param", "set_uint", "set_float", "set_ambient", "add_postproc", "set_background",
                 "set_texture_2d", "update_material", "set_data", "set_displacement", "setup_camera", "setup_light"):
        assert must in names, f"init_renderer never reached rt.{must}: {names}"
    disp = next(c for c in calls if c[0] == "set_displacement")
    assert disp[1][0] == "moon" and disp[1][1].dtype == np.float32 and disp[2].get("refresh") is False
    data = next(c for c in calls if c[0] == "set_data" and c[1][0] == "moon")
    assert data[2]["geom"] == "ParticleSetTextured" and data[2]["geom_attr"] == "DisplacedSurface" and data[2]["r"] == 10.0
    for must in ("setup_material",):
        assert must in names

    # update_view (moon_renderer.py:824-871) with the ephemeris and the Tk parts stubbed
    from datetime import datetime
    from moonrtx_b200.synth import synth_ephemeris
    mr.astro.calculate_moon_ephemeris = lambda dt, parallactic: synth_ephemeris(600.0)
    r.parallactic_mode = False
    r.dt_local = datetime(2026, 1, 1)
    r.in_observer_clock = lambda d: d
    r.update_overlays = lambda: None
    r.sync_datetime_dialog = lambda: None
    n0 = len(calls)
    r.update_view()
    later = [c[0] for c in calls[n0:]]
    assert later.count("update_data") == 2 and "update_light" in later and "refresh_scene" in later, later
    light = next(c for c in calls[n0:] if c[0] == "update_light")
    assert set(light[2]) == {"pos", "radius"}

    # navigation (renderer_navigation.py:27-73, 226-297, 299-354, 356-450, 494-523): camera read-modify-write
    import types
    n1 = len(calls)
    r.view_orientation = r.initial_view_orientation = mr.VIEW_ORIENTATION_NSWE if hasattr(mr, "VIEW_ORIENTATION_NSWE") else "NSWE"
    r.center_on_lat_lon(10.0, 20.0)
    r.navigate_view("left")
    r.pan_tilt_view(12.0, -7.0)
    r.rotate_around_moon_axis("left")
    r.rotate_around_view_direction("right")
    r.zoom_with_wheel(types.SimpleNamespace(delta=120, num=0))
    nav = [c[0] for c in calls[n1:]]
    assert nav.count("update_camera") >= 3 and "_optix.set_camera_fov" in nav, nav


def test_reference_video_overlay_calls_bind_to_the_dropin():
    """renderer_video.py:123-144, 222-260: overlay texture, Overlay post-process, encoder and accumulation callback."""
    from moonrtx_b200.optix import B200OptiX
    calls = []
    rt = make_proxy_class(calls)(width=64, height=48)
    rt.set_texture_2d("frame_overlay", np.zeros((48, 64, 4), np.uint8), filter_mode="Nearest", refresh=False)
    rt.add_postproc("Overlay")
    rt.encoder_create(fps=25, bitrate=16)
    rt.encoder_start("out.mp4", 120)
    rt.encoder_is_open()
    rt.set_accum_done_cb(lambda rt_: None)
    rt.set_accum_done_cb(None)
    rt.encoder_stop()
    rt.refresh_scene()
    rt._get_hit_at(3, 4)
    rt._get_image_xy(3, 4)
    rt.save_image("x.png", bps="Bps16")
    for name in ("_padlock", "_width", "_height", "_is_started", "_optix"):
        assert name in B200OptiX.__init__.__code__.co_names or hasattr(B200OptiX, name) or name in inspect.getsource(B200OptiX), name
    assert len(calls) == 13


def test_unmodified_reference_export_callback_call_order(tmp_path):
    """renderer_video.py:148-320 run UNMODIFIED against the recording proxy: per exported frame the reference draws the
    next frame's overlay, then update_view moves the scene and refreshes - the order tests/test_video_export_gpu.py
    restates on the GPU box (where the reference tree does not exist).  Also drives update_overlays with the grid and a pin
    enabled (renderer_labels.py, renderer_pins.py): set_graph / update_graph / delete_geometry must exist on the drop-in."""
    from datetime import datetime, timedelta
    mr = ref_stub.import_reference("moon_renderer")
    calls = []
    mr.TkOptiX = make_proxy_class(calls)
    r = bare_renderer(mr, tmp_path)
    r._init_video_export()
    r.init_renderer()
    from moonrtx_b200.synth import synth_ephemeris
    mr.astro.calculate_moon_ephemeris = lambda dt, parallactic: synth_ephemeris(600.0)
    r.parallactic_mode = False
    r.dt_local = datetime(2026, 1, 1)
    r.in_observer_clock = lambda d: d
    r.shifted_time = lambda minutes: r.dt_local + timedelta(minutes=minutes)
    r.update_overlays = lambda: None
    r.sync_datetime_dialog = lambda: None
    r._auto_advance_var = None
    r._preview_restore_id = None
    r._preview_active = False
    posted = []
    r.rt._root = type("Root", (), {"after": lambda self, ms, fn, *a: posted.append((fn, a)), "after_cancel": lambda self, i: None})()
    r.rt.encoder_is_open = lambda: True                 # (a started encoder is always open, renderer_video.py:246-252)
    progress = []
    err = r.start_video_export(str(tmp_path / "x.mp4"), 3, 10, 25, 16.0, lambda f, n, dt: progress.append(f), lambda e: progress.append(("done", e)))
    assert err is None
    names = [c[0] for c in calls]
    for must in ("encoder_create", "encoder_start", "set_accum_done_cb", "set_texture_2d", "add_postproc", "refresh_scene"):
        assert must in names, must
    cb = next(c for c in calls if c[0] == "set_accum_done_cb")[1][0]
    for frame in (1, 2):
        n0 = len(calls)
        cb(r.rt)                                        # what the render thread does after each accumulation cycle
        seq = [c[0] for c in calls[n0:]]
        i_ov, i_data, i_ref = seq.index("set_texture_2d"), seq.index("update_data"), seq.index("refresh_scene")
        assert i_ov < i_data < i_ref, seq                # overlay of the NEXT frame first, then update_view, then the refresh
        ov = calls[n0 + i_ov]
        assert ov[1][0] == "frame_overlay" and ov[2] == {"filter_mode": "Nearest", "refresh": False}
    n0 = len(calls)
    cb(r.rt)                                            # third frame: the export ends
    assert ("set_accum_done_cb", (None,), {}) in calls[n0:]

    # grid + pins through the unmodified overlay code: the geometry calls exist and bind
    for name in ("set_graph", "update_graph", "delete_geometry"):
        from moonrtx_b200.optix import B200OptiX
        assert callable(getattr(B200OptiX, name, None)), name
    # the UNMODIFIED grid / pin code of the reference (renderer_labels.py:263-305, 513-527; renderer_pins.py:18-69, 149-163)
    # and the real update_overlays (moon_renderer.py:781-789) with both visible
    del r.update_overlays
    r.view_orientation = r.initial_view_orientation = getattr(mr, "VIEW_ORIENTATION_NSWE", "NSWE")
    r.moon_grid, r.moon_grid_visible = None, False
    r.standard_labels_visible = r.spot_labels_visible = False
    r.pins, r.pins_visible = {}, True
    n0 = len(calls)
    r.setup_moon_grid()
    r.create_pin(3, 10.0, 20.0)
    r.update_overlays()
    r.remove_pin(3)
    geo = [c for c in calls[n0:] if c[0] in ("set_graph", "update_graph", "delete_geometry")]
    kinds = [c[0] for c in geo]
    assert kinds.count("set_graph") == 3 and "update_graph" in kinds and kinds[-1] == "delete_geometry", kinds
    grid = next(c for c in geo if c[0] == "set_graph")
    assert set(grid[2]) >= {"pos", "edges", "r", "c", "mat"}
