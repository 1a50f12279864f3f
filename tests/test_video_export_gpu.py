"""
The F11 export end to end on the drop-in: the accumulation-done callback loop of
renderer_video.py:276-320 - count the frame, draw the NEXT frame's overlay, update_view to the next time step, which
refreshes the scene and starts the next cycle - driven by B200OptiX's own render thread, with the encoder attached as
the reference attaches it (encoder_create / encoder_start / encoder_is_open, :222-252).  The reference's own acceptance
check for this loop is repeated: a marker sequence 30/70/110/150/190/230 burnt in through the overlay comes back from
the encoded frames in order with no off-by-one, and blending is exact alpha compositing (renderer_video.py:21-25).
(tests/test_dropin_surface.py runs the UNMODIFIED reference callback on the CPU and pins the call order restated here.)
"""
import threading

import numpy as np
import pytest

from helpers import make_gpu, sun_at_phase

pytestmark = pytest.mark.gpu

MARKERS = [30, 70, 110, 150, 190, 230, 90, 210]
W, H = 256, 160


def overlay_with_marker(v):
    buf = np.zeros((H, W, 4), np.uint8)
    buf[8:40, 8:72] = (v, v, v, 255)                # opaque patch: renders exactly v
    buf[H - 40:H - 8, 8:72] = (0, 0, 0, 128)        # 50 % black over whatever is underneath
    return buf


class CapturingSink:
    """stands where FrameSink stands (B200OptiX._encoder): keeps the frames the encoder is handed"""
    def __init__(self, n):
        self.frames, self.n = [], n
    def is_open(self):
        return len(self.frames) < self.n
    def grab(self, rgba):
        if self.is_open():
            self.frames.append(rgba.copy())
    def stop(self):
        self.n = len(self.frames)


def scene_states(n):
    from moonrtx_b200 import scene
    from moonrtx_b200.synth import synth_ephemeris
    return [scene.frame_state(synth_ephemeris(600.0 * i)) for i in range(n)]


def elevation():
    from moonrtx_b200.synth import synth_ldem
    from moonrtx_b200.data_loader import downscale_elevation
    return downscale_elevation(synth_ldem(720, 360, seed=9, craters=40), 1)[0]


def test_callback_driven_export_burns_each_frames_own_overlay(tmp_path):
    from moonrtx_b200.video import apply_frame_state
    n = len(MARKERS)
    states = scene_states(n)
    rt = make_gpu(elevation(), W, H, debug_hits=False, light_pos=sun_at_phase(90.0))
    rt.set_param(min_accumulation_step=1, max_accumulation_frames=4)
    rt.add_postproc("Overlay")
    done = threading.Event()
    st = {"frame": 0, "n": n, "error": None}

    def accum_done(r):                                   # renderer_video.py:276-320, on the render thread, padlock held
        st["frame"] += 1
        if st["frame"] < st["n"]:
            if not r.encoder_is_open():
                st["error"] = "encoder closed early"
                done.set()
                return
            with r._padlock:
                r.set_texture_2d("frame_overlay", overlay_with_marker(MARKERS[st["frame"]]), filter_mode="Nearest", refresh=False)
                apply_frame_state(r, states[st["frame"]])          # update_view's rt.* calls (moon_renderer.py:852-860) ...
                r.refresh_scene()                                   # ... and its refresh (:871)
        else:
            r.set_accum_done_cb(None)
            done.set()

    # the video file the reference asks for (mp4v sink), and beside it the exact frames the encoder was handed
    rt.encoder_create(fps=25, bitrate=16)
    path = str(tmp_path / "lapse.mp4")
    rt.encoder_start(path, n)
    assert rt.encoder_is_open()
    file_sink = rt._encoder
    cap = CapturingSink(n)

    class Tee:
        def is_open(self): return cap.is_open()
        def grab(self, img): file_sink.grab(img); cap.grab(img)
        def stop(self): file_sink.stop(); cap.stop()
    rt._encoder = Tee()
    rt.set_accum_done_cb(accum_done)
    with rt._padlock:
        rt.set_texture_2d("frame_overlay", overlay_with_marker(MARKERS[0]), filter_mode="Nearest", refresh=False)
        apply_frame_state(rt, states[0])
    rt.start()                                           # first cycle = first frame (renderer_video.py:262-268)
    assert done.wait(timeout=120), "export did not finish"
    assert st["error"] is None
    rt._stop.set()
    assert len(cap.frames) == n
    got = [int(f[20, 30, 0]) for f in cap.frames]
    assert got == MARKERS, got                           # in order, no off-by-one
    for f in cap.frames:
        assert np.all(f[8:40, 8:72, :3] == f[20, 30, 0])                      # opaque patch exact
    # 50 % black patch = exact alpha compositing of the tone-mapped pixel underneath (46 -> 23 in the reference's check)
    plain = make_gpu(elevation(), W, H, debug_hits=False, light_pos=sun_at_phase(90.0))
    plain.set_param(min_accumulation_step=1, max_accumulation_frames=4)
    apply_frame_state(plain, states[2])
    base = plain.render_cycle().copy()
    plain.close()
    under = base[H - 40:H - 8, 8:72, :3].astype(np.int32)
    want = (under * 127 + 127) // 255
    assert np.array_equal(cap.frames[2][H - 40:H - 8, 8:72, :3].astype(np.int32), want)
    # the file on disk has the frames too (lossy codec: the markers within a few levels, in order)
    import cv2
    vc = cv2.VideoCapture(path)
    levels = []
    while True:
        ok, fr = vc.read()
        if not ok:
            break
        levels.append(float(fr[12:36, 12:68].mean()))
    vc.release()
    assert len(levels) == n and all(abs(a - b) <= 8 for a, b in zip(levels, MARKERS)), levels
    rt.close()


def test_pipelined_export_delivers_the_same_frames():
    """the same export through submit_frame / wait_frame (video.render_timelapse(pipelined=True)): identical frames"""
    from moonrtx_b200.video import render_timelapse
    n = 6
    states = scene_states(n)
    outs = []
    for pipelined in (False, True):
        rt = make_gpu(elevation(), W, H, debug_hits=False, light_pos=sun_at_phase(90.0))
        rt.set_param(min_accumulation_step=4, max_accumulation_frames=4)
        rt.add_postproc("Overlay")
        cap = CapturingSink(n)
        rt._encoder = cap
        render_timelapse(rt, states, overlay_for=lambda i: overlay_with_marker(MARKERS[i]), keep=False, pipelined=pipelined)
        outs.append(cap.frames)
        rt.close()
    assert len(outs[0]) == n and len(outs[1]) == n
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert [int(f[20, 30, 0]) for f in outs[1]] == MARKERS[:n]
