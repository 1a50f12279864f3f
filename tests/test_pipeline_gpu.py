"""Pipelined frames (mrtx_frame_submit / mrtx_frame_wait, B200OptiX.submit_frame / wait_frame,
video.render_timelapse(pipelined=True)): same pixels as the synchronous sequence, frame for frame."""

import numpy as np
import pytest

from helpers import make_gpu, sun_at_phase

pytestmark = pytest.mark.gpu


def _overlay(h, w, i):
    buf = np.zeros((h, w, 4), np.uint8)
    buf[h - 20:h - 6, 6 + 3 * i:60 + 3 * i] = (255, 255, 255, 150 + 10 * i)
    buf[2:10, 2:30, :3] = 40 * i
    buf[2:10, 2:30, 3] = 255
    return buf


@pytest.mark.parametrize("spp", [1, 5])
def test_pipelined_timelapse_equals_synchronous_frames(spp):
    from moonrtx_b200 import scene
    from moonrtx_b200.synth import synth_ephemeris, synth_ldem
    from moonrtx_b200.data_loader import downscale_elevation
    from moonrtx_b200.video import render_timelapse
    elev, _ = downscale_elevation(synth_ldem(1440, 720, seed=9, craters=40), 2)
    W, H, n = 160, 96, 7
    states = [scene.frame_state(synth_ephemeris(600.0 * i)) for i in range(n)]
    results = []
    for pipelined in (False, True):
        rt = make_gpu(elev, W, H, debug_hits=False, light_pos=sun_at_phase(90.0))
        rt.set_param(min_accumulation_step=1, max_accumulation_frames=spp)
        rt.add_postproc("Overlay")
        seen = []
        frames = render_timelapse(rt, states, overlay_for=lambda i: _overlay(H, W, i), pipelined=pipelined,
                                  on_frame=lambda i, img: seen.append(i))
        assert seen == list(range(n)) and sorted(frames) == list(range(n))
        results.append(frames)
        rt.close()
    for i in range(n):
        assert np.array_equal(results[0][i], results[1][i]), f"frame {i} differs"
    # the frames are not all the same picture (the sun moves, the overlay moves)
    assert any(not np.array_equal(results[1][0], results[1][i]) for i in range(1, n))


def test_submit_without_overlay_and_callbacks():
    from moonrtx_b200.synth import synth_ldem
    from moonrtx_b200.data_loader import downscale_elevation
    elev, _ = downscale_elevation(synth_ldem(720, 360, seed=2, craters=20), 2)
    rt = make_gpu(elev, 64, 48, debug_hits=False, light_pos=sun_at_phase(60.0))
    fired = []
    rt.set_accum_done_cb(lambda r: fired.append(1))
    ref = rt.render_cycle().copy()
    t0 = rt.submit_frame()
    t1 = rt.submit_frame()
    a = rt.wait_frame(t0).copy()
    b = rt.wait_frame(t1).copy()
    t2 = rt.submit_frame()                                   # third submit reuses the first slot
    c = rt.wait_frame(t2).copy()
    assert (t0, t1, t2) == (0, 1, 0)
    assert np.array_equal(a, ref) and np.array_equal(b, ref) and np.array_equal(c, ref)
    assert len(fired) == 4
    with pytest.raises(ValueError):
        rt.submit_frame(np.zeros((3, 3, 4), np.uint8))
    rt.close()


def test_pipelined_frames_with_a_busy_referee_and_synchronous_calls_in_between():
    """Frames with a long referee phase (nearly every ray goes through it: long_walk = 2, tiny budgets) queued back to
    back, with synchronous calls (render_cycle, counters, hit read-back, texture updates) thrown in between."""
    from moonrtx_b200 import scene
    from moonrtx_b200.synth import synth_ephemeris, synth_ldem
    from moonrtx_b200.data_loader import downscale_elevation
    from moonrtx_b200.video import apply_frame_state
    elev, _ = downscale_elevation(synth_ldem(1440, 720, seed=11, craters=40), 2)
    W, H, n = 200, 120, 6
    states = [scene.frame_state(synth_ephemeris(900.0 * i)) for i in range(n)]

    def new_rt():
        rt = make_gpu(elev, W, H, debug_hits=False, light_pos=sun_at_phase(90.0))
        rt.set_param(min_accumulation_step=1, max_accumulation_frames=3)
        rt.set_uint("long_walk", 2)
        rt.set_uint("referee_budget", 8)
        return rt

    rt = new_rt()
    ref = []
    for i in range(n):
        apply_frame_state(rt, states[i])
        ref.append(rt.render_cycle().copy())
    rt.close()

    rt = new_rt()
    got = {}
    pending = None
    for i in range(n):
        apply_frame_state(rt, states[i])
        t = rt.submit_frame()
        if pending is not None:
            got[pending[0]] = rt.wait_frame(pending[1]).copy()
        pending = (i, t)
        if i == 2:                                           # synchronous work in the middle of the pipeline
            got[pending[0]] = rt.wait_frame(pending[1]).copy(); pending = None
            mid = rt.render_cycle().copy()
            assert np.array_equal(mid, ref[2])
            assert rt.counters()["primary_rays"] > 0
            rt._get_hit_at(W // 2, H // 2)
            rt.set_texture_2d("moon_color", np.full((8, 16, 4), 200, np.uint8))
            rt.set_texture_2d("moon_color", np.full((8, 16, 4), 255, np.uint8))
    got[pending[0]] = rt.wait_frame(pending[1]).copy()
    rt.close()
    # (frames 3.. were rendered with the white 8 x 16 albedo texture set in between: same as no texture)
    for i in range(n):
        assert np.array_equal(got[i], ref[i]), f"frame {i} differs"
