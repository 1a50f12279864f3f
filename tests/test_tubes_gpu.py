"""
Overlay tubes (rt.set_graph / update_graph / delete_geometry - the selenographic grid, its labels, the
feature labels and the pins, renderer_labels.py:263-305, renderer_pins.py:18-55) on the CUDA path against the float64
oracle, which tests every capsule by brute force: same pixels covered, same depth, same image; shadows untouched.
"""
import math

import numpy as np
import pytest

from helpers import image_metrics, make_gpu, make_oracle, sun_at_phase
from test_tubes_host import grid_graph, pin_graph

pytestmark = pytest.mark.gpu


def synth_elevation(W, H, seed=3):
    from moonrtx_b200.synth import synth_ldem
    from moonrtx_b200.data_loader import downscale_elevation
    return downscale_elevation(synth_ldem(W, H, seed=seed, craters=60), 1)


def _set_overlays(rt):
    pos, edges = grid_graph()
    rt.update_material("grid_material", {"flat": True})
    rt.set_graph("moon_grid_lines", pos=pos, edges=edges, r=0.006, c=[0.5, 0.5, 0.5], mat="grid_material")
    ppos, pedges = pin_graph(lat=12.0, lon=-20.0)
    rt.set_graph("pin_3", pos=ppos, edges=pedges, r=0.012, c=[1.0, 0.0, 0.0], mat="pin_material")
    # a label hidden on the night side: per-vertex radii, all zero (renderer_labels.py:128-130)
    rt.set_graph("hidden", pos=ppos + 0.3, edges=pedges, r=np.zeros(len(ppos), np.float32), c=[0.0, 1.0, 0.0])


@pytest.mark.parametrize("spp,fov,phase", [(1, 4.242192793, 80.0), (4, 4.242192793, 100.0), (1, 1.2, 60.0)])
def test_tubes_match_oracle(spp, fov, phase):
    elev, _ = synth_elevation(720, 360)
    kw = dict(light_pos=sun_at_phase(phase), fov=fov)
    W, H = 240, 180
    rt = make_gpu(elev, W, H, debug_hits=(spp == 1), **kw)
    _set_overlays(rt)
    seg = rt._tube_segments()
    assert len(seg) == 9 * 99 + 24 * 99 + 4 and np.all(seg[:, 3] > 0)         # the hidden label contributes nothing
    orc = make_oracle(elev, W, H, jitter=spp > 1, tubes=seg, **kw)
    if spp > 1:
        rt.set_param(max_accumulation_frames=spp, min_accumulation_step=spp)
    img = rt.render_cycle().copy()
    o = orc.render(nsamples=spp)
    ref = orc.tonemap(o["accum"])
    mae, psnr = image_metrics(img, ref)
    diff = np.abs(img[..., :3].astype(int) - ref[..., :3].astype(int)).max(axis=2)
    assert mae <= 0.05 and psnr >= 45.0 and int((diff > 2).sum()) <= 3, (mae, psnr, int((diff > 2).sum()))
    # the grid is there: grey pixels on the night side, where the surface is black, and red ones on the pin
    grey = (np.abs(img[..., 0].astype(int) - img[..., 1]) <= 1) & (img[..., 0] > 150)
    assert int(grey.sum()) > 200
    if spp == 1:
        assert int(((img[..., 0] > 200) & (img[..., 1] < 30) & (img[..., 2] < 30)).sum()) >= (2 if fov > 2 else 20)
    if spp == 1:
        g, oh = rt.get_hit_records_f64(), o["hit64"]
        tube_g, tube_o = g[..., 0] == -2.0, oh[..., 0] == -2.0
        assert int(tube_o.sum()) > 300
        assert int((tube_g != tube_o).sum()) == 0, int((tube_g != tube_o).sum())
        assert np.allclose(g[..., 3][tube_o], oh[..., 3][tube_o], rtol=0, atol=1e-6)      # distance to the tube (1e-7 R)
        # the hit buffer (rt._get_hit_at) holds the point on the tube: callers accept 0.9 R <= |h| <= 1.15 R as "on the Moon"
        hb, ob = rt.get_hit_buffer(), o["hit32"]
        assert np.allclose(hb[tube_o], ob[tube_o], rtol=0, atol=2e-4)
        rr = np.linalg.norm(hb[tube_o][:, :3], axis=1)
        assert rr.min() > 9.9 and rr.max() < 10.1
        # tubes beyond the limb: pixels whose rays miss the bounding sphere altogether
        yy, xx = np.nonzero(tube_o)
        t = math.tan(math.radians(fov) / 2)
        sx = ((xx + 0.5) / W * 2 - 1) * t * W / H
        sy = (1 - (yy + 0.5) / H * 2) * t
        b = 300.0 * np.sqrt(sx * sx + sy * sy) / np.sqrt(1 + sx * sx + sy * sy)
        if fov > 2:
            assert int((b > 10.0).sum()) >= 1
    rt.close()


def test_tubes_cast_no_shadow_and_can_be_removed():
    """The reference's overlay material lets shadow rays through (renderer_labels.py:133-139): with the tubes deleted the
    frame is the plain one again, and pixels not covered by a tube never change."""
    elev, _ = synth_elevation(720, 360)
    kw = dict(light_pos=sun_at_phase(75.0))
    rt = make_gpu(elev, 200, 150, **kw)
    plain = rt.render_cycle().copy()
    acc_plain = rt.get_accum_buffer().copy()
    _set_overlays(rt)
    with_tubes = rt.render_cycle().copy()
    covered = rt.get_hit_records_f64()[..., 0] == -2.0
    assert int(covered.sum()) > 100
    assert np.array_equal(rt.get_accum_buffer()[~covered], acc_plain[~covered])
    assert not np.array_equal(with_tubes[covered], plain[covered])
    rt.update_graph("moon_grid_lines", r=0.0)                     # show_moon_grid(False), renderer_labels.py:324
    rt.delete_geometry("pin_3")                                   # remove_pin, renderer_pins.py:70
    assert len(rt._tube_segments()) == 0
    assert np.array_equal(rt.render_cycle(), plain)
    rt.close()


def test_many_segments_in_one_tile_fall_back_to_the_whole_list():
    """More segments in a screen tile than its list holds: the tile tests every segment instead (same answer)."""
    elev = np.full((360, 720), 0.99, np.float32)      # (720 wide: a coarser map is rendered by the float64 A/B kernel, which draws no overlays)
    rng = np.random.default_rng(5)
    n = 300
    lat, lon = np.radians(rng.uniform(-3, 3, (n, 2))), np.radians(rng.uniform(-3, 3, (n, 2)))
    pts = np.stack([10.0 * np.cos(lat) * np.sin(lon), -10.0 * np.cos(lat) * np.cos(lon), 10.0 * np.sin(lat)], axis=2).reshape(-1, 3)
    edges = np.arange(2 * n).reshape(n, 2)
    rt = make_gpu(elev, 64, 64, fov=1.0)
    rt.set_graph("dense", pos=pts, edges=edges, r=0.004, c=[0.2, 0.9, 0.4])
    orc = make_oracle(elev, 64, 64, fov=1.0, tubes=rt._tube_segments())
    rt.render_cycle()
    g, oh = rt.get_hit_records_f64(), orc.render()["hit64"]
    assert int(((g[..., 0] == -2.0) != (oh[..., 0] == -2.0)).sum()) == 0 and int((oh[..., 0] == -2.0).sum()) > 80
    rt.close()
