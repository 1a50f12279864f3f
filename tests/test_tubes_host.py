"""
CPU side of the overlay tubes: the oracle's capsule test on analytic cases, and the flattening of
rt.set_graph geometry into the segment list of mrtx_set_tubes - checked on graphs built by the UNMODIFIED reference
(moon_grid.create_moon_grid / merge_segments_to_graph) when /root/reference is present.
"""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import make_oracle  # noqa: E402


def sphere_point(lat_deg, lon_deg, r=10.0):
    la, lo = np.radians(lat_deg), np.radians(lon_deg)
    return np.stack([r * np.cos(la) * np.sin(lo), -r * np.cos(la) * np.cos(lo), r * np.sin(la)], axis=-1)


def grid_graph(r=10.0, step=15.0, n=100):
    """lines of the selenographic grid as moon_grid.create_moon_grid lays them out (moon_grid.py:721-748): parallels
    -60..60, meridians 0..345, n points each, on the sphere of radius r; merged like merge_segments_to_graph (:27-46)"""
    lines = []
    for lat in np.arange(-60, 61, step):
        lines.append(sphere_point(np.full(n, lat), np.linspace(0, 360, n), r))
    for lon in np.arange(0, 360, step):
        lines.append(sphere_point(np.linspace(-90, 90, n), np.full(n, lon), r))
    pos = np.concatenate(lines, axis=0)
    edges, off = [], 0
    for ln in lines:
        idx = np.arange(off, off + len(ln))
        edges.append(np.column_stack((idx[:-1], idx[1:])))
        off += len(ln)
    return pos, np.concatenate(edges, axis=0)


def pin_graph(lat, lon, r=10.05):
    """a small square of four strokes 0.5 % above the sphere (the reference's glyph strokes, moon_grid.py:188)"""
    c = [sphere_point(lat + a, lon + b, r) for a, b in ((0, 0), (0, 2), (2, 2), (2, 0))]
    pos = np.array([c[0], c[1], c[1], c[2], c[2], c[3], c[3], c[0]])
    return pos, np.arange(8).reshape(4, 2)


def _seg(a, b, r, col):
    s = np.zeros((1, 12), np.float32)
    s[0, 0:3], s[0, 3], s[0, 4:7], s[0, 8:11] = a, r, b, col
    return s


def test_oracle_tube_in_front_of_the_sphere_is_analytic():
    elev = np.full((45, 90), 0.9, np.float32)                      # surface radius 9: the tube at 10 floats above it
    seg = _seg((-1.0, -10.0, 0.0), (1.0, -10.0, 0.0), 0.05, (0.25, 0.5, 0.75))
    orc = make_oracle(elev, 101, 101, shadows=False, tubes=seg)
    o = orc.render()
    h = o["hit64"]
    tube = h[..., 0] == -2.0
    # the centre pixel looks down the -y axis: it meets the tube's surface at y = -10.05
    assert tube[50, 50] and abs(h[50, 50, 3] - (300.0 - 10.05)) < 1e-9
    assert np.allclose(o["accum"][50, 50, :3], (0.25, 0.5, 0.75)) and np.allclose(o["hit32"][50, 50], (0, -10.05, 0, 289.95), atol=1e-4)
    # covered pixels: a band 2.1 wide and 0.1 high around the centre (pixel pitch of this camera: 0.22 at the tube)
    pitch = 2 * 290.0 * math.tan(math.radians(4.242192793) / 2) / 101
    yy, xx = np.nonzero(tube)
    assert set(yy) == {50} and abs(len(xx) - 2.1 / pitch) <= 1.5
    # behind the sphere nothing shows: the same tube on the far side
    far = _seg((-1.0, 10.0, 0.0), (1.0, 10.0, 0.0), 0.05, (1, 1, 1))
    o2 = make_oracle(elev, 101, 101, shadows=False, tubes=far).render()
    assert not (o2["hit64"][..., 0] == -2.0).any()
    # ... unless it sticks out beyond the limb
    far[0, 0:3], far[0, 4:7] = (9.5, 10.0, 0.0), (11.5, 10.0, 0.0)
    o3 = make_oracle(elev, 101, 101, shadows=False, tubes=far).render()
    t3 = o3["hit64"][..., 0] == -2.0
    assert t3.any() and np.nonzero(t3)[1].min() > 50 + 9.0 / (2 * 310.0 * math.tan(math.radians(4.242192793) / 2) / 101) - 1


def test_oracle_ray_along_the_tube_axis_meets_the_end_cap():
    elev = np.full((45, 90), 0.5, np.float32)
    seg = _seg((0.0, -12.0, 0.0), (0.0, -11.0, 0.0), 0.2, (1, 0, 0))
    h = make_oracle(elev, 3, 3, fov=0.01, shadows=False, tubes=seg).render()["hit64"]
    assert h[1, 1, 0] == -2.0 and abs(h[1, 1, 3] - (300.0 - 12.2)) < 1e-6


def test_reference_grid_flattens_to_the_segment_list():
    from oracle import ref_stub
    if not ref_stub.reference_available():
        pytest.skip("reference tree not on this box")
    mg = ref_stub.import_reference("moon_grid")
    create_moon_grid, merge_segments_to_graph = mg.create_moon_grid, mg.merge_segments_to_graph
    g = create_moon_grid(moon_radius=10.0, lat_step=15.0, lon_step=15.0, points_per_line=100, offset=0.0)
    pos, edges = merge_segments_to_graph(g.lat_lines + g.lon_lines)
    mine_pos, mine_edges = grid_graph()
    assert np.allclose(pos, mine_pos, atol=1e-9) and np.array_equal(edges, mine_edges)     # the test grid IS the reference's
    from moonrtx_b200.optix import tube_segments
    seg = tube_segments({"grid": {"geom": "Graph", "pos": pos, "edges": edges, "r": 0.006, "c": [0.5, 0.5, 0.5]},
                         "labels": {"geom": "Graph", "pos": pos[:10], "edges": edges[:9],
                                    "r": np.repeat([0.012, 0.0], 5).astype(np.float32), "c": [1.0, 0.0, 0.0]},
                         "moon": {"geom": "ParticleSetTextured"}})
    assert seg.shape == (len(edges) + 4, 12) and seg.dtype == np.float32
    assert np.allclose(seg[:len(edges), 0:3], pos[edges[:, 0]], atol=1e-6) and np.allclose(seg[:len(edges), 4:7], pos[edges[:, 1]], atol=1e-6)
    assert np.all(seg[:len(edges), 3] == np.float32(0.006)) and np.all(seg[len(edges):, 3] == np.float32(0.012))
    assert np.allclose(seg[len(edges):, 8:11], (1, 0, 0))


def test_tube_segments_flattening_rules():
    """rt.set_graph arguments -> the segment list: one radius or one per vertex (0 hides, the smaller end wins), one colour, a
    grey level or one colour per vertex; geometry that is no graph, or has no edges, contributes nothing."""
    from moonrtx_b200.optix import tube_segments
    pos = np.array([[0, 0, 10.0], [1, 0, 10.0], [2, 0, 10.0], [3, 0, 10.0]])
    edges = np.array([[0, 1], [1, 2], [2, 3]])
    seg = tube_segments({"a": {"geom": "Graph", "pos": pos, "edges": edges, "r": 0.02, "c": 0.5}})
    assert seg.shape == (3, 12) and np.all(seg[:, 3] == np.float32(0.02)) and np.allclose(seg[:, 8:11], 0.5)
    assert np.allclose(seg[1, 0:3], pos[1]) and np.allclose(seg[1, 4:7], pos[2]) and np.all(seg[:, [7, 11]] == 0)
    seg = tube_segments({"a": {"geom": "Graph", "pos": pos, "edges": edges, "r": [0.02, 0.02, 0.0, 0.01],
                               "c": [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]]}})
    assert seg.shape == (1, 12) and np.allclose(seg[0, 8:11], (1, 0, 0))                 # edges touching the hidden vertex are gone
    assert len(tube_segments({"m": {"geom": "ParticleSetTextured"}, "e": {"geom": "Graph", "pos": pos, "edges": np.zeros((0, 2), int)},
                              "n": {"geom": "Graph", "pos": None, "edges": None}})) == 0
    with pytest.raises(ValueError):
        tube_segments({"a": {"geom": "Graph", "pos": pos, "edges": edges, "r": [0.1, 0.2]}})
    two = tube_segments({"a": {"geom": "Graph", "pos": pos, "edges": edges[:1], "r": 0.01, "c": [1, 0, 0]},
                         "b": {"geom": "Graph", "pos": pos, "edges": edges[1:], "r": 0.03, "c": [0, 0, 1]}})
    assert two.shape == (3, 12) and sorted(np.unique(two[:, 3]).tolist()) == [np.float32(0.01), np.float32(0.03)]
