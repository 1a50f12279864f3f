"""Host logic of the frame-parallel time-lapse, incl. a world_size-2 gloo run on CPU."""
import os
import sys

import numpy as np
import pytest

from moonrtx_b200 import scene
from moonrtx_b200.synth import synth_ephemeris
from moonrtx_b200.video import frames_of_rank, merge_in_order


def test_round_robin_partition_covers_every_frame_once():
    for n in (0, 1, 7, 240):
        for world in (1, 2, 3, 8):
            seen = sorted(i for r in range(world) for i in frames_of_rank(n, r, world))
            assert seen == list(range(n))
    assert frames_of_rank(10, 1, 4) == [1, 5, 9]
    with pytest.raises(ValueError):
        frames_of_rank(10, 4, 4)


def test_merge_in_order_detects_gaps_and_duplicates():
    a = {0: np.zeros(1), 2: np.ones(1) * 2}
    b = {1: np.ones(1)}
    out = merge_in_order([a, b], 3)
    assert [int(x[0]) for x in out] == [0, 1, 2]
    with pytest.raises(ValueError):
        merge_in_order([a], 3)
    with pytest.raises(ValueError):
        merge_in_order([a, {0: np.zeros(1), 1: np.ones(1)}], 3)


def test_frame_state_matches_reference_vectors(golden_dir):
    """scene.py against the reference's MoonRenderer (fixture from oracle/make_golden.py)."""
    import json
    with open(os.path.join(golden_dir, "scene_vectors.json")) as f:
        sv = json.load(f)
    K = sv["constants"]
    assert scene.CAMERA_DISTANCE == K["CAMERA_DISTANCE"] and scene.SUN_BRIGHTNESS_SCALE == K["SUN_BRIGHTNESS_SCALE"]
    assert scene.SCENE_EPSILON == K["SCENE_EPSILON"] and scene.ACCUMULATION_FRAMES == K["ACCUMULATION_FRAMES"]
    for c in sv["cases"]:
        eph = synth_ephemeris(0.0)._replace(distance=c["distance"], sun_distance=c["sun_distance"],
                                            phase_angle=c["phase_angle"], bright_limb_angle=c["bright_limb_angle"])
        st = scene.frame_state(eph)
        assert np.allclose(st.light_pos, c["light_pos"], rtol=0, atol=1e-9)
        assert st.light_radius == pytest.approx(c["sun_light_radius"], rel=1e-15)
        assert np.allclose(st.eye, c["camera_eye"], rtol=0, atol=1e-12)
        assert st.fov == pytest.approx(c["camera_fov"], rel=1e-15)
        assert scene.moon_camera_distance(c["distance"]) == pytest.approx(c["camera_distance"], rel=1e-15)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moonrtx_b200.video import gather_frames
    n = 9
    mine = {i: np.full((4, 6, 4), i, dtype=np.uint8) for i in frames_of_rank(n, rank, world)}
    frames = gather_frames(mine, n, rank, world)
    if rank == 0:
        q.put([int(f[0, 0, 0]) for f in frames])
    else:
        assert frames is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_frames_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    order = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert order == list(range(9))


class _FakeRT:
    """The drop-in's frame-delivery surface over gloo: what render_timelapse_delivered needs of an rt, with torch.distributed
    point-to-point in place of NVLink (two sends in flight per producer, two receives pending on the consumer)."""
    def __init__(self, rank):
        import threading
        self.rank, self._padlock = rank, threading.RLock()
        self.state, self.tickets, self.recvs, self.next_ticket, self.calls = None, {}, {}, 0, []
    def update_camera(self, *a, **k): pass
    def update_light(self, *a, **k): pass
    def update_data(self, name, u=None, v=None, **k): self.state = float(u[0])       # the frame's identity rides in u
    def _render(self, overlay):
        img = np.full((4, 6, 4), int(self.state) % 251, np.uint8)
        if overlay is not None:
            img[0, 0] = overlay[0, 0]
        return img
    def submit_frame(self, overlay=None, dst=None):
        import torch, torch.distributed as dist
        assert len(self.tickets) < 2, "more than two frames in flight"
        t = self.next_ticket; self.next_ticket ^= 1
        img = self._render(overlay)
        self.tickets[t] = img if dst is None else dist.isend(torch.from_numpy(img.copy()), dst)
        return t
    def wait_frame(self, t):
        v = self.tickets.pop(t)
        if isinstance(v, np.ndarray):
            return v
        v.wait()
        return None
    def recv_frame(self, src):
        import torch, torch.distributed as dist
        assert len(self.recvs) < 2, "more than two receives pending"
        t = max(self.recvs, default=-1) + 1
        buf = torch.empty((4, 6, 4), dtype=torch.uint8)
        self.recvs[t] = (buf, dist.irecv(buf, src))
        return t
    def wait_recv(self, t):
        buf, req = self.recvs.pop(t)
        req.wait()
        return buf.numpy()


def _delivery_worker(rank, world, port, q):
    import torch.distributed as dist
    from collections import namedtuple
    from moonrtx_b200.video import render_timelapse_delivered
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    St = namedtuple("St", "eye target up fov u v light_pos light_radius")
    n = 11
    states = [St((0, -300, 0), (0, 0, 0), (0, 0, 1), 4.2, (float(i), 0, 1), (0, -1, 0), (1, 0, 0), 100.0) for i in range(n)]
    got = []
    rt = _FakeRT(rank)
    def overlay(i):
        ov = np.zeros((4, 6, 4), np.uint8); ov[0, 0] = (i, 2 * i, 3 * i, 255)
        return ov
    consumed = render_timelapse_delivered(rt, states, rank, world, consumer=0, on_frame=lambda i, img: got.append((i, img.copy())), overlay_for=overlay)
    if rank == 0:
        ok = consumed == n and [i for i, _ in got] == list(range(n))
        ok = ok and all(int(img[1, 1, 0]) == i % 251 and tuple(img[0, 0]) == (i, 2 * i, 3 * i, 255) for i, img in got)
        q.put(bool(ok))
    else:
        assert consumed == 0 and not got
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_delivered_timelapse_hands_rank0_every_frame_in_order_gloo(world):
    """renderer_video.py:276-364 feeds ONE encoder in frame order: frame i is rendered by rank i mod world and reaches the
    consumer's on_frame(i, img) in order, with its own overlay, never more than two frames in flight per rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 150 + world
    procs = [ctx.Process(target=_delivery_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
