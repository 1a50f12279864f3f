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
