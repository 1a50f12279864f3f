"""
Worker of the 2-GPU checks (launched by tests/test_multigpu_gpu.py through torch.distributed.run,
one process per GPU): a frame sharded by samples (ncclAllReduce of accumulators) and by interleaved
row bands (ncclAllGather of RGBA8 bands) must reproduce the single-GPU frame.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist          # noqa: E402
from helpers import make_gpu, sun_at_phase  # noqa: E402
from moonrtx_b200.optix import B200OptiX   # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    from moonrtx_b200.synth import synth_ldem
    from oracle import downscale_oracle as orc
    elev, _ = orc.load_elevation(synth_ldem(720, 360, seed=21, craters=60), 1)
    kw = dict(light_pos=sun_at_phase(80.0))
    W, H, N = 200, 150, 8
    rt = make_gpu(elev, W, H, debug_hits=False, **kw)
    rt.set_param(max_accumulation_frames=N, min_accumulation_step=N)
    uid = [B200OptiX.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    rt.comm_init(rank, world, uid[0])

    ref = rt.render_cycle().copy()                      # every rank: the whole frame on its own GPU
    ref_acc = rt.get_accum_buffer()

    img_s = rt.render_cycle(shard="samples").copy()
    acc_s = rt.get_accum_buffer()
    assert np.all(acc_s[..., 3] == N), "sample counts must add up"
    assert np.allclose(acc_s, ref_acc, rtol=1e-5, atol=1e-6), "sample-split accumulators"
    assert np.abs(img_s.astype(int) - ref.astype(int)).max() <= 1, "sample-split image"

    for tile in (16, 64):
        img_t = rt.render_cycle(shard="tiles", tile=tile).copy()
        assert np.array_equal(img_t, ref), f"interleaved tile split (tile={tile})"

    for tile_rows in (16, 64, 7):
        img_r = rt.render_cycle(shard="rows", tile_rows=tile_rows).copy()
        assert np.array_equal(img_r, ref), f"row-band split (tile_rows={tile_rows})"

    # frame-parallel time-lapse: frame i on rank i mod world, gathered in order on rank 0
    from moonrtx_b200 import scene
    from moonrtx_b200.synth import synth_ephemeris
    from moonrtx_b200.video import gather_frames, render_timelapse
    rt.set_param(max_accumulation_frames=1, min_accumulation_step=1)
    states = [scene.frame_state(synth_ephemeris(600.0 * i)) for i in range(5)]
    mine = render_timelapse(rt, states, rank, world)
    frames = gather_frames(mine, len(states), rank, world)
    if rank == 0:
        assert len(frames) == 5
        solo = render_timelapse(rt, states, 0, 1)
        for i in range(5):
            assert np.array_equal(frames[i], solo[i]), f"time-lapse frame {i}"
    # the same time-lapse with ordered delivery over NVLink (ncclSend / ncclRecv through the C ABI): rank 0 is handed
    # every frame, in order, with its overlay burnt in (renderer_video.py:137-144, 276-364)
    from moonrtx_b200.video import render_timelapse_delivered
    states = [scene.frame_state(synth_ephemeris(600.0 * i)) for i in range(7)]

    def overlay(i):
        ov = np.zeros((H, W, 4), np.uint8)
        ov[4:12, 4:4 + 8 * (i + 1)] = (255, 255, 255, 200)
        return ov
    solo = render_timelapse(rt, states, 0, 1, overlay_for=overlay) if rank == 0 else None
    for transport in ("nccl", "p2p"):
        if transport == "p2p":
            # peer-memory mailboxes (CUDA IPC + copy engines + stream memory operations): no kernel on either side
            handles = [None] * world
            dist.all_gather_object(handles, rt.p2p_open(rank, world))
            rt.p2p_connect(handles)
        for rep in range(2):                      # (twice: slots and sequence numbers carry on from one export to the next)
            got = []
            n = render_timelapse_delivered(rt, states, rank, world, consumer=0, on_frame=lambda i, img: got.append((i, img.copy())),
                                           overlay_for=overlay)
            if rank == 0:
                assert n == 7 and [i for i, _ in got] == list(range(7)), f"{transport}: frames must arrive in order"
                for i, img in got:
                    assert np.array_equal(img, solo[i]), f"{transport}: delivered frame {i}"
                    assert img[8, 6, 0] > 150 and (i == 6 or np.array_equal(img[8, 4 + 8 * (i + 1) + 2], solo[i][8, 4 + 8 * (i + 1) + 2]))
            else:
                assert n == 0
            dist.barrier()
    rt.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
