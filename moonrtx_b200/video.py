"""
Frame-parallel time-lapse: the reference's F11 export
(moonrtx/renderer_video.py:148-364) renders frame i at t0 + i*step strictly one after
another on one GPU, although only a few scalars change between frames
(moon_renderer.py:840-860).  Here frame i goes to rank i mod G (one process per GPU, the
height field replicated); no collective is needed while rendering, frames come back in
order through per-rank host buffers.

`FrameSink` stands in for PlotOptiX's NVENC encoder (encoder_create / encoder_start /
encoder_is_open / encoder_stop): B200 has no NVENC and this OpenCV build has no H.264,
so frames go to an `mp4v` cv2.VideoWriter.
"""

from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from .scene import FrameState


class FrameSink:
    def __init__(self, path: str, width: int, height: int, fps: int, n_frames: int = 0):
        import cv2
        self.path, self.n_frames, self.count = path, int(n_frames), 0
        self._w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (int(width), int(height)))
        if not self._w.isOpened():
            self._w = None
            raise RuntimeError(f"could not open a video writer for {path}")

    @staticmethod
    def check_available():
        import cv2  # noqa: F401

    def is_open(self) -> bool:
        return self._w is not None

    def grab(self, rgba: np.ndarray):
        """One frame per completed accumulation cycle; closes itself after n_frames
        (renderer_video.py:154-159)."""
        if self._w is None:
            return
        self._w.write(np.ascontiguousarray(rgba[..., 2::-1]))
        self.count += 1
        if self.n_frames and self.count >= self.n_frames:
            self.stop()

    def stop(self):
        if self._w is not None:
            self._w.release()
            self._w = None


def frames_of_rank(n_frames: int, rank: int, world: int) -> list[int]:
    """Frame i is rendered by rank i mod world (round-robin keeps every rank's frames
    spread over the whole terminator sweep, so per-rank cost is balanced)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the world")
    return list(range(rank, n_frames, world))


def apply_frame_state(rt, st: FrameState, moon_name: str = "moon", light_name: str = "sun",
                      camera_name: str = "cam1"):
    """The rt.* calls of update_view (moon_renderer.py:852-860) for one time step."""
    with rt._padlock:
        rt.update_camera(camera_name, eye=st.eye, target=st.target, up=st.up, fov=st.fov)
        rt.update_data(moon_name, u=st.u, v=st.v)
        rt.update_light(light_name, pos=st.light_pos, radius=st.light_radius)


def render_timelapse(rt, states: Sequence[FrameState], rank: int = 0, world: int = 1,
                     on_frame: Optional[Callable[[int, np.ndarray], None]] = None,
                     overlay_for: Optional[Callable[[int], Optional[np.ndarray]]] = None,
                     keep: bool = True, pipelined: bool = False) -> dict[int, np.ndarray]:
    """
    Render this rank's share of a time-lapse.  Each frame is one full accumulation cycle
    (rt.set_param(max_accumulation_frames=...) applies), exactly as the reference lets every
    frame converge before the encoder grabs it.  Returns {frame index: RGBA8 image}.

    pipelined=True: frame i + 1 is submitted before frame i is waited for (B200OptiX.submit_frame /
    wait_frame), so the overlay upload and the frame read-back overlap the tracing; same pixels.
    """
    out: dict[int, np.ndarray] = {}
    if pipelined:
        pending = None                                      # (frame index, ticket)
        def finish(p):
            img = rt.wait_frame(p[1])
            if on_frame is not None:
                on_frame(p[0], img)
            if keep:
                out[p[0]] = img.copy()
        for i in frames_of_rank(len(states), rank, world):
            apply_frame_state(rt, states[i])
            ticket = rt.submit_frame(overlay_for(i) if overlay_for is not None else None)
            if pending is not None:
                finish(pending)
            pending = (i, ticket)
        if pending is not None:
            finish(pending)
        return out
    for i in frames_of_rank(len(states), rank, world):
        if overlay_for is not None:
            ov = overlay_for(i)
            if ov is not None:
                rt.set_texture_2d("frame_overlay", ov, filter_mode="Nearest", refresh=False)
        apply_frame_state(rt, states[i])
        img = rt.render_cycle()
        if on_frame is not None:
            on_frame(i, img)
        if keep:
            out[i] = img.copy()
    return out


def render_timelapse_delivered(rt, states: Sequence[FrameState], rank: int, world: int, consumer: int = 0,
                               on_frame: Optional[Callable[[int, np.ndarray], None]] = None,
                               overlay_for: Optional[Callable[[int], Optional[np.ndarray]]] = None) -> int:
    """
    The F11 export across GPUs with ONE consumer (renderer_video.py:276-364 feeds one encoder, in frame order): frame i is
    rendered by rank i mod world; every rank but the consumer sends its resolved RGBA8 frames to the consumer over
    NVLink (B200OptiX.submit_frame(dst=consumer): ncclSend from the device, nothing pickled, nothing through the
    producer's host memory); the consumer renders its own share, posts the receives in frame order and hands every frame
    to on_frame(i, img) strictly in order.  Two frames are in flight per rank and two receives pending on the consumer,
    so tracing, NVLink transfers and the consumer's device-to-host copies overlap.  rt.comm_init() must have been called
    when world > 1.  Returns the number of frames this rank consumed (n on the consumer, 0 elsewhere).
    """
    n = len(states)
    own = frames_of_rank(n, rank, world)
    ov = (lambda i: overlay_for(i)) if overlay_for is not None else (lambda i: None)
    if rank != consumer:
        pending = None
        for i in own:
            apply_frame_state(rt, states[i])
            t = rt.submit_frame(ov(i), dst=consumer)
            if pending is not None:
                rt.wait_frame(pending)
            pending = t
        if pending is not None:
            rt.wait_frame(pending)
        return 0
    own_q, own_next = [], 0                  # (frame, ticket) submitted and not yet consumed; next index into own
    remote = [i for i in range(n) if i % world != rank]
    rem_q, rem_next = [], 0
    consumed = 0
    for i in range(n):
        while own_next < len(own) and len(own_q) < 2:
            f = own[own_next]
            apply_frame_state(rt, states[f])
            own_q.append((f, rt.submit_frame(ov(f))))
            own_next += 1
        while rem_next < len(remote) and len(rem_q) < 2:
            f = remote[rem_next]
            rem_q.append((f, rt.recv_frame(f % world)))
            rem_next += 1
        if i % world == rank:
            f, t = own_q.pop(0)
            img = rt.wait_frame(t)
        else:
            f, t = rem_q.pop(0)
            img = rt.wait_recv(t)
        assert f == i
        if on_frame is not None:
            on_frame(i, img)
        consumed += 1
    return consumed


def merge_in_order(per_rank: Iterable[dict[int, np.ndarray]], n_frames: int) -> list[np.ndarray]:
    """Frames of all ranks back in time order (what feeds the encoder on rank 0)."""
    merged: dict[int, np.ndarray] = {}
    for d in per_rank:
        for i, img in d.items():
            if i in merged:
                raise ValueError(f"frame {i} rendered twice")
            merged[i] = img
    missing = [i for i in range(n_frames) if i not in merged]
    if missing:
        raise ValueError(f"frames missing: {missing[:8]}")
    return [merged[i] for i in range(n_frames)]


def gather_frames(local: dict[int, np.ndarray], n_frames: int, rank: int, world: int, group=None):
    """
    Collect every rank's frames on rank 0, in time order (the encoder feed of the reference's
    export loop).  Uses torch.distributed's object gather on whatever backend the group has
    (gloo on CPU-only boxes); frames are host arrays at this point, 33 MB each at 4K.
    Returns the ordered list on rank 0, None elsewhere.
    """
    if world == 1:
        return merge_in_order([local], n_frames)
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    return merge_in_order(gathered, n_frames)
