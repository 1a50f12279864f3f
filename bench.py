#!/usr/bin/env python
"""
bench.py - the headline benchmark of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (mode frames)
    python bench.py --mode samples|tiles ...                      # the two single-frame sharding modes
    python bench.py --impl reference [--gpus N] ...               # CPU arm (oracle port, all host threads)

Workload of the default mode (`config.workload`): BASELINE config 3 frames along the config-4 sweep.  One step = 8
consecutive 3840x2160 frames of the F11 time-lapse (10-minute steps through a terminator sweep, default whole-disk camera),
each rendered from the full-size 92160x46080 int16 LOLA-shaped synthetic height map (8.5 GB, generated in HBM) and a colour
texture with 16 spp (primary ray + sun shadow ray + Lambert shading per sample) and resolved (Gamma / Overlay).  Frame f
is rendered by rank f mod N: the work of a step does not depend on N (strong scaling; --steps 30 is config 4's 240 frames).

Metric: Mrays/s = rays that walk the pyramid (primary rays that enter the bounding sphere + shadow rays) per second;
`value_incl_culled` adds the primary rays of pixels the cull pass rejects with one sphere test (SURVEY.md 8d counts them).
`value`: frames rendered back to back with everything resident in HBM, CUDA events on the launching stream, max over
ranks.  `e2e`: the same frames through the public drop-in API as the export loop uses it - every frame's overlay copied
to pinned memory and uploaded, scene update, accumulation cycle, resolve, and the RGBA8 frame DELIVERED IN FRAME ORDER
to rank 0's pinned ring (its own frames device -> host, the other ranks' frames ncclSend -> rank 0 -> host) where a
consumer reads it - wall clock between barriers.
"""

import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAP_W, MAP_H = 92160, 46080
IMG_W, IMG_H = 3840, 2160
COLOR_W, COLOR_H = 27360, 13680
COLOR_K = 4
SEED = 20240314
FRAME_STEP_MIN = 10.0            # config 4: 10-minute steps through the terminator sweep
FRAMES_PER_STEP = 8              # frames of the sweep per step (divisible by every N the driver uses)
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback
SCALE_F32 = float(np.float32(0.5 / 1737400.0))


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: libraries that print there (NCCL's version banner does, whatever
    NCCL_DEBUG_FILE says) are sent to stderr instead; emit() writes to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def quadtree_depth(W):
    k = 0
    while (W >> (k + 1)) >= 64:
        k += 1
    return k


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    nvidia-smi takes a while to come up (seconds on an 8-GPU box) and its start-up stalls CUDA calls of the
    processes it attaches to, so ONE sampler (rank 0, all visible GPUs) is started before the warm-up steps;
    only the rows that arrive between begin() and end() are used."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, enabled=True, gpus=(0,), period_ms=50):
        self.rows, self.proc, self.t0, self.t1 = [], None, None, None
        if not enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(g) for g in gpus), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits",
                                          "-lms", str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_ready(self, timeout=20.0):
        t = time.time()
        while self.proc and not self.rows and time.time() - t < timeout:
            time.sleep(0.05)

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, r in self.rows:
            if len(r) < 6 or self.t0 is None or not (self.t0 <= ts <= (self.t1 or ts) + 0.05):
                continue
            try:
                sm.append(float(r[0])); mx = max(mx or 0.0, float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no nvidia-smi sample inside the timed region"], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}



def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def sweep_state(frame):
    from moonrtx_b200 import scene
    from moonrtx_b200.synth import synth_ephemeris
    return scene.frame_state(synth_ephemeris(frame * FRAME_STEP_MIN))


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the primary-ray kernel (trace_kernel_pool; captures older
    than it hold trace_kernel_fast), from the newest profiles/*_raw.csv (`ncu --set full` of this workload; the csv is the raw page of the report)."""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_raw.csv"))):
        try:
            import csv
            rows = list(csv.reader(open(path)))
            hdr = rows[0]
            ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            units = rows[1]
            mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for r in rows[2:]:
                if "trace_kernel_pool" in r[ik] or "trace_kernel_fast" in r[ik]:
                    val = float(r[ir].replace(",", "")) * mul.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * mul.get(units[iw], 1.0)
                    best = (val, os.path.basename(path))
        except Exception:
            continue
    return best


# ------------------------------------------------------------------------------------------------
def build_scene(args, local_rank):
    """Synthetic maps in HBM + the drop-in renderer, set up with the reference's call sequence."""
    from moonrtx_b200 import _lib, scene
    from moonrtx_b200.optix import B200OptiX
    from moonrtx_b200.data_loader import albedo_lut

    rt = B200OptiX(width=args.img_w, height=args.img_h)
    dev = rt._dev
    W, H = args.map_w, args.map_h
    ldem = dev.alloc(W * H * 2)
    _lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, ldem.ptr, W, H, SEED))
    # radius_scale exactly as load_elevation_data(ds=1) would return it (data_loader.py:232-242)
    radius_scale = _radius_scale_ds1(dev, ldem, W, H)
    # colour map: synth BGR in HBM -> reduce + LUT kernel -> RGBA texture (data_loader.py:290-368)
    cw, ch, k = args.color_w, args.color_h, COLOR_K
    bgr = dev.alloc(cw * ch * 3)
    _lib.check(dev.lib.mrtx_synth_color_bgr_dev(dev.ctx, bgr.ptr, cw, ch, 4720))
    tex_dev = dev.alloc((cw // k) * (ch // k) * 4)
    lut = albedo_lut(2.2)
    _lib.check(dev.lib.mrtx_color_reduce_lut_dev(dev.ctx, bgr.ptr, cw, ch, k, lut.ctypes.data, tex_dev.ptr))
    tex = tex_dev.download((ch // k, cw // k, 4), np.uint8)
    bgr.free(); tex_dev.free()

    rt.set_param(min_accumulation_step=args.spp, max_accumulation_frames=args.spp)
    # direct light: camera segment + light segment, the north_star path.  The reference's own setting
    # (2, 4) adds two diffuse interreflection bounces; the line's "interreflection" block times that
    rt.set_uint("path_seg_range", 2, 2)
    rt.set_float("scene_epsilon", scene.SCENE_EPSILON)
    rt.set_float("marching_step", scene.MARCHING_STEP)
    rt.set_float("marching_step_eps", scene.MARCHING_STEP_EPS)
    rt.set_ambient(0)
    rt.set_float("tonemap_exposure", scene.TONEMAP_EXPOSURE)
    rt.set_float("tonemap_gamma", 2.2)
    rt.add_postproc("Gamma")
    rt.set_background(0)
    rt.set_texture_2d("moon_color", tex)
    rt.update_material("diffuse", {"ColorTextures": ["moon_color"]})
    rt.set_data("moon", geom="ParticleSetTextured", geom_attr="DisplacedSurface",
                pos=[0, 0, 0], u=[0, 0, 1], v=[0, -1, 0], r=scene.MOON_RADIUS)
    rt.set_displacement_i16("moon", (ldem, W, H), radius_scale=radius_scale)
    rt.setup_camera("cam1", cam_type="Pinhole", eye=[0, -scene.CAMERA_DISTANCE, 0], target=[0, 0, 0], up=[0, 0, 1],
                    fov=scene.default_fov())
    rt.setup_light("sun", color=scene.light_radiance(80.0), radius=scene.SUN_RADIUS, in_geometry=False)
    rt.add_postproc("Overlay")
    rt._bench_texture = tex                 # (the oracle of the parity block samples the same albedo texture)
    return rt, ldem, radius_scale


def _radius_scale_ds1(dev, ldem, W, H, band=512):
    """max of fl32(fl32(c*scale)+1) over the whole map, i.e. the radius_scale the reference computes at
    downscale 1 (data_loader.py:218-220, 232, 241): the ds=1 kernel run band by band into a scratch
    buffer (the full float32 map would be 17 GB), keeping only each band's maximum."""
    import ctypes as C
    from moonrtx_b200 import _lib
    out = dev.alloc(W * band * 4)
    best = 0.0
    for r0 in range(0, H, band):
        rows = min(band, H - r0)
        rs = C.c_float()
        _lib.check(dev.lib.mrtx_downscale_i16_dev(dev.ctx, C.c_void_p(ldem.ptr + r0 * W * 2), W, rows, 1, out.ptr,
                                                  C.byref(rs)))
        best = max(best, float(rs.value))
    out.free()
    return best



def radius_scale_numpy(ldem_host):
    """data_loader.py:218-220, 232 at downscale 1: max of fl32(fl32(c * scale) + 1) - monotone in c, so from the max count"""
    m = np.float32(ldem_host.max())
    return float(np.float32(np.float32(m * np.float32(SCALE_F32)) + np.float32(1)))


def downscale_lines(dev, peak, skip_cpu):
    """BASELINE config 1 beside the headline (data_loader.py:215-242 on the GPU): --downscale 4 of a 23040x11520 int16 map.
    HBM to HBM (L2 flushed between runs, algorithmic bytes 2*W*H + 4*(W/4)*(H/4)), the same call with HOST
    buffers (531 MB up through pinned chunks overlapped with the kernel, 66 MB down), the reference's numpy expression
    (oracle port pinned bit for bit on data_loader.py:223-242) on one host core, and the full-resolution 92160x46080 map at
    the application's default factors 3 and 4."""
    from moonrtx_b200 import _lib
    from moonrtx_b200.data_loader import downscale_elevation, downscale_elevation_dev
    out_lines = {}
    W, H, ds = 23040, 11520, 4
    src = dev.alloc(W * H * 2)
    _lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, src.ptr, W, H, SEED))
    out = dev.alloc((W // ds) * (H // ds) * 4)

    def time_dev(src, W, H, ds, out, reps=10):
        for _ in range(3):
            downscale_elevation_dev(src, W, H, ds, out)
        ts = []
        for _ in range(reps):
            dev.l2_flush(); dev.synchronize()
            dev.timer_start()
            _, rs = downscale_elevation_dev(src, W, H, ds, out)
            ts.append(dev.timer_stop())
        return statistics.median(ts), rs
    nbytes = 2 * W * H + 4 * (W // ds) * (H // ds)
    ms, rs = time_dev(src, W, H, ds, out)
    line = {"workload": f"{W}x{H} int16 -> downscale {ds} (BASELINE config 1), HBM to HBM, L2 flushed between runs, radius_scale returned to the host",
            "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 4),
            "algorithmic_bytes": nbytes, "launches": 2}
    host = src.download((H, W), np.int16)
    # the drop-in's own entry point with host buffers (what load_elevation_data calls)
    downscale_elevation(host, ds)
    t0 = time.perf_counter()
    got, rs_host = downscale_elevation(host, ds)
    dt = time.perf_counter() - t0
    line["e2e_host_buffers"] = {"ms": round(dt * 1e3, 2), "GBps": round(nbytes / dt / 1e9, 2), "h2d_bytes": int(host.nbytes), "d2h_bytes": int(got.nbytes),
                                "api": "moonrtx_b200.data_loader.downscale_elevation(ndarray, 4)"}
    if not skip_cpu:
        from oracle import downscale_oracle as orc
        t0 = time.perf_counter()
        ref, rs_ref = orc.load_elevation(host, ds)
        dtc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": round(nbytes / dtc / 1e9, 3), "unit": "GB/s", "seconds": round(dtc, 2), "cores": 1, "kind": "port",
                                "sample": "the whole config-1 map, numpy reshape/mean/normalise as data_loader.py:223-242"}
        line["bit_exact_vs_oracle"] = bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)) and rs_host == rs_ref
                                           and np.array_equal(out.download((H // ds, W // ds), np.float32).view(np.uint32), ref.view(np.uint32)))
    src.free(); out.free()
    out_lines["config1"] = line
    return out_lines


def downscale_fullres(dev, ldem, peak):
    """the full-resolution map at the application's factors (main.py default 3; 4): output 1.9 / 1.06 GB, larger than L2"""
    from moonrtx_b200.data_loader import downscale_elevation_dev
    res = {}
    W, H = MAP_W, MAP_H
    for ds in (3, 4):
        out = dev.alloc((W // ds) * (H // ds) * 4)
        ts = []
        for i in range(4):
            dev.l2_flush(); dev.synchronize()
            dev.timer_start()
            downscale_elevation_dev(ldem, W, H, ds, out)
            t = dev.timer_stop()
            if i:
                ts.append(t)
        out.free()
        nbytes = 2 * W * H + 4 * (W // ds) * (H // ds)
        ms = statistics.median(ts)
        res[f"ds{ds}"] = {"ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 4),
                          "algorithmic_bytes": nbytes}
    return res


def overlay_image(h, w, text_seed):
    """A frame_overlay like renderer_video.py:106-144 draws (time label box, bottom-left)."""
    buf = np.zeros((h, w, 4), dtype=np.uint8)
    bh, bw = max(8, h // 40), max(64, w // 6)
    m = max(6, int(round(h * 0.015)))
    buf[h - m - bh:h - m, m:m + bw] = (0, 0, 0, 150)
    buf[h - m - bh + 4:h - m - 4, m + 8:m + 8 + (text_seed % (bw - 16))] = (255, 255, 255, 255)
    return buf




class Dist:
    """torch.distributed plumbing (barriers and scalar reductions only; the data path uses the library's own NCCL calls)"""
    def __init__(self, rank, world, local):
        self.rank, self.world = rank, world
        if world > 1:
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"      # NCCL logs to stdout by default, which carries exactly one JSON line
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            self.torch, self.dist = torch, dist

    def barrier(self, dev):
        dev.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.world > 1 else None)

    def comm_init(self, rt):
        """the library's own communicator (frame delivery, all-reduce, all-gather): unique id from rank 0"""
        if self.world == 1:
            return
        from moonrtx_b200.optix import B200OptiX
        uid = [B200OptiX.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(uid, src=0)
        rt.comm_init(self.rank, self.world, uid[0])

    def p2p_connect(self, rt):
        """frame delivery through peer memory (copy engines over NVLink): every rank's mailbox handle to every rank.
        Returns False - on every rank - if CUDA IPC is not available to any of them (delivery then uses NCCL)."""
        if self.world == 1:
            return True
        handle, ok = None, 1.0
        try:
            handle = rt.p2p_open(self.rank, self.world)
        except Exception as e:
            sys.stderr.write(f"rank {self.rank}: peer-memory mailbox not available ({e})\n")
            ok = 0.0
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle)
        if ok and all(h is not None for h in handles):
            try:
                rt.p2p_connect(handles)
            except Exception as e:
                sys.stderr.write(f"rank {self.rank}: cannot map the peers' mailboxes ({e})\n")
                ok = 0.0
        else:
            ok = 0.0
        t = self.torch.tensor([ok], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        if float(t.item()) < 1.0:
            try:
                rt.p2p_close()
            except Exception:
                pass
            return False
        return True

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def walked(c):
    return c["primary_in_sphere"] + c["shadow_rays"]


def launches_per_frame(args):
    """kernels of this library per frame: cull, then per sample chunk and pixel wave trace_kernel_pool + shade_kernel +
    shadow_kernel + trace_kernel_referee + referee_hard_kernel + its finish kernel, fold, resolve"""
    chunks = (args.spp + 31) // 32
    per_chunk = min(args.spp, 32)
    npix = args.img_w * args.img_h
    cap = min(npix * per_chunk, 1 << 26)
    waves = -(-npix // (cap // per_chunk))
    return 1 + 6 * chunks * waves + 1 + 1


# ------------------------------------------------------------------------------------------------
def parity_block(args, rt, ldem_host, radius_scale, frame, budget_s):
    """Oracle parity on the very frame the bench renders (and the cpu_baseline timing): the float64 oracle on a strided
    sub-grid of the frame, (a) 1 spp through the pixel centres: hit / miss decisions and hit radius against the GPU's
    float64 hit records, (b) the full jittered sample set: 8-bit image against the GPU frame.  Tolerances are north_star's."""
    from oracle.render_oracle import OracleScene
    from moonrtx_b200 import _lib, scene
    from moonrtx_b200.video import apply_frame_state
    st = sweep_state(frame)
    dev = rt._dev
    lib, ctx = dev.lib, dev.ctx
    cores = os.cpu_count() or 1
    kw = dict(scale=SCALE_F32, radius_scale=radius_scale, img_w=args.img_w, img_h=args.img_h, u=st.u, v=st.v, eye=st.eye,
              target=st.target, up=st.up, fov=st.fov, light_pos=st.light_pos, light_radius=st.light_radius,
              light_radiance=scene.light_radiance(80.0), texture=rt._bench_texture)
    sc1 = OracleScene(ldem_host, jitter=False, **kw)
    t0 = time.perf_counter()
    o = sc1.render(stride=96, nsamples=1)
    t_cal = time.perf_counter() - t0
    per_px = t_cal / (o["accum"].shape[0] * o["accum"].shape[1]) * (args.spp + 1)
    want = max(1, int(budget_s / max(per_px, 1e-9)))
    stride = max(1, int(np.ceil(np.sqrt(args.img_w * args.img_h / want))))
    # (a) 1 spp, deterministic
    o1 = sc1.render(stride=stride, nsamples=1)
    apply_frame_state(rt, st)
    _lib.check(lib.mrtx_set_uint(ctx, b"debug_hits", 1, 0))
    _lib.check(lib.mrtx_set_uint(ctx, b"jitter", 0, 0))
    _lib.check(lib.mrtx_render(ctx, 0, 0, args.img_w, args.img_h, 0, 1, 1))
    g = rt.get_hit_records_f64()[::stride, ::stride]
    _lib.check(lib.mrtx_set_uint(ctx, b"debug_hits", 0, 0))
    oh, gh = o1["hit64"][..., 0] > 0, g[..., 0] > 0
    both = oh & gh
    texel = 2.0 * np.pi * scene.MOON_RADIUS / args.map_w
    dr = np.abs(g[..., 1] - o1["hit64"][..., 1])[both] / texel
    # (b) the frame as benchmarked
    scj = OracleScene(ldem_host, jitter=args.spp > 1, **kw)
    t0 = time.perf_counter()
    oj = scj.render(stride=stride, nsamples=args.spp)
    dt = time.perf_counter() - t0
    _lib.check(lib.mrtx_set_uint(ctx, b"jitter", 1 if args.spp > 1 else 0, 0))
    _lib.check(lib.mrtx_render(ctx, 0, 0, args.img_w, args.img_h, 0, args.spp, 1))
    _lib.check(lib.mrtx_resolve(ctx))
    img = np.empty((args.img_h, args.img_w, 4), np.uint8)
    _lib.check(lib.mrtx_read_rgba8(ctx, img.ctypes.data))
    gi = img[::stride, ::stride, :3].astype(np.float64)
    oi = scj.tonemap(oj["accum"])[..., :3].astype(np.float64)
    mae = float(np.abs(gi - oi).mean())
    mse = float(((gi - oi) ** 2).mean())
    psnr = 99.0 if mse == 0 else float(10.0 * np.log10(255.0 ** 2 / mse))
    npx = oi.shape[0] * oi.shape[1]
    parity = {"frame": frame, "sub_grid": f"every {stride}th pixel in x and y ({npx} pixels)",
              "hit_mask_mismatches": int((oh != gh).sum()), "hits_compared": int(both.sum()),
              "hit_radius_max_err_texel": float(dr.max()) if dr.size else 0.0, "hit_radius_over_1e-3_texel": int((dr > 1e-3).sum()),
              "image_mae_8bit": round(mae, 4), "image_psnr_db": round(psnr, 2), "spp": args.spp,
              "tolerance": "hit radius <= 1e-3 texel, 8-bit MAE <= 1, PSNR >= 40 dB (north_star)"}
    parity["ok"] = bool(parity["hit_mask_mismatches"] == 0 and parity["hit_radius_over_1e-3_texel"] == 0 and mae <= 1.0 and psnr >= 40.0)
    inside = int((oj["stats"][..., 0] > 0).sum()) * args.spp
    shadow = int((oj["stats"][..., 1] > 0).sum()) * args.spp          # last-sample estimate
    cpu = {"value": round((inside + shadow) / dt / 1e6, 4), "unit": "Mrays/s", "cores": cores, "kind": "port", "seconds": round(dt, 3),
           "sample": f"every {stride}th pixel in x and y of frame {frame} ({npx} pixels x {args.spp} spp), {dt:.1f} s, float64 oracle "
                     f"(exhaustive cell walk, no pyramid), OpenMP {cores} threads; rays counted as for `value`"}
    return parity, cpu


def run_frames(args):
    rank, world, local = dist_env()
    D = Dist(rank, world, local)
    from moonrtx_b200 import _lib
    from moonrtx_b200.video import apply_frame_state, render_timelapse_delivered
    if FRAMES_PER_STEP % world:
        raise SystemExit(f"--gpus must divide {FRAMES_PER_STEP}")
    t_setup = time.time()
    rt, ldem, radius_scale = build_scene(args, local)
    dev = rt._dev
    lib, ctx = dev.lib, dev.ctx
    D.comm_init(rt)
    if args.delivery == "p2p" and not D.p2p_connect(rt):
        args.delivery = "nccl"
    n_warm, n_timed = args.warmup * FRAMES_PER_STEP, args.steps * FRAMES_PER_STEP
    frames = list(range(n_warm + n_timed))
    mine_warm = [f for f in frames[:n_warm] if f % world == rank]
    mine = [f for f in frames[n_warm:] if f % world == rank]
    states = {f: sweep_state(f) for f in frames}
    t_setup = time.time() - t_setup
    _lib.check(lib.mrtx_set_uint(ctx, b"jitter", 1 if args.spp > 1 else 0, 0))

    def device_frame(f):
        apply_frame_state(rt, states[f])
        _lib.check(lib.mrtx_render(ctx, 0, 0, args.img_w, args.img_h, 0, args.spp, 1))
        _lib.check(lib.mrtx_resolve(ctx))

    # ---- device-resident throughput (value) + the kernels' own launch durations (roofline) -------------------------
    clocks = ClockSampler(enabled=rank == 0, gpus=range(world))
    clocks.wait_ready()
    for f in mine_warm:
        device_frame(f)
    rt.counters(reset=True); rt.defer_stats(reset=True)
    _lib.check(lib.mrtx_set_uint(ctx, b"profile", 1, 0))
    rt.kernel_times(reset=True)
    D.barrier(dev)
    clocks.begin()
    dev.timer_start()
    for f in mine:
        device_frame(f)
    ms_total = dev.timer_stop()
    D.barrier(dev)
    clocks.end()
    clock_info = clocks.stop()
    kt = rt.kernel_times(reset=True)
    _lib.check(lib.mrtx_set_uint(ctx, b"profile", 0, 0))
    c = rt.counters()
    defer = rt.defer_stats()
    ms_total = D.max(ms_total)
    rays_all, rays_all_culled = D.sum(float(walked(c))), D.sum(float(c["primary_rays"] + c["shadow_rays"]))
    value = rays_all / (ms_total * 1e-3) / 1e6

    depth = quadtree_depth(args.map_w)
    b_floor = 32 * (depth + 3)
    peak, peak_src = measured_peak()
    nl = max(1, kt["launches"])
    fast_ms, shadow_ms, ref_ms = kt["trace_kernel_fast"] / nl, kt["shadow_kernel"] / nl, kt["trace_kernel_referee"] / nl
    shade_ms = kt["shade_kernel"] / nl
    path_ms = sum(kt[k] for k in ("cull_kernel", "beam_kernel", "trace_kernel_fast", "shade_kernel", "shadow_kernel", "trace_kernel_referee", "fold_kernel")) / nl
    nf = max(1, len(mine))
    prim_bytes = b_floor * c["primary_in_sphere"] / nf
    shad_bytes = b_floor * c["shadow_rays"] / nf
    traffic = ncu_traffic() if (args.map_w, args.img_w, args.spp) == (MAP_W, IMG_W, 16) else None

    def gbs(nbytes, ms):
        return nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": "trace_kernel_pool<int16> (primary ray to its first hit, which goes to the hit queue; undecided rays parked in per-warp pools)",
        "achieved": round(gbs(prim_bytes, fast_ms), 2), "peak": peak, "unit": "GB/s", "frac": round(gbs(prim_bytes, fast_ms) / peak, 5),
        "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
        "algorithmic_bytes_per_launch": int(prim_bytes), "bytes_per_ray": b_floor, "rays_per_launch": int(c["primary_in_sphere"] / nf),
        "kernel_ms": round(fast_ms, 3), "peak_source": peak_src, "launches_timed": kt["launches"],
        "timing": "CUDA events on the launching stream at the kernel boundaries of every timed frame (mrtx_kernel_times)",
        "shadow_kernel": {"kernel_ms": round(shadow_ms, 3), "achieved": round(gbs(shad_bytes, shadow_ms), 2), "frac": round(gbs(shad_bytes, shadow_ms) / peak, 5),
                          "rays_per_launch": int(c["shadow_rays"] / nf), "algorithmic_bytes_per_launch": int(shad_bytes)},
        "whole_path": {"kernels": "cull + trace_kernel_pool + shade_kernel + shadow_kernel + trace_kernel_referee + fold (one mrtx_render)",
                       "ms": round(path_ms, 3), "referee_ms": round(ref_ms, 3), "shade_ms": round(shade_ms, 3), "achieved": round(gbs(prim_bytes + shad_bytes, path_ms), 2),
                       "frac": round(gbs(prim_bytes + shad_bytes, path_ms) / peak, 5)},
        "counted_bytes_per_ray": round(32.0 * (c["node_visits"] + 2 * c["patch_tests"]) / max(1.0, walked(c)), 1),
        "kernel_share_of_step": round(fast_ms * nf / ms_total, 4) if world == 1 else None,
    }

    # ---- end to end through the public API (e2e): the export loop with ordered delivery to rank 0 -------------------------
    e2e = None
    if not args.skip_e2e:
        ov_cache = {}

        def overlay_for(f):
            if f not in ov_cache:
                ov_cache.clear()
                ov_cache[f] = overlay_image(args.img_h, args.img_w, 37 * f)
            return ov_cache[f]
        sums = []
        warm_states = [states[f] for f in frames[:n_warm]]
        render_timelapse_delivered(rt, warm_states, rank, world, 0, on_frame=None, overlay_for=overlay_for)
        D.barrier(dev)
        rt.counters(reset=True)
        t0 = time.perf_counter()
        timed_states = [states[f] for f in frames[n_warm:]]
        render_timelapse_delivered(rt, timed_states, rank, world, 0, on_frame=lambda i, img: sums.append(int(img[::97, ::89, :3].sum())),
                                   overlay_for=lambda i: overlay_for(n_warm + i))
        D.barrier(dev)
        dt = D.max(time.perf_counter() - t0)
        ce = rt.counters()
        e2e_rays = D.sum(float(walked(ce)))
        frame_bytes = args.img_w * args.img_h * 4
        e2e = {"value": round(e2e_rays / dt / 1e6, 2), "unit": "Mrays/s",
               "h2d_bytes_per_step": int(FRAMES_PER_STEP * (frame_bytes + 1024)), "d2h_bytes_per_step": int(FRAMES_PER_STEP * frame_bytes),
               "nvlink_bytes_per_step": int(FRAMES_PER_STEP * (world - 1) // world * frame_bytes),
               "ms_per_step": round(dt * 1e3 / args.steps, 2), "frames_per_s": round(n_timed / dt, 3),
               "frames_delivered_in_order_to_rank0": len(sums) if rank == 0 else None,
               "frame_checksum": (sum(sums) & 0xffffffff) if rank == 0 else None,
               "delivery": None if world == 1 else ("peer-memory mailboxes: copy-engine copies over NVLink + stream memory operations (mrtx_p2p_*)"
                                                    if args.delivery == "p2p" else "ncclSend / ncclRecv"),
               "api": "video.render_timelapse_delivered: B200OptiX.submit_frame(dst=0) / recv_frame / wait_frame, two frames in flight per rank"}

    # ---- the reference's own path_seg_range (2, 4): two interreflection bounces, device-resident, N = 1 -----------------
    bounce = None
    if rank == 0 and world == 1 and not args.skip_e2e:
        rt.set_uint("path_seg_range", 2, 4)
        device_frame(frames[n_warm])
        rt.counters(reset=True)
        dev.synchronize()
        dev.timer_start()
        for f in frames[n_warm:n_warm + 2]:
            device_frame(f)
        bms = dev.timer_stop() / 2
        cb = rt.counters()
        bounce = {"path_seg_range": [2, 4], "ms_per_frame": round(bms, 3),
                  "rays_walked_per_frame": int(walked(cb) // 2), "node_visits_per_frame": int(cb["node_visits"] // 2),
                  "note": "camera ray + 2 diffuse bounces, direct light with a shadow ray at each of the 3 hits (SURVEY.md 8f N2); "
                          "bounce rays are not in the ray counters, their node visits are"}
        rt.set_uint("path_seg_range", 2, 2)

    # ---- oracle parity on the benchmarked frame + CPU baseline (rank 0, N = 1 only) -----------------------------------
    parity, cpu = None, None
    if rank == 0 and world == 1 and not args.skip_cpu:
        parity, cpu = parity_block(args, rt, ldem.download((args.map_h, args.map_w), np.int16), radius_scale, n_warm, budget_s=20.0)

    # ---- config 1 beside it (rank 0): the data_loader downscale against the HBM peak -----------------------------------
    downscale = None
    if rank == 0 and not args.skip_downscale:
        downscale = downscale_lines(dev, peak, args.skip_cpu)
        if (args.map_w, args.map_h) == (MAP_W, MAP_H) and world == 1:
            downscale["full_resolution_92160x46080"] = downscale_fullres(dev, ldem, peak)

    if rank == 0:
        line = {
            "metric": "Mrays/s (primary+shadow) @4K", "value": round(value, 2), "unit": "Mrays/s",
            "value_incl_culled": round(rays_all_culled / (ms_total * 1e-3) / 1e6, 2),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (pyramid walk and patch test in a cell-local frame re-based in f64) over int16 texels; f64 referee for undecided samples", "data": "synthetic",
            "config": {"workload": f"{FRAMES_PER_STEP} frames per step of the F11 terminator sweep (BASELINE config 4; each frame = config 3: "
                                   f"{args.img_w}x{args.img_h}, {args.map_w}x{args.map_h} int16 synthetic LOLA map + "
                                   f"{args.color_w // COLOR_K}x{args.color_h // COLOR_K} colour texture, {args.spp} spp); frame f on rank f mod N",
                       "spp": args.spp, "camera": "default whole-disk, fov 4.2422 deg", "sun": "terminator sweep from phase 90 deg, 10 min per frame",
                       "light": "direct (path_seg_range 2, 2: camera ray + sun shadow ray, the north_star path)",
                       "frames_per_step": FRAMES_PER_STEP, "frames_timed": n_timed,
                       "l2_hygiene": "inputs_larger_than_L2 (8.5 GB map + 2.9 GB pyramid)", "setup_s": round(t_setup, 1)},
            "rays": {"walked_per_frame": int(rays_all / n_timed), "primary_in_sphere_per_frame": int(D.sum(float(c["primary_in_sphere"])) / n_timed) if world == 1 else None,
                     "primary_incl_culled_per_frame": c["primary_rays"] // nf, "shadow_per_frame": c["shadow_rays"] // nf,
                     "node_visits_per_frame": c["node_visits"] // nf, "patch_tests_per_frame": c["patch_tests"] // nf, "overflow": c["overflow"],
                     "samples_deferred_to_f64_referee_per_frame": defer["deferred_samples"] // nf,
                     "defer_reasons_primary": defer["primary_reasons"], "defer_reasons_shadow": defer["shadow_reasons"]},
            "frames_per_s": round(n_timed / (ms_total * 1e-3), 3), "ms_per_frame_per_gpu": round(ms_total / nf, 3),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "interreflection": bounce,
            "gpu_launches": launches_per_frame(args) * n_timed, "clocks": clock_info, "downscale": downscale,
        }
        emit(line)
    rt.close()
    D.close()
    if parity is not None and not parity["ok"]:
        sys.stderr.write(f"PARITY OUT OF TOLERANCE: {parity}\n")
        sys.exit(3)


# ------------------------------------------------------------------------------------------------
def run_single_frame_mode(args):
    """--mode samples: ONE config-3 frame, its 16 samples split across the ranks, float4 accumulators summed with
    ncclAllReduce (mrtx_allreduce_accum).  --mode tiles: ONE config-5 frame (7680x4320, fov 2.5 deg, the terminator at
    the disk centre: grazing incidence everywhere), interleaved 64x64 tiles split across the ranks in one launch, owned
    tiles tone-mapped into the send buffer and exchanged with ncclAllGather (mrtx_allgather_tiles).  Strong scaling of a
    single frame; every rank checks its gathered frame against the frame it renders alone."""
    rank, world, local = dist_env()
    D = Dist(rank, world, local)
    from moonrtx_b200 import _lib
    from moonrtx_b200.video import apply_frame_state
    tiles = args.mode == "tiles"
    if tiles:
        args.img_w, args.img_h = 7680, 4320
    rt, ldem, radius_scale = build_scene(args, local)
    dev = rt._dev
    lib, ctx = dev.lib, dev.ctx
    D.comm_init(rt)
    if world == 1:
        from moonrtx_b200.optix import B200OptiX
        rt.comm_init(0, 1, B200OptiX.comm_unique_id())
    st = sweep_state(0)
    apply_frame_state(rt, st)
    if tiles:
        cam = rt.get_camera("cam1")
        rt.update_camera("cam1", eye=cam["Eye"], target=cam["Target"], up=cam["Up"], fov=2.5)
    _lib.check(lib.mrtx_set_uint(ctx, b"jitter", 1 if args.spp > 1 else 0, 0))
    n = args.spp
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world

    def step(timed=None):
        if tiles:
            _lib.check(lib.mrtx_render_tiles(ctx, 64, 0, n, 1))
        else:
            _lib.check(lib.mrtx_render(ctx, 0, 0, args.img_w, args.img_h, lo, hi - lo, 1))
        if timed is not None:
            dev.synchronize()
            dev.timer_start()
        if tiles:
            _lib.check(lib.mrtx_allgather_tiles(ctx, 64))
        else:
            _lib.check(lib.mrtx_allreduce_accum(ctx))
            _lib.check(lib.mrtx_resolve(ctx))
        if timed is not None:
            timed.append(dev.timer_stop())

    # the frame this rank renders alone (outside the timed region): what the sharded frame must reproduce
    _lib.check(lib.mrtx_render(ctx, 0, 0, args.img_w, args.img_h, 0, n, 1))
    _lib.check(lib.mrtx_resolve(ctx))
    solo = np.empty((args.img_h, args.img_w, 4), np.uint8)
    _lib.check(lib.mrtx_read_rgba8(ctx, solo.ctypes.data))
    clocks = ClockSampler(enabled=rank == 0, gpus=range(world))
    clocks.wait_ready()
    for _ in range(args.warmup):
        step()
    rt.counters(reset=True)
    D.barrier(dev)
    clocks.begin()
    dev.timer_start()
    for _ in range(args.steps):
        step()
    ms_total = dev.timer_stop()
    D.barrier(dev)
    clocks.end()
    clock_info = clocks.stop()
    c = rt.counters()
    ms_total = D.max(ms_total)
    rays_all = D.sum(float(walked(c)))
    coll = []
    for _ in range(3):
        step(coll)
    got = np.empty_like(solo)
    _lib.check(lib.mrtx_read_rgba8(ctx, got.ctypes.data))
    diff = np.abs(got[..., :3].astype(np.int32) - solo[..., :3].astype(np.int32))
    ok = bool(diff.max() <= (0 if tiles else 1))
    all_ok = D.sum(1.0 if ok else 0.0) == world
    coll_ms = D.max(statistics.median(coll))
    if rank == 0:
        frame_bytes = args.img_w * args.img_h * 4
        emit({
            "metric": "Mrays/s (primary+shadow) @8K, one frame screen-tiled" if tiles else "Mrays/s (primary+shadow) @4K, one frame sample-split",
            "mode": args.mode, "value": round(rays_all / (ms_total * 1e-3) / 1e6, 2), "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 walk / f64 re-basing over int16 texels", "data": "synthetic",
            "config": {"workload": (f"BASELINE config 5: {args.img_w}x{args.img_h}, fov 2.5 deg on the terminator at the disk centre, {n} spp, "
                                    f"interleaved 64x64 tiles over {world} ranks + ncclAllGather of RGBA8 tiles") if tiles else
                                   (f"BASELINE config 3: {args.img_w}x{args.img_h}, {n} spp split over {world} ranks + ncclAllReduce of the float4 accumulators"),
                       "map": f"{args.map_w}x{args.map_h} int16 synthetic LOLA map", "spp": n},
            "collective": {"what": "resolve of owned tiles into the send buffer + ncclAllGather + untile" if tiles else "ncclAllReduce(sum, f32) of the accumulators + resolve",
                           "ms": round(coll_ms, 3), "bytes": int(frame_bytes if tiles else frame_bytes * 4),
                           "timing": "CUDA events around the collective step, median of 3, max over ranks"},
            "matches_single_gpu_frame": all_ok, "max_abs_diff_8bit": int(diff.max()),
            "frames_per_s": round(args.steps / (ms_total * 1e-3), 3), "clocks": clock_info,
            "gpu_launches": (launches_per_frame(args) + (1 if tiles else 0)) * args.steps,
        })
    rt.close()
    D.close()
    if not all_ok:
        sys.exit(3)


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the CPU arm.  The reference's own engine (PlotOptiX) is a closed binary that cannot be installed
    offline, so the float64 oracle port is what runs (kind 'port'), on every host thread (torchrun exports
    OMP_NUM_THREADS=1 to its workers: set back explicitly), on a bounded sub-grid of the same frames.  Its inputs do not come
    from the code under test except the synthetic map itself: radius_scale is numpy's."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from moonrtx_b200 import scene
    try:
        from moonrtx_b200 import _lib
        from moonrtx_b200.device import Device
        dev = Device(local)
        buf = dev.alloc(args.map_w * args.map_h * 2)
        _lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, buf.ptr, args.map_w, args.map_h, SEED))
        ldem_host = buf.download((args.map_h, args.map_w), np.int16)
        buf.free(); dev.close()
    except Exception as e:
        emit({"impl": "reference", "unavailable": f"could not generate the synthetic map: {e}"})
        return
    radius_scale = radius_scale_numpy(ldem_host)
    from oracle.render_oracle import OracleScene, lib as orc_lib
    try:
        orc_lib().omp_set_num_threads(cores)             # libgomp, loaded with the oracle
    except Exception:
        pass
    import ctypes as C
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except Exception:
        pass
    per_step_budget = max(2.0, min(20.0, 120.0 / (args.steps + args.warmup)))
    vals, secs, last = [], [], None
    for j in range(args.warmup + args.steps):
        st = sweep_state(j * FRAMES_PER_STEP)
        sc = OracleScene(ldem_host, scale=SCALE_F32, radius_scale=radius_scale, img_w=args.img_w, img_h=args.img_h, u=st.u, v=st.v,
                         eye=st.eye, target=st.target, up=st.up, fov=st.fov, light_pos=st.light_pos, light_radius=st.light_radius,
                         light_radiance=scene.light_radiance(80.0), jitter=args.spp > 1)
        if last is None:
            t0 = time.perf_counter()
            o = sc.render(stride=96, nsamples=1)
            per_px = (time.perf_counter() - t0) / (o["accum"].shape[0] * o["accum"].shape[1]) * args.spp
            stride = max(1, int(np.ceil(np.sqrt(args.img_w * args.img_h / max(1, int(per_step_budget / max(per_px, 1e-9)))))))
        t0 = time.perf_counter()
        o = sc.render(stride=stride, nsamples=args.spp)
        dt = time.perf_counter() - t0
        npx = o["accum"].shape[0] * o["accum"].shape[1]
        rays = (int((o["stats"][..., 0] > 0).sum()) + int((o["stats"][..., 1] > 0).sum())) * args.spp
        last = (rays / dt / 1e6, dt, npx)
        if j >= args.warmup:
            vals.append(last[0]); secs.append(dt)
    v = statistics.mean(vals)
    cpu = {"value": round(v, 4), "unit": "Mrays/s", "cores": cores, "kind": "port", "omp_threads": cores,
           "sample": f"every {stride}th pixel in x and y of the first frame of each step ({last[2]} pixels x {args.spp} spp per step), "
                     f"float64 oracle (exhaustive cell walk, no pyramid), OpenMP {cores} threads; rays counted as for the CUDA arm's `value`"}
    emit({
        "impl": "reference", "metric": "Mrays/s (primary+shadow) @4K", "value": round(v, 4), "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * statistics.mean(secs), 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{FRAMES_PER_STEP} frames per step of the F11 terminator sweep (BASELINE config 4; each frame = config 3: "
                               f"{args.img_w}x{args.img_h}, {args.map_w}x{args.map_h} int16 synthetic LOLA map, {args.spp} spp); "
                               f"bounded sample per step: a sub-grid of the step's first frame"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(v, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="frames", choices=["frames", "samples", "tiles"])
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--map-w", type=int, default=MAP_W)
    ap.add_argument("--map-h", type=int, default=MAP_H)
    ap.add_argument("--img-w", type=int, default=IMG_W)
    ap.add_argument("--img-h", type=int, default=IMG_H)
    ap.add_argument("--color-w", type=int, default=COLOR_W)
    ap.add_argument("--color-h", type=int, default=COLOR_H)
    ap.add_argument("--delivery", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1, frames mode: frames reach rank 0 through peer-memory mailboxes (copy engines) or ncclSend / ncclRecv")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-downscale", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "frames":
        run_frames(args)
    else:
        run_single_frame_mode(args)


if __name__ == "__main__":
    main()
