"""Development helper: time the trace kernel on BASELINE configs (kernel-only, CUDA events)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moonrtx_b200 import _lib
if os.environ.get("MRTX_LIB"):
    _lib.LIB_PATH = os.environ["MRTX_LIB"]          # tuning variants (tools/build_variant.py)
from moonrtx_b200.device import Device
from moonrtx_b200.optix import B200OptiX
from moonrtx_b200.data_loader import downscale_elevation_dev
from moonrtx_b200 import scene
from moonrtx_b200.synth import synth_ephemeris

def setup(W, H, iw, ih, ds=1, tex=True):
    rt = B200OptiX(width=iw, height=ih)
    dev = rt._dev
    t = time.time()
    src = dev.alloc(W * H * 2)
    _lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, src.ptr, W, H, 20240314))
    dev.synchronize(); t_synth = time.time() - t
    rt.set_data("moon", geom="ParticleSetTextured", geom_attr="DisplacedSurface", pos=[0, 0, 0], u=[0, 0, 1], v=[0, -1, 0], r=10.0)
    t = time.time()
    if ds == 1:
        # radius_scale from the max count, as data_loader.py:232-242 would compute it
        if os.environ.get("BANDS"):
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            import bench
            rs = bench._radius_scale_ds1(dev, src, W, H)
        else:
            out, rs = downscale_elevation_dev(src, W, H, 1, want_scale=True)
            out.free()
        print("radius_scale", rs)
        rt.set_displacement_i16("moon", (src, W, H), radius_scale=rs)
    else:
        out, rs = downscale_elevation_dev(src, W, H, ds)
        _lib.check(dev.lib.mrtx_set_displacement_f32_dev(dev.ctx, out.ptr, W // ds, H // ds, 0))
        rt._keep = out
    dev.synchronize(); t_pyr = time.time() - t
    if tex:
        cw, ch = 6840, 3420
        cb = dev.alloc(cw * ch * 3)
        _lib.check(dev.lib.mrtx_synth_color_bgr_dev(dev.ctx, cb.ptr, cw, ch, 4720))
        bgr = cb.download((ch, cw, 3), np.uint8)
        from moonrtx_b200.data_loader import color_texture
        rt.set_texture_2d("moon_color", color_texture(bgr, 2.2, 1))
    st = scene.frame_state(synth_ephemeris(0.0))
    from moonrtx_b200.video import apply_frame_state
    rt.setup_camera("cam1", eye=st.eye, target=st.target, up=st.up, fov=st.fov)
    rt.setup_light("sun", color=scene.light_radiance(80), radius=st.light_radius)
    apply_frame_state(rt, st)
    return rt, {"synth_s": round(t_synth, 2), "setup_s": round(t_pyr, 2)}

def time_frame(rt, spp, reps=3):
    dev = rt._dev
    lib, ctx = dev.lib, dev.ctx
    _lib.check(lib.mrtx_set_uint(ctx, b"jitter", 1 if spp > 1 else 0, 0))
    ts = []
    _lib.check(lib.mrtx_set_uint(ctx, b"profile", 1, 0))
    for i in range(reps + 1):
        rt.kernel_times(reset=True)
        rt.counters(reset=True)
        rt.defer_stats(reset=True)
        dev.synchronize()
        dev.timer_start()
        _lib.check(lib.mrtx_render(ctx, 0, 0, rt._width, rt._height, 0, spp, 1))
        ms = dev.timer_stop()
        if i:
            ts.append(ms)
    c = rt.counters()
    kt = rt.kernel_times()
    _lib.check(lib.mrtx_set_uint(ctx, b"profile", 0, 0))
    rays = c["primary_rays"] + c["shadow_rays"]
    ms = float(np.median(ts))
    n = max(1, kt.pop("launches"))
    return {"spp": spp, "ms": round(ms, 2), "kernel_ms": {k.replace("_kernel", "").replace("trace_", ""): round(v / n, 2) for k, v in kt.items()}, "defer": rt.defer_stats(), "Mrays_s": round(rays / ms / 1e3, 1), **c,
            "nodes_per_inray": round(c["node_visits"] / max(1, c["primary_in_sphere"] + c["shadow_rays"]), 1)}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    if which == "cfg2":
        rt, info = setup(23040, 11520, 1920, 1080, ds=4)
    elif which == "cfg3":
        rt, info = setup(92160, 46080, 3840, 2160, ds=1)
    elif which == "cfg5":
        # BASELINE config 5: 8K narrow-FOV view of the terminator at the disk centre (fov 2.5 deg)
        rt, info = setup(92160, 46080, 7680, 4320, ds=1)
        cam = rt.get_camera("cam1")
        rt.update_camera("cam1", eye=cam["Eye"], target=cam["Target"], up=cam["Up"], fov=2.5)
    elif which == "cfg1k":
        rt, info = setup(23040, 11520, 3840, 2160, ds=1)
    print(json.dumps(info))
    spps = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else ((1, 4) if which != "cfg3" else (1, 16))
    kernels = [int(v) for v in os.environ.get("KERNELS", "2").split(",")]
    lib, ctx = rt._dev.lib, rt._dev.ctx
    imgs = {}
    if os.environ.get("START"):
        a, b = [int(v) for v in os.environ["START"].split(",")]
        _lib.check(lib.mrtx_set_uint(ctx, b"start_levels", a, b))
    # VARIANTS="beam=0,0;beam=1,2;start_levels=3,1": engine switches (mrtx_set_uint name=a,b) tried one after the other on kernel 2,
    # every frame compared with the first variant's
    variants = [v for v in os.environ.get("VARIANTS", "").split(";") if v]
    if variants:
        kernels = list(range(100, 100 + len(variants)))
    for spp in spps:
        for k in kernels:
            if k >= 100:
                _lib.check(lib.mrtx_set_uint(ctx, b"kernel", 2, 0))
                for item in variants[k - 100].split("+"):
                    name, val = item.split("=")
                    a, b = (val.split(",") + ["0"])[:2]
                    _lib.check(lib.mrtx_set_uint(ctx, name.encode(), int(a), int(b)))
            else:
                _lib.check(lib.mrtx_set_uint(ctx, b"kernel", k, 0))
            r = time_frame(rt, spp)
            r["kernel"] = k if k < 100 else variants[k - 100]
            print(json.dumps(r), flush=True)
            _lib.check(lib.mrtx_resolve(ctx))
            img = np.empty((rt._height, rt._width, 4), np.uint8)
            _lib.check(lib.mrtx_read_rgba8(ctx, img.ctypes.data))
            acc = np.empty((rt._height, rt._width, 4), np.float32)
            _lib.check(lib.mrtx_read_accum_f32(ctx, acc.ctypes.data))
            imgs[k] = (img, acc)
        for other in kernels[1:]:
            a, b = imgs[kernels[0]], imgs[other]
            d = np.abs(a[0][..., :3].astype(np.int32) - b[0][..., :3].astype(np.int32))
            da = np.abs(a[1][..., :3] - b[1][..., :3])
            print(json.dumps({"spp": spp, "compare": [kernels[0], other], "img_mae": float(d.mean()), "img_max": int(d.max()),
                              "pixels_differ": int((d.max(axis=2) > 0).sum()), "pixels_differ_gt2": int((d.max(axis=2) > 2).sum()),
                              "accum_max_abs": float(da.max()), "accum_w_equal": bool(np.array_equal(a[1][..., 3], b[1][..., 3]))}), flush=True)
