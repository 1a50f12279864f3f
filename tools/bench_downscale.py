"""Kernel-only timing of the elevation downscale (config 1) - development helper."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moonrtx_b200 import _lib
from moonrtx_b200.device import get_device
from moonrtx_b200.data_loader import downscale_elevation_dev

def main():
    W, H = 23040, 11520
    dev = get_device()
    src = dev.alloc(W * H * 2)
    _lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, src.ptr, W, H, 20240314))
    for ds in (4, 3, 2, 16, 1, 8):
        out = dev.alloc((W // ds) * (H // ds) * 4)
        for _ in range(3):
            downscale_elevation_dev(src, W, H, ds, out, want_scale=False)
        ts = []
        for _ in range(10):
            dev.l2_flush(); dev.synchronize()
            dev.timer_start()
            downscale_elevation_dev(src, W, H, ds, out, want_scale=False)
            ts.append(dev.timer_stop())
        b = 2 * W * H + 4 * (W // ds) * (H // ds)
        t = float(np.median(ts))
        print(json.dumps({"ds": ds, "ms": round(t, 4), "min_ms": round(min(ts), 4), "GBps": round(b / t / 1e6, 1), "bytes": b}))
        out.free()

if __name__ == "__main__":
    main()
