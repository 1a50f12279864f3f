"""Development helper: write the GPU-synthesised LDEM (int16) of a given size to gpurun_out/ for host-side debugging."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moonrtx_b200 import _lib
from moonrtx_b200.device import get_device
W, H = int(sys.argv[1]), int(sys.argv[2])
dev = get_device()
src = dev.alloc(W * H * 2)
_lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, src.ptr, W, H, 20240314))
a = src.download((H, W), np.int16)
np.save(f"gpurun_out/synth_{W}x{H}.npy", a)
print(a.min(), a.max(), a.std())
