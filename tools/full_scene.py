"""Development helper: one config-3 frame with everything the reference draws switched on - star map, Sun disk, grid tubes,
path_seg_range (2, 4) - timed feature by feature (kernel-only, CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bench_trace as bt
from test_tubes_host import grid_graph
rt, info = bt.setup(92160, 46080, 3840, 2160, ds=1)
def t(label):
    r = bt.time_frame(rt, 16, reps=2)
    print(json.dumps({"what": label, "ms": r["ms"], "kernel_ms": r["kernel_ms"], "nodes": r["node_visits"], "shadow_rays": r["shadow_rays"]}), flush=True)
t("plain (direct light, black sky)")
rng = np.random.default_rng(1)
rt.set_background_mode("TextureEnvironment")
rt.set_background((rng.random((1024, 2048, 3)) ** 8).astype(np.float32), gamma=2.2, rt_format="UByte4")
rt.set_data("sun_disk", geom="ParticleSet", mat="flat", pos=[[2000.0, -21000.0, 300.0]], r=100.0, c=1.0)
t("+ star map and Sun disk")
pos, edges = grid_graph()
rt.set_graph("moon_grid_lines", pos=pos, edges=edges, r=0.006, c=[0.5, 0.5, 0.5])
t("+ grid tubes (3 267 segments)")
rt.set_uint("path_seg_range", 2, 4)
t("+ path_seg_range (2, 4)")
rt.close()
