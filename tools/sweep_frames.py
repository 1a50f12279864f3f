"""Development helper: kernel time of single frames along the terminator sweep (BASELINE config 4 frame indices)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_trace as bt
from moonrtx_b200 import _lib, scene
from moonrtx_b200.synth import synth_ephemeris
from moonrtx_b200.video import apply_frame_state
rt, info = bt.setup(92160, 46080, 3840, 2160, ds=1)
lib, ctx = rt._dev.lib, rt._dev.ctx
_lib.check(lib.mrtx_set_uint(ctx, b"kernel", int(os.environ.get("KERNEL", "2")), 0))
for name in ("long_walk", "referee_budget"):
    if os.environ.get(name.upper()):
        _lib.check(lib.mrtx_set_uint(ctx, name.encode(), int(os.environ[name.upper()]), 0))
for f in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,8,24,48,87,160,239").split(",")]:
    apply_frame_state(rt, scene.frame_state(synth_ephemeris(f * 10.0)))
    r = bt.time_frame(rt, 16, reps=2)
    print(json.dumps({"frame": f, "ms": r["ms"], "kernel_ms": r["kernel_ms"], "defer_shadow": r["defer"]["shadow_reasons"], "nodes": r["node_visits"], "tests": r["patch_tests"], "shadow": r["shadow_rays"], "occluded": r["shadow_occluded"],
                      "in_sphere": r["primary_in_sphere"], "deferred": r["defer"]["deferred_samples"], "nodes_per_inray": r["nodes_per_inray"]}), flush=True)
