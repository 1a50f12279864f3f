"""Development helper: with a -DMRTX_REFEREE_TIMING build (MRTX_LIB=...), why the referee ran float64 tests, per frame."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_trace as bt
from moonrtx_b200 import scene
from moonrtx_b200.synth import synth_ephemeris
from moonrtx_b200.video import apply_frame_state
rt, info = bt.setup(92160, 46080, 3840, 2160, ds=1)
names = ["test_phases", "test_phase_lanes", "trav_steps", "trav_step_lanes", "start_phases", "start_phase_lanes", "refills"]
reasons = [1, 2, 3, 7, 8, 9, 12]
for f in [int(v) for v in sys.argv[1].split(",")]:
    apply_frame_state(rt, scene.frame_state(synth_ephemeris(f * 10.0)))
    r = bt.time_frame(rt, 16, reps=1)
    print(f, r["ms"], {f"r{k}": r[n] // 2 for k, n in zip(reasons, names)}, flush=True)
