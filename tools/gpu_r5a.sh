#!/bin/bash
# round-2 GPU session A: tests, then beam / ceiling variants on config 3 and config 5
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r5a_tests.log
V="beam=0,0+ceiling=0;beam=1,1+ceiling=0;beam=1,2+ceiling=0;beam=1,3+ceiling=0;beam=1,2+ceiling=3;beam=1,2+ceiling=4;beam=1,2+ceiling=2;beam=0,0+ceiling=3"
( VARIANTS="$V" timeout 600 python tools/bench_trace.py cfg3 16 2>&1 ) > gpurun_out/r5a_cfg3.log
( VARIANTS="beam=0,0+ceiling=0;beam=1,2+ceiling=3" timeout 600 python tools/bench_trace.py cfg5 16 2>&1 ) > gpurun_out/r5a_cfg5.log
python - <<'PY'
import json
for f in ("gpurun_out/r5a_cfg3.log", "gpurun_out/r5a_cfg5.log"):
    print("==", f)
    for line in open(f):
        if not line.startswith("{"):
            if "rror" in line: print(line.strip()[:300])
            continue
        d = json.loads(line)
        if "compare" in d: print("   cmp", d["compare"], "mae", round(d["img_mae"], 5), "max", d["img_max"], "px", d["pixels_differ"], "gt2", d["pixels_differ_gt2"], "accmax", d["accum_max_abs"], d["accum_w_equal"])
        elif "ms" in d: print(d["kernel"], "ms", d["ms"], "nodes/ray", d["nodes_per_inray"], "nodes", d["node_visits"], "tests", d.get("patch_tests"), "defer", d["defer"]["deferred_samples"])
PY
cat gpurun_out/r5a_tests.log
