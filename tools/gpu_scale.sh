#!/bin/bash
# Development helper: bench.py under torchrun at N ranks (as the driver launches it), default mode + the two sharding modes
N=${1:-2}; tag=${2:-r6c}; steps=${3:-3}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps $steps --warmup 1 > gpurun_out/${tag}_frames_n$N.json 2> gpurun_out/${tag}_frames_n$N.err; echo "frames rc=$?"
run --mode samples --steps 10 --warmup 3 > gpurun_out/${tag}_samples_n$N.json 2> gpurun_out/${tag}_samples_n$N.err; echo "samples rc=$?"
run --mode tiles --steps 5 --warmup 2 > gpurun_out/${tag}_tiles_n$N.json 2> gpurun_out/${tag}_tiles_n$N.err; echo "tiles rc=$?"
for m in frames samples tiles; do tail -3 gpurun_out/${tag}_${m}_n$N.err | cut -c1-300; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${m}_n$N.json"))
    keys = ("mode", "value", "ms_per_step", "frames_per_s", "scaling", "matches_single_gpu_frame", "max_abs_diff_8bit", "collective", "e2e")
    print("$m", {k: d[k] for k in keys if k in d})
except Exception as e:
    print("$m: no line", e)
PY
done
