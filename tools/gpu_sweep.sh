#!/bin/bash
# Development helper: kernel-only timing of config 3 (and others) for the main library and every variant library in
# moonrtx_b200/_variants, each with the engine-switch VARIANTS of tools/bench_trace.py.  One gpurun call.
#   CFGS="cfg3:16 cfg5:16" VARIANTS="beam=0,0;beam=1,2" bash tools/gpu_sweep.sh tag
tag=${1:-sweep}; mkdir -p gpurun_out
CFGS=${CFGS:-cfg3:16}
for c in $CFGS; do
  cfg=${c%%:*}; spp=${c##*:}
  for f in main moonrtx_b200/_variants/*.so; do
    [ "$f" != main ] && [ ! -f "$f" ] && continue
    name=$(basename $f .so)
    if [ "$f" = main ]; then timeout 900 python tools/bench_trace.py $cfg $spp > gpurun_out/${tag}_${cfg}_${name}.log 2>&1
    else MRTX_LIB=$f timeout 900 python tools/bench_trace.py $cfg $spp > gpurun_out/${tag}_${cfg}_${name}.log 2>&1; fi
  done
done
python - "$tag" <<'PY'
import glob, json, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_*.log")):
    print("==", f)
    for line in open(f):
        if not line.startswith("{"):
            if "rror" in line: print(line.strip()[:300])
            continue
        d = json.loads(line)
        if "compare" in d: print("   cmp", d["compare"], "max", d["img_max"], "px", d["pixels_differ"], "gt2", d["pixels_differ_gt2"], "accmax", round(d["accum_max_abs"], 4), d["accum_w_equal"])
        elif "ms" in d: print(f'{str(d["kernel"]):34s} ms {d["ms"]:7.2f}  nodes/ray {d["nodes_per_inray"]:5.1f}  tests {d.get("patch_tests")}  defer {d["defer"]["deferred_samples"]}  {d.get("kernel_ms")}  hits {d["primary_hits"]} occl {d["shadow_occluded"]}')
PY
