#!/bin/bash
# Evidence of a round for profiles/ (one gpurun call, one GPU): the full GPU test suite, bench.py as the driver runs it,
# the CPU arm, the launch list of a short bench run, `ncu --set full` of the production kernels and of the downscale kernels.
#   bash tools/gpu_round.sh r07
tag=${1:-r07}; out=gpurun_out; mkdir -p $out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $out/${tag}_gpu_tests.log; tail -3 $out/${tag}_gpu_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1 | tee $out/${tag}_smoke.log
timeout 1500 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?"; tail -2 $out/${tag}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_reference_n1.json 2> $out/${tag}_reference_n1.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-downscale > $out/${tag}_launches.log 2>&1
timeout 1300 ncu --set full --import-source on --clock-control none -k regex:'trace_kernel_pool|shade_kernel|shadow_kernel' -s 6 -c 3 -o $out/${tag}_walk -f python tools/bench_trace.py cfg3 16 > $out/${tag}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'downscale_vec_kernel|normalise_kernel' -s 6 -c 2 -o $out/${tag}_downscale -f python tools/bench_downscale.py > $out/${tag}_ncu_ds.log 2>&1
python - "$tag" <<'PY'
import json, sys
t = sys.argv[1]
d = json.load(open(f"gpurun_out/{t}_bench_n1.json"))
print({k: d[k] for k in ("value", "value_incl_culled", "ms_per_step", "frames_per_s", "ms_per_frame_per_gpu", "gpu_launches")})
print(d["roofline"]); print(d["e2e"]); print(d["parity"]); print(d["cpu_baseline"]); print(d["interreflection"]); print(json.dumps(d["downscale"])); print(d["clocks"])
print(open(f"gpurun_out/{t}_reference_n1.json").read()[:900])
PY
ls -la $out/${tag}_*
