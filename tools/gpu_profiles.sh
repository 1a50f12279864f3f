#!/bin/bash
# Round-2 evidence for profiles/: bench line, launch list of the same command, `ncu --set full` of the two walk kernels
tag=${1:-r06}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-downscale > gpurun_out/${tag}_bench_short.json 2> gpurun_out/${tag}_bench_short.err || { tail -5 gpurun_out/${tag}_bench_short.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-downscale > gpurun_out/${tag}_launches.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'trace_kernel_fast|shadow_kernel' -s 4 -c 2 -o gpurun_out/${tag}_walk -f python tools/bench_trace.py cfg3 16 > gpurun_out/${tag}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'downscale_vec_kernel|normalise_kernel' -s 6 -c 2 -o gpurun_out/${tag}_downscale -f python tools/bench_downscale.py > gpurun_out/${tag}_ncu_ds.log 2>&1
ls -la gpurun_out/${tag}_*
