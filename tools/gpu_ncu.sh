#!/bin/bash
# Development helper: one `ncu --set full` capture of the dominant kernel of config 3 (after a plain run has exited 0).
#   bash tools/gpu_ncu.sh tag [kernel-regex] [launch-skip]
tag=${1:-r05}; k=${2:-trace_kernel_fast}; skip=${3:-1}
mkdir -p gpurun_out
timeout 600 python tools/bench_trace.py cfg3 16 > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
grep '"spp": 16' gpurun_out/${tag}_plain.log | cut -c1-120
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:$k -s $skip -c 1 -o gpurun_out/${tag}_fast -f python tools/bench_trace.py cfg3 16 > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/${tag}_fast.ncu-rep
