#!/bin/bash
# Development helper: `ncu --set full` of the two production kernels of one config-3 frame (first launches after warm-up).
tag=${1:-r05}
mkdir -p gpurun_out
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'trace_kernel_fast|shadow_kernel' -s 4 -c 2 -o gpurun_out/${tag}_k -f python tools/bench_trace.py cfg3 16 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/${tag}_k.ncu-rep
