#!/bin/bash
# Development helper: time kernel 3 on config 3 (16 spp) for the main library and every variant in moonrtx_b200/_variants
out=${1:-gpurun_out/sweep.log}; : > $out
echo "== main" >> $out; KERNELS=${KERNELS:-3} python tools/bench_trace.py cfg3 16 2>&1 | grep '"spp"' | cut -c1-40 >> $out
for f in moonrtx_b200/_variants/*.so; do
  echo "== $f" >> $out; MRTX_LIB=$f KERNELS=${KERNELS:-3} python tools/bench_trace.py cfg3 16 2>&1 | grep '"spp"\|Error' | cut -c1-40 >> $out
done
cat $out
