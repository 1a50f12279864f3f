#!/bin/bash
# N = 1 lines for profiles/: bench.py as the driver runs it, the CPU arm, the downscale kernels with hot L2 (ncu --cache-control none)
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r06_bench_n1.json 2> gpurun_out/r06_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r06_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r06_reference_n1.json 2> gpurun_out/r06_reference_n1.err; echo "ref rc=$?"; tail -3 gpurun_out/r06_reference_n1.err
timeout 600 ncu --set full --clock-control none --cache-control none -k regex:'downscale_vec_kernel|normalise_kernel' -s 6 -c 2 -o gpurun_out/r06_downscale_hotl2 -f python tools/bench_downscale.py > gpurun_out/r06_ncu_ds2.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/r06_bench_n1.json"))
print({k: d[k] for k in ("value", "value_incl_culled", "ms_per_step", "frames_per_s", "ms_per_frame_per_gpu", "gpu_launches")})
print(d["roofline"]); print(d["e2e"]); print(d["parity"]); print(d["cpu_baseline"]); print(json.dumps(d["downscale"])); print(d["clocks"])
print(open("gpurun_out/r06_reference_n1.json").read()[:900])
PY
