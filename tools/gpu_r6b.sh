mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r6b_tests.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r6b_bench.json 2> gpurun_out/r6b_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r6b_bench.err; cat gpurun_out/r6b_tests.log
python -c "
import json
d=json.load(open('gpurun_out/r6b_bench.json'))
for k in ('value','value_incl_culled','ms_per_step','frames_per_s','ms_per_frame_per_gpu','gpu_launches'): print(k, d[k])
print('roofline', json.dumps(d['roofline'], indent=1))
print('e2e', d['e2e']); print('parity', d['parity']); print('cpu', d['cpu_baseline']); print('downscale', json.dumps(d['downscale'], indent=1)); print(d['clocks'])
"
