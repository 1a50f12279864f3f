( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r5d_tests.log
VARIANTS='shadow_queue=0+beam=0,0+ceiling=0;shadow_queue=1+beam=0,0+ceiling=0;shadow_queue=1+beam=0,0+ceiling=2;shadow_queue=1+beam=1,2+ceiling=2;shadow_queue=1+beam=0,0+ceiling=3' bash tools/gpu_sweep.sh r5d
cat gpurun_out/r5d_tests.log
