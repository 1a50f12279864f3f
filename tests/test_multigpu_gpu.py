"""2-GPU checks of the frame-sharding collectives; skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _gpu_count():
    import ctypes as C
    try:
        from moonrtx_b200 import _lib
        n = C.c_int()
        if _lib.load().mrtx_device_count(C.byref(n)) != 0:
            return 0
        return n.value
    except Exception:                       # library not built / no driver: the CPU suite must still collect
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sample_split_and_row_split_match_single_gpu():
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(here, "mgpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
