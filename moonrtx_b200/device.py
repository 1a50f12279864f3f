"""
One `Device` per GPU: owns the C-ABI context (stream, scratch, scene, frame
buffers) and gives numpy-facing helpers for device buffers.  Host code above this
never touches CUDA directly; PyTorch is not needed here at all.
"""

import ctypes as C
import os
import threading

import numpy as np

from . import _lib


class DeviceBuffer:
    """A raw allocation in HBM, freed with the object."""

    def __init__(self, dev: "Device", nbytes: int):
        self.dev = dev
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _lib.check(dev.lib.mrtx_dev_alloc(dev.ctx, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, array: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(array)
        if a.nbytes > self.nbytes:
            raise ValueError("array larger than the device buffer")
        _lib.check(self.dev.lib.mrtx_h2d(self.dev.ctx, self.ptr, a.ctypes.data, a.nbytes))
        self.dev.synchronize()          # `a` may be a temporary
        return self

    def download(self, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        if out.nbytes > self.nbytes:
            raise ValueError("request larger than the device buffer")
        _lib.check(self.dev.lib.mrtx_d2h(self.dev.ctx, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.dev.lib.mrtx_dev_free(self.dev.ctx, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Device:
    def __init__(self, index: int = 0):
        self.lib = _lib.load()
        self.index = int(index)
        ctx = _lib.c_ctx()
        _lib.check(self.lib.mrtx_create(self.index, C.byref(ctx)))
        self.ctx = ctx
        sm, l2, hbm = C.c_int(), C.c_int(), C.c_size_t()
        _lib.check(self.lib.mrtx_device_props(ctx, C.byref(sm), C.byref(l2), C.byref(hbm)))
        self.sm_count, self.l2_bytes, self.hbm_bytes = sm.value, l2.value, hbm.value

    # -- memory ---------------------------------------------------------------
    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def to_device(self, array: np.ndarray) -> DeviceBuffer:
        a = np.ascontiguousarray(array)
        return self.alloc(a.nbytes).upload(a)

    def synchronize(self):
        _lib.check(self.lib.mrtx_synchronize(self.ctx))

    def use_stream(self, cuda_stream_handle):
        """Issue this device's work on a caller-owned stream (int handle), None = own."""
        _lib.check(self.lib.mrtx_set_stream(self.ctx, C.c_void_p(cuda_stream_handle or 0)))

    def l2_flush(self):
        _lib.check(self.lib.mrtx_l2_flush(self.ctx))

    # -- CUDA-event stopwatch on the device stream ------------------------------------
    def timer_start(self):
        _lib.check(self.lib.mrtx_timer_start(self.ctx))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _lib.check(self.lib.mrtx_timer_stop(self.ctx, C.byref(ms)))
        return float(ms.value)

    def close(self):
        if self.ctx:
            self.lib.mrtx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_devices: dict[int, Device] = {}
_devices_lock = threading.Lock()


def default_index() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def get_device(index: int | None = None) -> Device:
    """Process-wide Device for a GPU (default: LOCAL_RANK, one process per GPU)."""
    i = default_index() if index is None else int(index)
    with _devices_lock:
        d = _devices.get(i)
        if d is None or not d.ctx:
            d = Device(i)
            _devices[i] = d
        return d
