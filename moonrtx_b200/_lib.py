"""
ctypes binding of libmoonb200.so (the C ABI declared in include/moonb200.h).

There is no fallback: if the library has not been built, or a call fails, an
exception is raised.  Nothing in this package computes the hot path on the CPU.
"""

import ctypes as C
import os
import re
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmoonb200.so")
HEADER = os.path.join(HERE, "..", "include", "moonb200.h")


class MoonB200Error(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()

c_ctx = C.c_void_p
_D3 = C.POINTER(C.c_double)

_SIGNATURES = {
    # name: (restype, [argtypes])
    "mrtx_abi_version": (C.c_int, []),
    "mrtx_last_error": (C.c_char_p, []),
    "mrtx_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mrtx_create": (C.c_int, [C.c_int, C.POINTER(c_ctx)]),
    "mrtx_destroy": (C.c_int, [c_ctx]),
    "mrtx_synchronize": (C.c_int, [c_ctx]),
    "mrtx_set_stream": (C.c_int, [c_ctx, C.c_void_p]),
    "mrtx_get_stream": (C.c_int, [c_ctx, C.POINTER(C.c_void_p)]),
    "mrtx_device_props": (C.c_int, [c_ctx, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "mrtx_timer_start": (C.c_int, [c_ctx]),
    "mrtx_timer_stop": (C.c_int, [c_ctx, C.POINTER(C.c_float)]),
    "mrtx_dev_alloc": (C.c_int, [c_ctx, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mrtx_dev_free": (C.c_int, [c_ctx, C.c_void_p]),
    "mrtx_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "mrtx_host_free": (C.c_int, [C.c_void_p]),
    "mrtx_h2d": (C.c_int, [c_ctx, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mrtx_d2h": (C.c_int, [c_ctx, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mrtx_l2_flush": (C.c_int, [c_ctx]),
    "mrtx_downscale_i16": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float)]),
    "mrtx_downscale_i16_dev": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float)]),
    "mrtx_color_reduce_lut": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mrtx_color_reduce_lut_dev": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mrtx_synth_ldem_i16_dev": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_uint32]),
    "mrtx_synth_color_bgr_dev": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_uint32]),
    "mrtx_set_displacement_f32": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int]),
    "mrtx_set_displacement_f32_dev": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mrtx_set_displacement_i16": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]),
    "mrtx_set_displacement_i16_dev": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]),
    "mrtx_set_texture_rgba8": (C.c_int, [c_ctx, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "mrtx_set_frame": (C.c_int, [c_ctx, _D3, _D3, _D3, C.c_double]),
    "mrtx_set_camera": (C.c_int, [c_ctx, _D3, _D3, _D3, C.c_double]),
    "mrtx_set_light": (C.c_int, [c_ctx, _D3, C.c_double, C.c_double]),
    "mrtx_set_float": (C.c_int, [c_ctx, C.c_char_p, C.c_double]),
    "mrtx_set_uint": (C.c_int, [c_ctx, C.c_char_p, C.c_uint, C.c_uint]),
    "mrtx_resize": (C.c_int, [c_ctx, C.c_int, C.c_int]),
    "mrtx_render": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int]),
    "mrtx_resolve": (C.c_int, [c_ctx]),
    "mrtx_frame_submit": (C.c_int, [c_ctx, C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_int)]),
    "mrtx_frame_wait": (C.c_int, [c_ctx, C.c_int]),
    "mrtx_frame_submit_to": (C.c_int, [c_ctx, C.c_void_p, C.c_uint, C.c_int, C.POINTER(C.c_int)]),
    "mrtx_frame_recv": (C.c_int, [c_ctx, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "mrtx_frame_recv_wait": (C.c_int, [c_ctx, C.c_int]),
    "mrtx_tiff_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mrtx_downscale_tiff_i16": (C.c_int, [c_ctx, C.c_char_p, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_char_p]),
    "mrtx_set_tubes": (C.c_int, [c_ctx, C.c_void_p, C.c_int]),
    "mrtx_p2p_open": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_uint8)]),
    "mrtx_p2p_connect": (C.c_int, [c_ctx, C.POINTER(C.c_uint8)]),
    "mrtx_p2p_close": (C.c_int, [c_ctx]),
    "mrtx_read_rgba8": (C.c_int, [c_ctx, C.c_void_p]),
    "mrtx_read_accum_f32": (C.c_int, [c_ctx, C.c_void_p]),
    "mrtx_read_hit_f32": (C.c_int, [c_ctx, C.c_void_p]),
    "mrtx_read_hit_f64": (C.c_int, [c_ctx, C.c_void_p]),
    "mrtx_hit_at": (C.c_int, [c_ctx, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "mrtx_frame_buffers_dev": (C.c_int, [c_ctx, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "mrtx_counters": (C.c_int, [c_ctx, C.POINTER(C.c_uint64), C.c_int]),
    "mrtx_set_background_f32": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_float]),
    "mrtx_read_background_rgba8": (C.c_int, [c_ctx, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mrtx_set_sun_disk": (C.c_int, [c_ctx, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_float)]),
    "mrtx_resize_cubic_f32": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "mrtx_render_tiles": (C.c_int, [c_ctx, C.c_int, C.c_uint, C.c_uint, C.c_int]),
    "mrtx_allgather_tiles": (C.c_int, [c_ctx, C.c_int]),
    "mrtx_kernel_times": (C.c_int, [c_ctx, C.POINTER(C.c_double), C.c_int]),
    "mrtx_defer_stats": (C.c_int, [c_ctx, C.POINTER(C.c_uint64), C.c_int]),
    "mrtx_comm_unique_id": (C.c_int, [C.c_char_p, C.c_void_p]),
    "mrtx_comm_init": (C.c_int, [c_ctx, C.c_char_p, C.c_int, C.c_int, C.c_void_p]),
    "mrtx_comm_destroy": (C.c_int, [c_ctx]),
    "mrtx_allreduce_accum": (C.c_int, [c_ctx]),
    "mrtx_allgather_rows": (C.c_int, [c_ctx, C.c_int]),
}

MRTX_ERR_INVALID = -1


def header_symbols() -> list[str]:
    """Every function include/moonb200.h declares (the tests check the .so exports them all)."""
    with open(HEADER, "r", encoding="utf-8") as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrtx_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """The loaded library; raises MoonB200Error when it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise MoonB200Error(
                f"{LIB_PATH} is missing: build it with `python -m moonrtx_b200.build` "
                "(there is no CPU fallback for the hot path)")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.mrtx_abi_version() != 1:
            raise MoonB200Error("libmoonb200.so ABI version mismatch; rebuild it")
        _lib = lib
        return lib


def check(rc: int) -> None:
    """Raise for a non-zero return code: ValueError for bad arguments (what numpy /
    the reference raise for the same mistakes), MoonB200Error for everything else."""
    if rc == 0:
        return
    msg = load().mrtx_last_error().decode("utf-8", "replace")
    if rc == MRTX_ERR_INVALID:
        raise ValueError(msg)
    raise MoonB200Error(f"libmoonb200 error {rc}: {msg}")


def vec3(v) -> C.Array:
    a = [float(x) for x in v]
    if len(a) != 3:
        raise ValueError("expected a 3-vector")
    return (C.c_double * 3)(*a)


def nccl_library_path() -> str:
    """The torch-bundled libnccl.so.2 (no torch import needed to find it)."""
    import importlib.util
    spec = importlib.util.find_spec("nvidia.nccl")
    if spec is not None and spec.submodule_search_locations:
        p = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
        if os.path.isfile(p):
            return p
    return "libnccl.so.2"
