"""
Drop-in for the hot functions of the reference's `moonrtx/data_loader.py`:

    load_elevation_data(filepath, downscale) -> (float32 ndarray, radius_scale)   [:166-247]
    load_color_data(filepath, gamma=2.2, downscale=1) -> uint8 RGBA ndarray       [:290-342]
    downscale_cache_available(filepath, downscale) -> bool                        [:63-86]

Same names, arguments, return values, exceptions and on-disk cache format
(`<src>.ds<N>.npy` + `.json` sidecar, data_loader.py:19-95) - but the block-mean
downscale + normalisation and the colour reduce + LUT run on the B200
(`mrtx_downscale_i16`, `mrtx_color_reduce_lut`), bit-exact against the numpy /
OpenCV results.  There is no CPU fallback: without the CUDA library these raise.

The array-level entry points (`downscale_elevation`, `color_texture`) are what
the file-level functions call after decoding, and what the tests and bench use.
"""

import ctypes as C
import json
import os
from typing import Optional

import numpy as np

from . import _lib
from .device import Device, DeviceBuffer, get_device

# data_loader.py:160-163, 253, 266-267
LDEM_METERS_PER_UNIT = 0.5
MOON_REFERENCE_RADIUS_M = 1_737_400.0
COLOR_DOWNSCALE_FACTORS = (1, 2, 4, 8)
COLOR_ALBEDO_MIN = 0.2
COLOR_ALBEDO_RANGE = 0.75
_CACHE_VERSION = 1          # data_loader.py:17 - same number, so caches are interchangeable


# --------------------------------------------------------------------------------------
# array level (device work)
# --------------------------------------------------------------------------------------
def downscale_elevation(src_i16: np.ndarray, downscale: int,
                        device: Optional[Device] = None) -> tuple[np.ndarray, float]:
    """
    int16 LDEM counts (H, W) -> (float32 displacement factors (H/ds, W/ds), radius_scale),
    the array part of load_elevation_data (data_loader.py:215-247).  Host buffers in and
    out; the copies both ways are part of the call.
    """
    src = np.ascontiguousarray(src_i16)
    if src.dtype == np.uint16:
        src = src.view(np.int16)            # data_loader.py:215 reinterprets in place
    if src.dtype != np.int16 or src.ndim != 2:
        raise ValueError("elevation source must be a 2-D int16/uint16 array")
    H, W = src.shape
    ds = int(downscale)
    if ds < 1:
        raise ValueError("downscale must be >= 1")
    dev = device or get_device()
    out = np.empty((H // ds, W // ds), dtype=np.float32)
    rs = C.c_float()
    _lib.check(dev.lib.mrtx_downscale_i16(dev.ctx, src.ctypes.data, W, H, ds, out.ctypes.data, C.byref(rs)))
    return out, float(rs.value)


def downscale_elevation_dev(src_dev: DeviceBuffer, W: int, H: int, downscale: int,
                            out_dev: Optional[DeviceBuffer] = None,
                            want_scale: bool = True) -> tuple[DeviceBuffer, Optional[float]]:
    """Same, HBM to HBM (no copies): the kernel-only path bench.py times."""
    dev = src_dev.dev
    ds = int(downscale)
    if out_dev is None:
        out_dev = dev.alloc((W // max(ds, 1)) * (H // max(ds, 1)) * 4)
    rs = C.c_float()
    _lib.check(dev.lib.mrtx_downscale_i16_dev(dev.ctx, src_dev.ptr, W, H, ds, out_dev.ptr,
                                              C.byref(rs) if want_scale else None))
    return out_dev, (float(rs.value) if want_scale else None)


def albedo_lut(gamma: float) -> np.ndarray:
    """The 256-entry albedo/gamma table of data_loader.py:272-287 (host; 256 elements)."""
    lut = np.arange(256, dtype=np.float32)
    lut = COLOR_ALBEDO_MIN + (COLOR_ALBEDO_RANGE / 255) * lut
    lut = np.power(lut, gamma, dtype=np.float32)
    lut *= 255
    return lut.astype(np.uint8)


def color_texture(bgr: np.ndarray, gamma: float = 2.2, downscale: int = 1,
                  device: Optional[Device] = None) -> np.ndarray:
    """
    uint8 BGR (H, W, 3) as cv2.imread returns it -> uint8 RGBA texture (H/k, W/k, 4):
    the reduce of cv2's IMREAD_REDUCED_COLOR_k (data_loader.py:331) fused with
    _moon_texture (data_loader.py:345-368).
    """
    src = np.ascontiguousarray(bgr)
    if src.dtype != np.uint8 or src.ndim != 3 or src.shape[2] != 3:
        raise ValueError("color source must be a (H, W, 3) uint8 array")
    k = int(downscale)
    if k not in COLOR_DOWNSCALE_FACTORS:
        raise ValueError(f"color downscale must be one of {COLOR_DOWNSCALE_FACTORS}")
    return _color_reduce_lut(src, albedo_lut(gamma), k, device or get_device())


def _color_reduce_lut(src: np.ndarray, lut: np.ndarray, k: int, dev: Device) -> np.ndarray:
    H, W = src.shape[:2]
    out = np.empty((H // k, W // k, 4), dtype=np.uint8)
    _lib.check(dev.lib.mrtx_color_reduce_lut(dev.ctx, src.ctypes.data, W, H, k, lut.ctypes.data, out.ctypes.data))
    return out


def resize_cubic(img: np.ndarray, width: int, height: int, device: Optional[Device] = None) -> np.ndarray:
    """cv2.resize(img, (width, height), interpolation=cv2.INTER_CUBIC) of a float32 (H, W, C) image followed by the clip
    to [0, 1] - the two lines of load_starmap that touch every pixel (data_loader.py:412-415) - on the device."""
    src = np.ascontiguousarray(img, dtype=np.float32)
    if src.ndim == 2:
        src = src[..., None]
    if src.ndim != 3 or not (1 <= src.shape[2] <= 4):
        raise ValueError("image must be (H, W) or (H, W, C <= 4) float32")
    dev = device or get_device()
    out = np.empty((int(height), int(width), src.shape[2]), dtype=np.float32)
    _lib.check(dev.lib.mrtx_resize_cubic_f32(dev.ctx, src.ctypes.data, src.shape[1], src.shape[0], src.shape[2],
                                             out.ctypes.data, int(width), int(height)))
    return out


# --------------------------------------------------------------------------------------
# file level (same behaviour as the reference functions)
# --------------------------------------------------------------------------------------
def _fingerprint(filepath: str, **params) -> dict:
    fp = {"version": _CACHE_VERSION, **params}
    if os.path.isfile(filepath):
        fp["source_size"] = os.path.getsize(filepath)
        fp["source_mtime"] = int(os.path.getmtime(filepath))
    return fp


def _cache_meta(base: str, fingerprint: dict) -> Optional[dict]:
    try:
        with open(base + ".json", "r", encoding="utf-8") as f:
            meta = json.load(f)
    except Exception:
        return None
    for key, value in fingerprint.items():
        if meta.get(key) != value:
            return None
    return meta if os.path.isfile(base + ".npy") else None


def _cache_load(base: str, fingerprint: dict):
    meta = _cache_meta(base, fingerprint)
    if meta is None:
        return None, {}
    try:
        return np.load(base + ".npy"), meta
    except Exception:
        return None, {}


def _cache_save(base: str, array: np.ndarray, meta: dict) -> None:
    try:
        np.save(base + ".npy", array)
        with open(base + ".json", "w", encoding="utf-8") as f:
            json.dump(meta, f)
        print(f"  Cached to {base}.npy for faster next start")
    except Exception as e:                      # a broken cache may only cost time
        print(f"Warning: could not write cache {base}.npy: {e}")


def downscale_cache_available(filepath: str, downscale: int) -> bool:
    if downscale <= 1:
        return False
    return _cache_meta(f"{filepath}.ds{downscale}", _fingerprint(filepath, downscale=downscale)) is not None


def read_image(filepath: str) -> Optional[np.ndarray]:
    """
    Decode the LDEM TIFF to a 2-D 16-bit array.  The reference uses
    plotoptix.utils.read_image (data_loader.py:206); when PlotOptiX is not installed
    OpenCV's TIFF reader is used (set OPENCV_IO_MAX_IMAGE_PIXELS for the 4.2 Gpx map).
    """
    try:
        from plotoptix.utils import read_image as _po_read      # type: ignore
        return _po_read(filepath)
    except ImportError:
        import cv2
        return cv2.imread(filepath, cv2.IMREAD_UNCHANGED)


def tiff_info(filepath: str) -> Optional[dict]:
    """Size of a TIFF and whether its strips can be streamed to the device as they lie (mrtx_tiff_info: an uncompressed
    little-endian single-channel 16-bit strip TIFF, classic or BigTIFF - the LDEM as NASA ships it).  Host work only."""
    w, h, bits, ok = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    if _lib.load().mrtx_tiff_info(os.fsencode(filepath), C.byref(w), C.byref(h), C.byref(bits), C.byref(ok)) != 0:
        return None
    return {"width": w.value, "height": h.value, "bits": bits.value, "streamable": bool(ok.value)}


def downscale_elevation_file(filepath: str, downscale: int, cache_npy: Optional[str] = None,
                             device: Optional[Device] = None) -> tuple[np.ndarray, float]:
    """
    read_image + block mean + normalise of load_elevation_data (data_loader.py:206-242) in one pass over the FILE: the
    strips of the TIFF go through pinned staging buffers straight to the device (no decoded copy of the 8.5 GB map in
    host memory) and the result is written to `cache_npy` in .npy format while it comes down.  ValueError if the file is
    not laid out for that (tiff_info()["streamable"]).
    """
    info = tiff_info(filepath)
    if info is None or not info["streamable"]:
        raise ValueError(f"{filepath}: not an uncompressed 16-bit strip TIFF")
    ds = int(downscale)
    dev = device or get_device()
    out = np.empty((info["height"] // ds, info["width"] // ds), dtype=np.float32)
    rs = C.c_float()
    _lib.check(dev.lib.mrtx_downscale_tiff_i16(dev.ctx, os.fsencode(filepath), ds, out.ctypes.data, C.byref(rs),
                                               os.fsencode(cache_npy) if cache_npy else None))
    return out, float(rs.value)


def load_elevation_data(filepath: str, downscale: int) -> tuple[np.ndarray, float]:
    print(f"Loading elevation data from {filepath}...")
    base = f"{filepath}.ds{downscale}"
    fingerprint = None
    if downscale > 1:
        fingerprint = _fingerprint(filepath, downscale=downscale)
        elevation, meta = _cache_load(base, fingerprint)
        if elevation is not None:
            print(f"  Loaded from cache: {base}.npy, dimensions {elevation.shape}")
            return elevation, float(meta["radius_scale"])
    if not os.path.isfile(filepath):
        raise FileNotFoundError(
            f"Elevation file not found: {filepath}, and no cache of it downscaled by {downscale} beside it.")
    info = tiff_info(filepath)
    if info is not None and info["streamable"] and info["height"] % downscale == 0 and info["width"] % downscale == 0:
        # the file's strips are the array: streamed to the device, the cache written as the result comes down
        print(f"  Original dimensions: {(info['height'], info['width'])}")
        try:
            elevation, radius_scale = downscale_elevation_file(filepath, downscale, base + ".npy" if fingerprint is not None else None)
        except _lib.MoonB200Error as e:
            if fingerprint is None or "cache" not in str(e) and "write" not in str(e):
                raise
            print(f"Warning: could not write cache {base}.npy: {e}")        # a broken cache may only cost time
            elevation, radius_scale = downscale_elevation_file(filepath, downscale, None)
            fingerprint = None
        print(f"  Downscaled dimensions: {elevation.shape}")
        if fingerprint is not None:
            try:
                with open(base + ".json", "w", encoding="utf-8") as f:
                    json.dump({**fingerprint, "radius_scale": radius_scale}, f)
                print(f"  Cached to {base}.npy for faster next start")
            except Exception as e:
                print(f"Warning: could not write cache {base}.npy: {e}")
        return elevation, radius_scale
    src = read_image(filepath)
    if src is None:
        raise ValueError(f"Failed to read elevation file: {filepath}")
    print(f"  Original dimensions: {src.shape}")
    elevation, radius_scale = downscale_elevation(src, downscale)
    print(f"  Downscaled dimensions: {elevation.shape}")
    if fingerprint is not None:
        _cache_save(base, elevation, {**fingerprint, "radius_scale": radius_scale})
    return elevation, radius_scale


def load_color_data(filepath: str, gamma: float = 2.2, downscale: int = 1) -> np.ndarray:
    print(f"Loading color data from {filepath}...")
    base = f"{filepath}.ds{downscale}"
    fingerprint = None
    if downscale > 1:
        fingerprint = _fingerprint(filepath, downscale=downscale)
        reduced, _ = _cache_load(base, fingerprint)
        if reduced is not None:
            # the reference caches the reduced BGR image and applies gamma after reading it
            print(f"  Loaded from cache: {base}.npy, dimensions {reduced.shape}")
            return color_texture(reduced, gamma, 1)
    if not os.path.isfile(filepath):
        raise FileNotFoundError(
            f"Color file not found: {filepath}, and no cache of it downscaled by {downscale} beside it.")
    import cv2
    k = downscale if downscale in COLOR_DOWNSCALE_FACTORS else 1        # data_loader.py:331 falls back to IMREAD_COLOR
    reduced = None
    if k > 1 and filepath.lower().endswith((".tif", ".tiff")):
        # a TIFF is decoded in full by OpenCV and reduced afterwards (central 2x2 of every k x k block):
        # that reduce runs on the GPU, once, to BGR - the image the reference caches
        src = cv2.imread(filepath, cv2.IMREAD_COLOR)
        if src is None:
            raise ValueError(f"Failed to read color file: {filepath}")
        if src.shape[0] % k == 0 and src.shape[1] % k == 0:
            ident = _color_reduce_lut(np.ascontiguousarray(src), np.arange(256, dtype=np.uint8), k, get_device())
            reduced = np.ascontiguousarray(ident[..., 2::-1])
        del src
    if reduced is None:
        # other formats are reduced inside their codecs (JPEG by DCT scaling), and sizes the factor does not divide are
        # resampled with other weights: the reference's own call, on the host, keeps those bit-exact too
        flag = {2: cv2.IMREAD_REDUCED_COLOR_2, 4: cv2.IMREAD_REDUCED_COLOR_4, 8: cv2.IMREAD_REDUCED_COLOR_8}.get(downscale, cv2.IMREAD_COLOR)
        reduced = cv2.imread(filepath, flag)
        if reduced is None:
            raise ValueError(f"Failed to read color file: {filepath}")
    print(f"  Dimensions: {reduced.shape}" + (f" (decoded at 1/{downscale})" if downscale > 1 else ""))
    if fingerprint is not None:
        _cache_save(base, reduced, fingerprint)         # (also for a factor OpenCV has no flag for: the full image, as the reference does)
    return color_texture(reduced, gamma, 1)


def load_starmap(filepath: str, target_width: int) -> Optional[np.ndarray]:
    """
    data_loader.py:371-425: the star map for the background, float32 RGB (h, w, 3) in [0, 1], at most `target_width` wide;
    None if the file is missing or unreadable.  Same cache files (`<file>.w<width>.npy` + sidecar); the bicubic resize
    of the 16k source runs on the device.
    """
    if not os.path.isfile(filepath):
        print(f"Star map not found: {filepath}")
        return None
    print(f"Loading star map from {filepath}...")
    cache_base = f"{filepath}.w{target_width}"
    fingerprint = _fingerprint(filepath, target_width=target_width)
    star_map, _ = _cache_load(cache_base, fingerprint)
    if star_map is not None:
        print(f"  Loaded from cache: {cache_base}.npy, dimensions {star_map.shape}")
        return star_map
    import cv2
    star_src = cv2.imread(filepath)
    if star_src is None:
        print(f"Failed to read star map: {filepath}")
        return None
    star_src = star_src[..., ::-1].astype(np.float32)
    star_src *= 1 / 255
    if target_width < star_src.shape[1]:
        target_height = int(star_src.shape[0] * target_width / star_src.shape[1])
        star_map = resize_cubic(star_src, target_width, target_height)
    else:
        star_map = star_src
    print(f"  Dimensions: {star_map.shape}")
    _cache_save(cache_base, star_map, fingerprint)
    return star_map
