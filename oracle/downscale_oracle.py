"""
TEST INFRASTRUCTURE - CPU oracle for the data_loader hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (moonrtx_b200) never does.

Pinned: every function here is checked bit-for-bit against the *unmodified*
reference `moonrtx/data_loader.py` imported under stubs (oracle/ref_stub.py) by
tests/test_oracle_pinned.py, and against the fixtures under tests/golden/ that
oracle/make_golden.py produced from the reference in the build container.

Two statements of the elevation path are kept:

* `block_mean_numpy`     - the reference's own numpy expression (data_loader.py:223-226),
                           used as the timed CPU baseline ("port" of a Python reference
                           is the expression itself);
* `block_mean_two_stage` - the arithmetic that expression performs, written out
                           (what the CUDA kernel implements), so the kernel is
                           checked against semantics and not against a numpy quirk.
"""

import numpy as np

# data_loader.py:160-163
LDEM_METERS_PER_UNIT = 0.5
MOON_REFERENCE_RADIUS_M = 1_737_400.0
# data_loader.py:266-267
COLOR_ALBEDO_MIN = 0.2
COLOR_ALBEDO_RANGE = 0.75


def block_mean_numpy(src_i16: np.ndarray, ds: int) -> np.ndarray:
    """data_loader.py:223-226 verbatim semantics (reshape raises ValueError when
    H or W is not a multiple of ds, like the reference)."""
    h = src_i16.shape[0] // ds
    w = src_i16.shape[1] // ds
    return src_i16.reshape(1, h, ds, w, ds).mean(4, dtype=np.float32).mean(2, dtype=np.float32).reshape(h, w)


def block_mean_two_stage(src_i16: np.ndarray, ds: int) -> np.ndarray:
    """
    What data_loader.py:225-226 computes, spelled out:
      stage 1: m[r, j]  = fl32( exact_sum(src[r, j*ds:(j+1)*ds]) / fl32(ds) )
      stage 2: out[i,j] = fl32( (((m[i*ds,j] + m[i*ds+1,j]) + ...) each add rounded to f32) / fl32(ds) )
    """
    H, W = src_i16.shape
    if H % ds or W % ds:
        raise ValueError(f"cannot reshape array of size {H * W} into blocks of {ds}")
    h, w = H // ds, W // ds
    dsf = np.float32(ds)
    # exact integer row sums (|sum| < 2^24 for ds <= 512, so fl32(sum) is exact)
    rows = src_i16.reshape(H, w, ds).astype(np.int64).sum(axis=2)
    m = rows.astype(np.float32) / dsf                       # one rounding
    m = m.reshape(h, ds, w)
    acc = m[:, 0, :].copy()
    for k in range(1, ds):
        acc = acc + m[:, k, :]                              # f32 add, sequential order
    return (acc / dsf).astype(np.float32)


def normalise(elev_f32: np.ndarray) -> tuple[np.ndarray, float]:
    """data_loader.py:216, 227, 232, 241-242: *scale, +1, /max - three f32 roundings."""
    scale = LDEM_METERS_PER_UNIT / MOON_REFERENCE_RADIUS_M   # python float, applied as f32
    e = elev_f32 * np.float32(scale)
    e = e + np.float32(1.0)
    radius_scale = float(e.max())
    e = e / np.float32(radius_scale)
    return e.astype(np.float32, copy=False), radius_scale


def load_elevation(src_i16: np.ndarray, ds: int, explicit: bool = False) -> tuple[np.ndarray, float]:
    """The array part of load_elevation_data (data_loader.py:215-247), cache/IO removed."""
    if ds == 1:
        e = src_i16.astype(np.float32)                       # data_loader.py:218-220
    else:
        e = block_mean_two_stage(src_i16, ds) if explicit else block_mean_numpy(src_i16, ds)
    return normalise(e)


def albedo_lut(gamma: float) -> np.ndarray:
    """data_loader.py:272-287."""
    lut = np.arange(256, dtype=np.float32)
    lut = COLOR_ALBEDO_MIN + (COLOR_ALBEDO_RANGE / 255) * lut
    lut = np.power(lut, gamma, dtype=np.float32)
    lut *= 255
    return lut.astype(np.uint8)


def color_reduce(bgr: np.ndarray, k: int) -> np.ndarray:
    """
    cv2.imread(path, IMREAD_REDUCED_COLOR_k) for a TIFF (data_loader.py:250-259, 331):
    OpenCV 4.13 decodes in full, then resizes with INTER_LINEAR_EXACT to (W//k, H//k).
    For dimensions divisible by k that is, per channel, the round-half-up mean of the
    central 2x2 of each k x k block (pinned by tests against cv2).
    """
    if k == 1:
        return bgr
    H, W = bgr.shape[:2]
    if H % k or W % k:
        raise ValueError("color_reduce: dimensions must be divisible by the factor")
    o = k // 2 - 1
    b = bgr.reshape(H // k, k, W // k, k, 3).astype(np.uint16)
    s = b[:, o, :, o] + b[:, o, :, o + 1] + b[:, o + 1, :, o] + b[:, o + 1, :, o + 1]
    return ((s + 2) >> 2).astype(np.uint8)


def moon_texture(bgr: np.ndarray, gamma: float) -> np.ndarray:
    """data_loader.py:345-368: RGBA = (lut[R], lut[G], lut[B], 255) from a BGR source."""
    lut = albedo_lut(gamma)
    out = np.empty(bgr.shape[:2] + (4,), dtype=np.uint8)
    out[..., 0] = lut[bgr[..., 2]]
    out[..., 1] = lut[bgr[..., 1]]
    out[..., 2] = lut[bgr[..., 0]]
    out[..., 3] = 255
    return out


def load_color(bgr: np.ndarray, gamma: float = 2.2, k: int = 1) -> np.ndarray:
    return moon_texture(color_reduce(bgr, k), gamma)
