"""
Parity of the CUDA data_loader path (through the C ABI) against the reference fixtures
and the pinned oracle.  Bit-exact: float32 results are compared as uint32 bit patterns.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import downscale_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dl():
    from moonrtx_b200 import data_loader
    return data_loader


@pytest.fixture(scope="module")
def elev_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "elevation_small.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", ["synth", "uniform"])
@pytest.mark.parametrize("ds", [1, 2, 3, 4, 5, 6, 9, 12])
def test_downscale_matches_reference_fixture(dl, elev_golden, name, ds):
    src = elev_golden[f"{name}_src"]                # 360 x 180: W % 16 != 0 -> generic kernel
    e, rs = dl.downscale_elevation(src, ds)
    ref = elev_golden[f"{name}_ds{ds}"]
    assert e.shape == ref.shape and e.dtype == np.float32
    assert np.array_equal(bits(e), bits(ref))
    assert rs == float(elev_golden[f"{name}_ds{ds}_rs"])


def test_downscale_matches_reference_digests(dl, elev_golden, golden_dir):
    with open(os.path.join(golden_dir, "elevation_digests.json")) as f:
        dig = json.load(f)
    big = np.tile(elev_golden["synth_src"], (8, 8))         # 2880 x 1440: vector kernels
    big = (big.astype(np.int32) + (np.arange(big.shape[1])[None, :] % 97) * 3
           - (np.arange(big.shape[0])[:, None] % 89) * 5).astype(np.int16)
    for ds, case in dig["cases"].items():
        e, rs = dl.downscale_elevation(big, int(ds))
        assert list(e.shape) == case["shape"]
        assert hashlib.sha256(e.tobytes()).hexdigest() == case["sha256"], f"ds={ds}"
        assert rs == case["radius_scale"]


@pytest.mark.parametrize("ds,W,H", [
    (1, 1024, 512), (2, 2048, 1024), (3, 3072, 1536), (4, 4096, 2048), (5, 2560, 1280),
    (6, 3072, 1536), (7, 1792, 896), (8, 4096, 2048), (12, 3072, 1536), (16, 4096, 2048),
    (32, 4096, 2048), (64, 4096, 2048), (4, 4104, 2052),        # w % 8 != 0 -> generic
    (3, 1000 * 3, 7 * 3), (512, 1024, 512),
])
def test_downscale_matches_oracle_random(dl, ds, W, H):
    rng = np.random.default_rng(ds * 1000 + W)
    src = rng.integers(-32768, 32768, size=(H, W), dtype=np.int32).astype(np.int16)
    e, rs = dl.downscale_elevation(src, ds)
    ref, rs_ref = orc.load_elevation(src, ds, explicit=True)
    assert np.array_equal(bits(e), bits(ref))
    assert rs == rs_ref
    assert e.max() == np.float32(1.0)


def test_downscale_full_config1_size_bit_exact(dl):
    """BASELINE config 1: 23040 x 11520 int16, ds = 4, against the reference's numpy expression."""
    from moonrtx_b200.device import get_device
    import ctypes as C
    dev = get_device()
    W, H, ds = 23040, 11520, 4
    src_dev = dev.alloc(W * H * 2)
    from moonrtx_b200 import _lib
    _lib.check(dev.lib.mrtx_synth_ldem_i16_dev(dev.ctx, src_dev.ptr, W, H, 20240314))
    src = src_dev.download((H, W), np.int16)
    assert src.min() >= -18200 and src.max() <= 21600 and src.std() > 1000
    out_dev, rs = dl.downscale_elevation_dev(src_dev, W, H, ds)
    e = out_dev.download((H // ds, W // ds), np.float32)
    ref, rs_ref = orc.load_elevation(src, ds)                # numpy expression of data_loader.py:223-242
    assert np.array_equal(bits(e), bits(ref))
    assert rs == rs_ref
    # host-buffer entry point gives the same
    e2, rs2 = dl.downscale_elevation(src, ds)
    assert np.array_equal(bits(e2), bits(ref)) and rs2 == rs_ref


def test_downscale_rejects_like_numpy(dl):
    src = np.zeros((180, 360), dtype=np.int16)
    with pytest.raises(ValueError):
        dl.downscale_elevation(src, 7)                       # reshape ValueError in the reference
    with pytest.raises(ValueError):
        dl.downscale_elevation(src.astype(np.float32), 2)
    with pytest.raises(ValueError):
        dl.downscale_elevation(src, 0)


def test_downscale_uint16_view_and_constant_map(dl):
    src = np.full((64, 128), -5, dtype=np.int16)
    e, rs = dl.downscale_elevation(src.view(np.uint16), 4)   # the TIFF reader hands back uint16
    ref, rs_ref = orc.load_elevation(src, 4)
    assert np.array_equal(bits(e), bits(ref)) and rs == rs_ref
    assert np.all(e == 1.0)


def test_color_matches_reference_fixture(dl, golden_dir):
    g = np.load(os.path.join(golden_dir, "color_small.npz"))
    src = g["src_bgr"]
    for k in (1, 2, 4, 8):
        for gamma in (2.2, 1.0):
            tex = dl.color_texture(src, gamma, k)
            assert tex.dtype == np.uint8
            assert np.array_equal(tex, g[f"k{k}_g{gamma}"]), (k, gamma)


@pytest.mark.parametrize("k,W,H", [(1, 333, 77), (2, 1002, 334), (4, 2052, 1028), (8, 4104, 2056), (2, 8192, 4096)])
def test_color_matches_oracle_random(dl, k, W, H):
    rng = np.random.default_rng(k * 77 + W)
    src = rng.integers(0, 256, size=(H, W, 3), dtype=np.int32).astype(np.uint8)
    tex = dl.color_texture(src, 2.2, k)
    assert np.array_equal(tex, orc.load_color(src, 2.2, k))


def test_color_rejects_bad_arguments(dl):
    src = np.zeros((64, 66, 3), dtype=np.uint8)
    with pytest.raises(ValueError):
        dl.color_texture(src, 2.2, 3)
    with pytest.raises(ValueError):
        dl.color_texture(src, 2.2, 4)                        # 66 % 4 != 0
    with pytest.raises(ValueError):
        dl.color_texture(src[..., 0], 2.2, 1)


def test_file_level_functions_and_cache_format(dl, tmp_path):
    import cv2
    rng = np.random.default_rng(5)
    ldem = rng.integers(-18200, 21600, size=(96, 192), dtype=np.int32).astype(np.int16)
    p = str(tmp_path / "ldem.tif")
    cv2.imwrite(p, ldem.view(np.uint16))
    assert not dl.downscale_cache_available(p, 3)
    e, rs = dl.load_elevation_data(p, 3)
    ref, rs_ref = orc.load_elevation(ldem, 3)
    assert np.array_equal(bits(e), bits(ref)) and rs == rs_ref
    assert dl.downscale_cache_available(p, 3)
    meta = json.load(open(p + ".ds3.json"))
    assert meta["version"] == 1 and meta["downscale"] == 3 and meta["radius_scale"] == rs
    e2, rs2 = dl.load_elevation_data(p, 3)                   # from cache
    assert np.array_equal(e2, e) and rs2 == rs
    with pytest.raises(FileNotFoundError):
        dl.load_elevation_data(str(tmp_path / "missing.tif"), 2)

    col = rng.integers(0, 256, size=(64, 128, 3), dtype=np.int32).astype(np.uint8)
    pc = str(tmp_path / "color.tif")
    cv2.imwrite(pc, col)
    tex = dl.load_color_data(pc, 2.2, 4)
    assert np.array_equal(tex, orc.load_color(col, 2.2, 4))
    cached = np.load(pc + ".ds4.npy")
    assert np.array_equal(cached, orc.color_reduce(col, 4))  # the reference caches reduced BGR
    assert np.array_equal(dl.load_color_data(pc, 1.0, 4), orc.load_color(col, 1.0, 4))


def test_starmap_bicubic_resize_matches_opencv(tmp_path):
    """load_starmap (data_loader.py:371-425): BGR -> RGB / 255, cv2.resize(INTER_CUBIC) to the target width, clip - the
    resize on the device against OpenCV's own on the same image; and the file-level function with its cache."""
    import cv2
    from moonrtx_b200.data_loader import load_starmap, resize_cubic
    rng = np.random.default_rng(3)
    src = rng.uniform(0, 1, (257, 512, 3)).astype(np.float32)
    src[rng.integers(0, 257, 300), rng.integers(0, 512, 300)] = 1.0
    for (w, h) in ((128, 64), (200, 100), (511, 256), (37, 19)):
        want = np.clip(cv2.resize(src, (w, h), interpolation=cv2.INTER_CUBIC), 0, 1)
        got = resize_cubic(src, w, h)
        # (OpenCV's float path rounds its coefficients and sums differently: 2e-5 of full scale = 0.005 of an 8-bit step)
        assert got.shape == want.shape and float(np.abs(got - want).max()) <= 5e-5, float(np.abs(got - want).max())
    img8 = (rng.uniform(0, 1, (128, 256, 3)) ** 6 * 255).astype(np.uint8)
    path = str(tmp_path / "stars.tif")
    cv2.imwrite(path, img8)
    got = load_starmap(path, 96)
    ref = cv2.imread(path)[..., ::-1].astype(np.float32) * (1 / 255)
    want = np.clip(cv2.resize(ref, (96, 48), interpolation=cv2.INTER_CUBIC), 0, 1)
    assert got.dtype == np.float32 and got.shape == (48, 96, 3) and float(np.abs(got - want).max()) <= 5e-5
    again = load_starmap(path, 96)                       # served from <file>.w96.npy
    assert np.array_equal(again, got)
    assert load_starmap(str(tmp_path / "missing.tif"), 96) is None


@pytest.mark.parametrize("big,rps,ds", [(False, 16, 3), (True, 7, 4), (True, 1000, 1), (False, 1, 12)])
def test_streamed_tiff_equals_the_decoded_array_and_writes_the_cache(dl, tmp_path, big, rps, ds):
    """the LDEM's strips go from the file through pinned staging to the device (no decoded copy on the
    host) and the .npy cache is written while the result comes down - same bits as the array path, same file as np.save."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from tiff_util import write_strip_tiff
    rng = np.random.default_rng(rps + ds)
    ldem = rng.integers(-18200, 21600, size=(192, 384), dtype=np.int32).astype(np.int16)
    p = str(tmp_path / "ldem.tif")
    write_strip_tiff(p, ldem, rps, big=big)
    ref, rs_ref = orc.load_elevation(ldem, ds)
    cache = str(tmp_path / "direct.npy")
    e, rs = dl.downscale_elevation_file(p, ds, cache)
    assert np.array_equal(bits(e), bits(ref)) and rs == rs_ref
    back = np.load(cache)
    assert back.dtype == np.float32 and back.flags["C_CONTIGUOUS"] and np.array_equal(bits(back), bits(ref))
    np.save(str(tmp_path / "numpy.npy"), ref)
    assert open(cache, "rb").read() == open(str(tmp_path / "numpy.npy"), "rb").read()      # byte for byte what np.save writes
    # through the reference-shaped entry point: cache + sidecar written, second call served from them
    e2, rs2 = dl.load_elevation_data(p, ds)
    assert np.array_equal(bits(e2), bits(ref)) and rs2 == rs_ref
    if ds > 1:
        assert dl.downscale_cache_available(p, ds)
        meta = json.load(open(f"{p}.ds{ds}.json"))
        assert meta["radius_scale"] == rs and meta["downscale"] == ds
        os.remove(p)                                          # the cache alone serves the next start (data_loader.py:63-86)
        e3, rs3 = dl.load_elevation_data(p, ds)
        assert np.array_equal(bits(e3), bits(ref)) and rs3 == rs_ref


def test_streamed_tiff_rejects_what_it_cannot_stream(dl, tmp_path):
    import cv2
    a = np.arange(96 * 192, dtype=np.uint16).reshape(96, 192)
    lzw = str(tmp_path / "lzw.tif")
    cv2.imwrite(lzw, a)
    with pytest.raises(ValueError):
        dl.downscale_elevation_file(lzw, 2)
    e, rs = dl.load_elevation_data(lzw, 2)                   # ... and the file-level function decodes it on the host instead
    ref, rs_ref = orc.load_elevation(a.view(np.int16), 2)
    assert np.array_equal(bits(e), bits(ref)) and rs == rs_ref
    raw = str(tmp_path / "raw.tif")
    cv2.imwrite(raw, a, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    with pytest.raises(Exception):
        dl.downscale_elevation_file(raw, 5)                  # 96 is not divisible by 5: numpy's reshape ValueError
    assert not os.path.exists(raw + ".ds5.npy")


def test_color_files_the_gpu_reduce_does_not_cover_still_match_opencv(dl, tmp_path):
    """load_color_data (data_loader.py:290-343) for inputs outside the GPU reduce: a size the factor does not divide, a JPEG
    (reduced inside the codec) and a factor OpenCV has no flag for - each must give what the reference's own call gives
    (cv2.imread with the same flag, then _moon_texture), and cache the same BGR image."""
    import cv2
    rng = np.random.default_rng(11)
    col = rng.integers(0, 256, size=(70, 130, 3), dtype=np.int32).astype(np.uint8)
    odd = str(tmp_path / "odd.tif")
    cv2.imwrite(odd, col)
    want = cv2.imread(odd, cv2.IMREAD_REDUCED_COLOR_4)
    assert np.array_equal(dl.load_color_data(odd, 2.2, 4), orc.load_color(want, 2.2, 1))
    assert np.array_equal(np.load(odd + ".ds4.npy"), want)
    smooth = cv2.GaussianBlur(rng.integers(0, 256, size=(64, 128, 3), dtype=np.int32).astype(np.uint8), (0, 0), 3)
    jpg = str(tmp_path / "c.jpg")
    cv2.imwrite(jpg, smooth)
    want = cv2.imread(jpg, cv2.IMREAD_REDUCED_COLOR_2)
    assert np.array_equal(dl.load_color_data(jpg, 2.2, 2), orc.load_color(want, 2.2, 1))
    tif = str(tmp_path / "c.tif")
    cv2.imwrite(tif, col[:64, :128])
    full = cv2.imread(tif, cv2.IMREAD_COLOR)
    assert np.array_equal(dl.load_color_data(tif, 2.2, 3), orc.load_color(full, 2.2, 1))          # no flag for 3: full size ...
    assert np.array_equal(np.load(tif + ".ds3.npy"), full)                                       # ... cached all the same
