"""
The CPU oracle against (i) fixtures produced by the unmodified reference
(oracle/make_golden.py) and (ii) the live reference when /root/reference exists.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import downscale_oracle as orc
from oracle import ref_stub

DS_SMALL = (1, 2, 3, 4, 5, 6, 9, 12)


@pytest.fixture(scope="module")
def elev_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "elevation_small.npz"))


@pytest.mark.parametrize("name", ["synth", "uniform"])
@pytest.mark.parametrize("ds", DS_SMALL)
@pytest.mark.parametrize("explicit", [False, True])
def test_elevation_oracle_matches_reference_fixture(elev_golden, name, ds, explicit):
    src = elev_golden[f"{name}_src"]
    e, rs = orc.load_elevation(src, ds, explicit=explicit)
    ref = elev_golden[f"{name}_ds{ds}"]
    assert e.dtype == np.float32 and e.shape == ref.shape
    assert np.array_equal(e.view(np.uint32), ref.view(np.uint32))      # bit-exact
    assert rs == float(elev_golden[f"{name}_ds{ds}_rs"])
    assert e.max() == np.float32(1.0)


def _big_map(elev_golden):
    big = np.tile(elev_golden["synth_src"], (8, 8))
    return (big.astype(np.int32) + (np.arange(big.shape[1])[None, :] % 97) * 3
            - (np.arange(big.shape[0])[:, None] % 89) * 5).astype(np.int16)


@pytest.mark.parametrize("explicit", [False, True])
def test_elevation_oracle_matches_reference_digests(elev_golden, golden_dir, explicit):
    with open(os.path.join(golden_dir, "elevation_digests.json")) as f:
        dig = json.load(f)
    big = _big_map(elev_golden)
    for ds, case in dig["cases"].items():
        e, rs = orc.load_elevation(big, int(ds), explicit=explicit)
        assert list(e.shape) == case["shape"]
        assert hashlib.sha256(e.tobytes()).hexdigest() == case["sha256"], f"ds={ds}"
        assert rs == case["radius_scale"]


def test_elevation_not_divisible_raises(elev_golden):
    src = elev_golden["synth_src"]
    with pytest.raises(ValueError):
        orc.load_elevation(src, 7)           # 360 % 7 != 0, like reshape in data_loader.py:225
    with pytest.raises(ValueError):
        orc.load_elevation(src, 7, explicit=True)


def test_color_oracle_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "color_small.npz"))
    src = g["src_bgr"]
    for k in (1, 2, 4, 8):
        for gamma in (2.2, 1.0):
            tex = orc.load_color(src, gamma, k)
            assert np.array_equal(tex, g[f"k{k}_g{gamma}"]), (k, gamma)
    for gamma in (0.5, 1.0, 1.8, 2.2, 5.0):
        assert np.array_equal(orc.albedo_lut(gamma), g[f"lut_g{gamma}"])
    lut = orc.albedo_lut(2.2)
    assert (lut[0], lut[128], lut[255]) == (7, 75, 227)      # SURVEY.md §8 A4 [probed]


@pytest.mark.skipif(not ref_stub.reference_available(), reason="reference tree not on this box")
@pytest.mark.parametrize("ds", [2, 3, 4, 7, 16, 64])
def test_elevation_oracle_vs_live_reference(tmp_path, ds):
    dl = ref_stub.import_reference("data_loader")
    rng = np.random.default_rng(100 + ds)
    src = rng.integers(-18200, 21600, size=(ds * 37, ds * 53), dtype=np.int32).astype(np.int16)
    p = tmp_path / "ldem.tif"
    p.write_bytes(b"stub")
    ref_stub.set_read_image(str(p), src.view(np.uint16).copy())
    ref, rs_ref = dl.load_elevation_data(str(p), ds)
    for explicit in (False, True):
        e, rs = orc.load_elevation(src, ds, explicit=explicit)
        assert np.array_equal(e.view(np.uint32), np.asarray(ref).view(np.uint32))
        assert rs == rs_ref


@pytest.mark.skipif(not ref_stub.reference_available(), reason="reference tree not on this box")
def test_color_oracle_vs_live_reference(tmp_path):
    import cv2
    dl = ref_stub.import_reference("data_loader")
    rng = np.random.default_rng(9)
    src = rng.integers(0, 256, size=(96, 208, 3), dtype=np.int32).astype(np.uint8)
    p = str(tmp_path / "c.tif")
    cv2.imwrite(p, src)
    for k in (1, 2, 4, 8):
        ref = dl.load_color_data(p, 2.2, k)
        assert np.array_equal(orc.load_color(src, 2.2, k), ref)
        for ext in (".npy", ".json"):
            c = f"{p}.ds{k}{ext}"
            if os.path.exists(c):
                os.remove(c)
