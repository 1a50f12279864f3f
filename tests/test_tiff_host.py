"""The TIFF directory parser behind the streaming LDEM reader (mrtx_tiff_info: host work only, no GPU needed)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tiff_util import write_strip_tiff  # noqa: E402


@pytest.fixture(scope="module")
def dl():
    from moonrtx_b200 import data_loader
    return data_loader


@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("rps", [1, 7, 16, 96, 1000])
def test_strip_tiffs_are_streamable_and_opencv_reads_the_same_pixels(dl, tmp_path, big, rps):
    import cv2
    rng = np.random.default_rng(rps)
    a = rng.integers(-18200, 21600, size=(96, 160), dtype=np.int32).astype(np.int16)
    p = str(tmp_path / f"t{rps}{big}.tif")
    write_strip_tiff(p, a, rps, big=big)
    assert dl.tiff_info(p) == {"width": 160, "height": 96, "bits": 16, "streamable": True}
    back = cv2.imread(p, cv2.IMREAD_UNCHANGED)               # the writer makes files a real TIFF reader accepts
    assert back is not None and np.array_equal(back.view(np.int16), a)


def test_files_that_cannot_be_streamed_say_so(dl, tmp_path):
    import cv2
    a = np.arange(96 * 160, dtype=np.uint16).reshape(96, 160)
    lzw = str(tmp_path / "lzw.tif")
    cv2.imwrite(lzw, a)                                     # OpenCV compresses (LZW) by default
    i = dl.tiff_info(lzw)
    assert i["width"] == 160 and i["height"] == 96 and i["bits"] == 16 and not i["streamable"]
    raw = str(tmp_path / "raw.tif")
    cv2.imwrite(raw, a, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    assert dl.tiff_info(raw)["streamable"]
    be = str(tmp_path / "be.tif")
    write_strip_tiff(be, a, 16, byteorder=">")
    assert dl.tiff_info(be) == {"width": 160, "height": 96, "bits": 16, "streamable": False}     # big-endian samples
    rgb = str(tmp_path / "rgb.tif")
    cv2.imwrite(rgb, np.zeros((8, 8, 3), np.uint8), [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    assert not dl.tiff_info(rgb)["streamable"]
    junk = str(tmp_path / "junk.tif")
    open(junk, "wb").write(b"not a tiff at all, just bytes")
    assert dl.tiff_info(junk) is None and dl.tiff_info(str(tmp_path / "missing.tif")) is None
