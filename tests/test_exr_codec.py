"""The host library's EXR reader / writer (slr_b200/host/assets/exr.h; the reference uses OpenEXR's RgbaInputFile,
libSLRSceneGraph/Helper/image_loader.cpp:38-62) against an INDEPENDENT codec: OpenCV's bundled OpenEXR. Files written here
must read back exactly there, and files written there (uncompressed half RGBA scanlines, the subset the reader supports)
must read back exactly here -- the reader is no longer checked only against its own writer."""
import os

import numpy as np
import pytest

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
cv2 = pytest.importorskip("cv2")

from slr_b200 import capi  # noqa: E402


def _pattern(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.random((h, w, 4)).astype(np.float32)
    img[..., :3] *= np.float32(8.0)                          # HDR range
    img[0, 0] = (0.0, 1.0, 65504.0, 1.0)                    # zero, one, the largest half
    return img


@pytest.mark.parametrize("size", [(17, 23), (64, 128), (1, 5)])
def test_our_writer_is_read_exactly_by_openexr(size, tmp_path):
    h, w = size
    rgba = _pattern(h, w, h)
    path = str(tmp_path / "ours.exr")
    capi.write_exr(path, rgba)
    back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if back is None:
        pytest.skip("this OpenCV build has no OpenEXR codec")
    assert back.shape == (h, w, 4)
    want = rgba.astype(np.float16).astype(np.float32)
    assert np.array_equal(back[..., [2, 1, 0, 3]], want)     # OpenCV delivers BGRA


@pytest.mark.parametrize("size", [(19, 31), (48, 96)])
def test_openexr_written_file_is_read_exactly_by_our_reader(size, tmp_path):
    h, w = size
    rgba = _pattern(h, w, w)
    path = str(tmp_path / "theirs.exr")
    ok = cv2.imwrite(path, np.ascontiguousarray(rgba[..., [2, 1, 0, 3]]),
                     [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_NO])
    if not ok:
        pytest.skip("this OpenCV build cannot write EXR")
    got = capi.read_exr(path)
    assert got.shape == (h, w, 4)
    assert np.array_equal(got, rgba.astype(np.float16).astype(np.float32))


def test_compressed_exr_fails_loudly(tmp_path):
    path = str(tmp_path / "zip.exr")
    if not cv2.imwrite(path, _pattern(8, 8, 1)[..., [2, 1, 0, 3]], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF,
                                                                   cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_ZIP]):
        pytest.skip("this OpenCV build cannot write EXR")
    with pytest.raises(capi.SlrError, match="compress"):
        capi.read_exr(path)
