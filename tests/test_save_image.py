"""ImageSensor::saveImage is byte work (strata -> XYZ -> sRGB, 1 - exp(-Y) tone map, gamma, 24-bit bottom-up BMP:
libSLR/Core/ImageSensor.cpp:138-186, Helper/bmp_exporter.cpp:14-53) and must be bit-exact: the reference's own sensor
sums, fed to the host library's saveImage, have to give the bytes of the NNN.bmp the reference wrote -- header and
every pixel byte; the row padding (width % 4 bytes per row) is uninitialised heap in the reference and is excluded.
Also the export cadence of PathTracingRenderer.cpp:63-65,83-94: images after 1, 2, 4, ... samples, at most 16.
"""
import glob
import os

import numpy as np
import pytest

import render_util as ru
from slr_b200 import capi


def _assert_same_bmp(mine, ref, w, h):
    hm, pm = ru.bmp_pixels(mine, w, h)
    hr, pr = ru.bmp_pixels(ref, w, h)
    assert np.array_equal(hm, hr), "BMP header differs"
    diff = int((pm != pr).sum())
    assert diff == 0, f"{diff} of {pm.size} pixel bytes differ"


def test_save_image_bytes_match_golden(tmp_path):
    g = np.load(os.path.join(ru.GOLDEN, "save_image.npz"))
    accum, spp = g["accum"], int(g["spp"])
    h, w, _ = accum.shape
    out = str(tmp_path / "mine.bmp")
    capi.save_bmp(out, accum, float(g["brightness"]) / spp, float(g["sensitivity"]))
    names = sorted(k for k in g.files if k.startswith("bmp_"))
    assert names == ["bmp_000", "bmp_001", "bmp_002", "bmp_003"]          # 1, 2, 4, 8 samples
    with open(out, "rb") as f:
        _assert_same_bmp(f.read(), g[names[-1]].tobytes(), w, h)


@pytest.mark.parametrize("name,w,h,spp,brightness", [("diffuse", 33, 20, 4, 1.0), ("ibl", 64, 48, 2, 4.0)])
def test_save_image_bytes_match_live_reference(name, w, h, spp, brightness, tmp_path):
    if not ru.have_ref_render():
        pytest.skip("oracle/_ref/ref_render not built")
    path = ru.scene_file(name, str(tmp_path), w, h, spp)
    accum, timing, bmps = ru.run_ref_render_with_bmps(path, spp, w, h, seed=11)
    assert len(bmps) == int(np.log2(spp)) + 1
    out = str(tmp_path / "mine.bmp")
    capi.save_bmp(out, accum, brightness / spp, timing["sensitivity"])
    with open(out, "rb") as f:
        _assert_same_bmp(f.read(), bmps[-1], w, h)


@pytest.mark.gpu
def test_gpu_renderer_export_cadence_and_bytes(tmp_path):
    """GPUPathTracingRenderer writes NNN.bmp after 1, 2, 4, ... samples like the reference, and the last one is the
    tone-mapped sensor it returns."""
    assert capi.gpu.slrgpu_device_count() > 0, "needs a CUDA device"
    w, h, spp = 50, 37, 12            # 12 is not a power of two: images at 1, 2, 4, 8 only
    path = ru.scene_file("diffuse", str(tmp_path), w, h, spp)
    hs = capi.read_scene(path)
    bdir = str(tmp_path / "bmps")
    os.makedirs(bdir)
    accum8, _ = capi.host_render(hs, w, h, 8)
    accum, st = capi.host_render(hs, w, h, spp, bmp_dir=bdir)
    assert st["paths"] == w * h * spp
    files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(bdir, "*.bmp")))
    assert files == ["000.bmp", "001.bmp", "002.bmp", "003.bmp"]
    # 003.bmp = the sensor after 8 samples (same RNG keys: the first 8 samples of the 12-sample call)
    mine = str(tmp_path / "mine.bmp")
    cam = hs.desc.camera        # PerspectiveCamera's default sensitivity: 1 / lens area (PerspectiveCamera.cpp:15-24)
    sens = float(cam.sensitivity) if cam.sensitivity > 0 else float(np.float32(1.0 / (np.pi * cam.lens_radius * cam.lens_radius)))
    capi.save_bmp(mine, accum8, 1.0 / 8, sens)
    with open(mine, "rb") as f, open(os.path.join(bdir, "003.bmp"), "rb") as g:
        a, b = ru.bmp_pixels(f.read(), w, h)[1], ru.bmp_pixels(g.read(), w, h)[1]
    # summation order of the progressive segments differs from one 8-sample call by fp32 rounding: at most one code value
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1 and (a != b).mean() < 0.01
