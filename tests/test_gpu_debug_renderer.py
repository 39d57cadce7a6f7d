"""The GPU debug (AOV) renderer against the reference's DebugRenderer (libSLR/Renderers/DebugRenderer.cpp:29-217).

Goldens: tests/golden/debug_<scene>.npz = the three BMPs the reference writes (made by tests/golden/make_debug_golden.py
from oracle/_ref/ref_render ... debug). Both renderers take ONE jittered camera sample per pixel, with different random
numbers (the reference's are per-thread xorshift streams), so pixels on a silhouette, a curved surface or a facet
boundary differ by the jitter; flat surfaces must agree to the quantisation step. The goldens therefore carry the noise
floor: floor_<channel> = the fractions of pixels within 6/255 and within 1/255 between TWO reference runs with
different seeds. Tolerance: the GPU image's fractions against the golden are at most 2 % of the pixels below that
floor, and the images' means agree within 1.0.
"""
import os

import numpy as np
import pytest

from slr_b200 import capi
import render_util as ru

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CHANNELS = ("geometric_normal", "shading_normal", "shading_tangent")


def quantise(v):
    """DebugRenderer.cpp:158-160: (uint8_t)clamp((0.5 v + 0.5) * 255, 0, 255)"""
    return np.clip((np.float32(0.5) * v + np.float32(0.5)) * np.float32(255), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("name,w,h", [("spheres", 128, 128), ("materials", 128, 128), ("instanced", 160, 90)])
def test_debug_images_match_reference(name, w, h, tmp_path):
    gold = np.load(os.path.join(GOLD, f"debug_{name}.npz"))
    path = ru.scene_file(name, str(tmp_path), w, h, 1)
    hs = capi.read_scene(path)
    out, st = capi.host_render_debug(hs, w, h, bmp_dir=str(tmp_path))
    assert out.shape == (h, w, capi.DEBUG_FLOATS) and st["paths"] == w * h
    for k, ch in enumerate(CHANNELS):
        got = quantise(out[:, :, 1 + 3 * k:4 + 3 * k])
        want = gold[ch]
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32)).max(-1)
        floor6, floor1 = gold["floor_" + ch]
        assert (diff <= 6).mean() >= floor6 - 0.02, f"{name} {ch}: {(diff <= 6).mean():.4f} of the pixels within 6/255, reference vs reference {floor6:.4f}"
        assert (diff <= 1).mean() >= floor1 - 0.02, f"{name} {ch}: {(diff <= 1).mean():.4f} of the pixels within 1/255, reference vs reference {floor1:.4f}"
        assert abs(float(got.mean()) - float(want.mean())) < 1.0
        # the BMP the renderer wrote holds exactly the quantised vectors (bottom-up BGR rows, the reference's padding)
        raw = open(os.path.join(str(tmp_path), ch + ".bmp"), "rb").read()
        row = 3 * w + w % 4
        px = np.frombuffer(raw, np.uint8, count=row * h, offset=54).reshape(h, row)[:, :3 * w].reshape(h, w, 3)[::-1, :, ::-1]
        assert np.array_equal(px, got)


def test_debug_vectors_are_unit_and_consistent(tmp_path):
    path = ru.scene_file("spheres", str(tmp_path), 96, 96, 1)
    hs = capi.read_scene(path)
    out, _ = capi.host_render_debug(hs, 96, 96)
    hit = out[:, :, 0] == 1.0
    assert hit.mean() > 0.9                      # a closed box: nearly every pixel sees a surface
    for k in range(3):
        n = np.linalg.norm(out[:, :, 1 + 3 * k:4 + 3 * k], axis=-1)
        assert np.allclose(n[hit], 1.0, atol=1e-4)
        assert np.all(n[~hit] == 0.0)            # a miss leaves the reference's zero vectors
    sn, st = out[:, :, 4:7][hit], out[:, :, 7:10][hit]
    # Triangle::getSurfacePoint re-orthogonalises the tangent only when |n . t| >= 0.01 (TriangleMesh.cpp:196-199)
    assert np.abs((sn * st).sum(-1)).max() < 0.0101
    again, _ = capi.host_render_debug(hs, 96, 96)
    assert np.array_equal(again, out)                # counter-based RNG: a re-run draws the same samples
    other, _ = capi.host_render_debug(hs, 96, 96, seed=77)
    assert not np.array_equal(other, out)
