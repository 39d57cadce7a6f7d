"""The light-selection tables the GPU path samples from, against the reference's own (CPU, bit-exact).

Next event estimation picks a light with SurfaceObjectAggregate's RegularConstantDiscrete1D over the emitters' importances
(libSLR/Core/SurfaceObject.cpp:232-243, 279-299; distributions.cpp:97-119, compensated sums) inside the world sphere
Scene::build computes (SurfaceObject.cpp:396-406). The host library restates both when it flattens a scene
(slr_b200/host/scene.cpp, shading.cpp). oracle/_ref/ref_render <scene> ... lights prints the reference's values as raw
float bits; tests/golden/lights_<scene>.txt holds them (made by `python tests/test_light_tables.py --make-golden`), and the
host's SlrGpuSceneDesc must carry the same bits: world centre and radius, number of top-level lights, their importance
integral, every pmf and both cdf ends."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_util as ou  # noqa: E402
import render_util as ru  # noqa: E402
from slr_b200 import capi  # noqa: E402

SCENES = ["diffuse", "spheres", "materials", "instanced", "ibl", "scatter", "lamps"]     # lamps: emitters inside instances


def golden_path(name):
    return os.path.join(ru.GOLDEN, f"lights_{name}.txt")


def reference_dump(path):
    p = subprocess.run([os.path.join(ou.REF_DIR, "ref_render"), os.path.basename(path), "out.bin", "1", "8", "8", "0", "0", "lights"],
                       cwd=os.path.dirname(path), capture_output=True, text=True)
    return "\n".join(l for l in p.stdout.splitlines() if l.startswith(("world ", "camera ", "lights ", "pmf ", "env ", "envpdf "))) + "\n"


def parse(text):
    world, n, integral, rows = None, None, None, []
    for l in text.splitlines():
        t = l.split()
        if not t:
            continue
        if t[0] == "world":
            world = [int(x, 16) for x in t[1:5]]
        elif t[0] == "lights":
            n, integral = int(t[1]), int(t[3], 16)
        elif t[0] == "pmf":
            rows.append((int(t[2], 16), int(t[4], 16), int(t[5], 16)))
    return world, n, integral, rows


def bits(x):
    return int(np.float32(x).view(np.uint32))


def check(name, text, tmp):
    world, n, integral, rows = parse(text)
    assert world is not None and n == len(rows)
    hs = capi.read_scene(ru.scene_file(name, tmp, 8, 8, 1))
    d = hs.desc
    assert [bits(d.world_center[i]) for i in range(3)] + [bits(d.world_radius)] == world
    assert d.num_top_lights == n
    if n:
        assert bits(d.top_light_importance) == integral
    for i, (pmf, lo, hi) in enumerate(rows):
        L = d.lights[i]
        assert (bits(L.pmf), bits(L.cdf_lo), bits(L.cdf_hi)) == (pmf, lo, hi), f"light {i}"
    # the camera: local-to-world matrix (column-major) and its inverse, aspect, fovY, lens radius, image / object plane distances, sensitivity
    cam = [l.split()[1:] for l in text.splitlines() if l.startswith("camera ")]
    assert len(cam) == 1
    want = [int(x, 16) for x in cam[0]]
    c = d.camera
    got = [bits(c.mat[i]) for i in range(16)] + [bits(c.mat_inv[i]) for i in range(16)] + \
          [bits(c.aspect), bits(c.fov_y), bits(c.lens_radius), bits(c.img_plane_dist), bits(c.obj_plane_dist)]
    assert got == want[:37], "camera differs from the reference's"
    assert np.isclose(np.float32(c.sensitivity), np.array(want[37], np.uint32).view(np.float32), rtol=1e-6) or c.sensitivity == 0
    # the environment's importance map (IBLEmission::createIBLImportanceMap -> RegularConstantContinuous2D): every pdf, cdf and
    # row integral of the reference's map enters an FNV-1a hash in a fixed order; the host's arrays must hash to the same value
    env = [l.split() for l in text.splitlines() if l.startswith("env ")]
    e = d.environment
    assert bool(e.present) == bool(env)
    if env:
        w, h = int(env[0][1]), int(env[0][2])
        assert (e.map_width, e.map_height) == (w, h)
        f32 = lambda ptr, n: np.ctypeslib.as_array(ptr, shape=(n,)).view(np.uint32)
        row_pdf, row_cdf = f32(e.row_pdf, w * h).reshape(h, w), f32(e.row_cdf, (w + 1) * h).reshape(h, w + 1)
        row_int, mpdf, mcdf = f32(e.row_integral, h), f32(e.marginal_pdf, h), f32(e.marginal_cdf, h + 1)
        stream = np.concatenate([np.concatenate([row_pdf[y], row_cdf[y], row_int[y:y + 1]]) for y in range(h)] + [mpdf, mcdf]).astype(np.uint32)
        hsh = 1469598103934665603
        for b in stream.view(np.uint8).tolist():
            hsh = ((hsh ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        assert bits(e.marginal_integral) == int(env[0][6], 16)
        assert hsh == int(env[0][8], 16), "importance map differs from the reference's"
        for l in text.splitlines():
            if l.startswith("envpdf "):
                t = l.split()
                x, y = int(t[1]), int(t[2])
                assert (int(row_pdf[y, x]), int(row_cdf[y, x]), int(mpdf[y])) == (int(t[3], 16), int(t[4], 16), int(t[5], 16))


@pytest.mark.parametrize("name", SCENES)
def test_light_tables_match_golden(name, tmp_path):
    check(name, open(golden_path(name)).read(), str(tmp_path))


@pytest.mark.skipif(not ou.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("name", ["spheres", "instanced"])
def test_light_tables_match_live_reference(name, tmp_path):
    path = ru.scene_file(name, str(tmp_path / "ref"), 8, 8, 1)
    check(name, reference_dump(path), str(tmp_path))


if __name__ == "__main__" and "--make-golden" in sys.argv:
    import tempfile
    for name in SCENES:
        with tempfile.TemporaryDirectory() as d:
            text = reference_dump(ru.scene_file(name, d, 8, 8, 1))
        with open(golden_path(name), "w") as f:
            f.write(text)
        print(name, len(text.splitlines()), "lines")
