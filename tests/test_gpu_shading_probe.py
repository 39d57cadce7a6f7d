"""GPU function-level parity of the shade stage: slrgpu_probe_shading runs the material kernels' device functions
(closest hit -> surface point incl. normal maps and instance transforms -> material -> BSDF with its textures and
spectra -> BSDF::sample / evaluate / evaluatePDF, emittance) on probe rays with given random numbers; the
reference's own classes do the same in oracle/_ref/ref_probe (Intersection::getSurfacePoint,
SurfacePoint::createBSDF, BSDF::sample/evaluate/evaluatePDF: the calls of PathTracingRenderer.cpp:147-210).

Bars (libm vs CUDA math differ by a few ulp, so floating-point values are tolerance-based; decisions are exact):
  * hit / miss / environment status, hasNonDelta, isEmitting: identical for every probe;
  * t: bit-equal; position, shading normal and tangent: within 5e-6 absolute;
  * sampled direction TYPE (which lobe / reflect vs transmit was chosen): identical for >= 99.8 % of the probes
    (a uniform within rounding distance of a selection threshold may flip);
  * sampled fs (16 wavelengths), direction and pdf; evaluated fs and pdf: within 1e-4 relative for >= 99.5 % of
    the probes and within 2e-3 for >= 99.9 %; emittance within 1e-5 relative.
Scenes: the five image-parity scenes -- together they contain every BSDF model, the checker / Voronoi spectrum,
float and normal textures, image (environment) textures, regular / irregular / up-sampled spectra, instancing.
"""
import os

import numpy as np
import pytest

import render_util as ru
from slr_b200 import capi

pytestmark = pytest.mark.gpu
SCENES = ["diffuse", "spheres", "materials", "ibl", "instanced", "cutout", "textured", "motion", "nested"]


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert capi.gpu.slrgpu_device_count() > 0, "these tests need a CUDA device"


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("probe_scenes"))


def check(got, want, composed_transforms=False):
    """composed_transforms (`nested`): instancing nested in instancing reaches the device as ONE level with the transforms of
    a chain multiplied on the host (host/scene.h PlacedSubtree), where the reference transforms the ray once per level: the
    hit is the same, its distance and frame agree to rounding instead of bit for bit."""
    r = ru.compare_probes(got, want, rel=1e-4)
    loose = ru.compare_probes(got, want, rel=2e-3)
    assert r["nondelta_mismatch"] == 0.0 and r["emitting_mismatch"] == 0.0, r
    if composed_transforms:
        assert r["status_mismatch"] <= 0.001 and r["t_worst_rel"] <= 1e-3 and r["frame_worst_abs"] <= 5e-5, r      # distances: tests/test_nested_instancing.py
    else:
        assert r["status_mismatch"] == 0.0 and r["t_worst_rel"] == 0.0, r
        assert r["frame_worst_abs"] <= 5e-6, r
    assert r["sample_type_mismatch"] <= 0.002, r
    assert r["sample_value_mismatch"] <= 0.005 and r["eval_mismatch"] <= 0.005, r
    assert loose["sample_value_mismatch"] <= 0.001 and loose["eval_mismatch"] <= 0.001, loose
    assert r["emittance_worst_rel"] <= 1e-5, r
    return r


def gpu_scene(name, workdir):
    path = ru.scene_file(name, workdir, 64, 64, 1)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    return path, hs, capi.GpuScene(hs)


@pytest.mark.parametrize("name", SCENES)
def test_probe_matches_golden(name, workdir):
    g = np.load(os.path.join(ru.GOLDEN, f"probe_{name}.npz"))
    _, hs, gs = gpu_scene(name, workdir)
    got = capi.probe_shading(gs, g["probes"])
    r = check(got, g["reference"], composed_transforms=name == "nested")
    assert r["hits"] >= 500


@pytest.mark.parametrize("name", SCENES)
def test_probe_matches_live_reference(name, workdir):
    if not ru.have_ref_probe():
        pytest.skip("oracle/_ref/ref_probe not built")
    path, hs, gs = gpu_scene(name, workdir)
    center = [hs.desc.world_center[i] for i in range(3)]
    probes = ru.make_probes(center, hs.desc.world_radius, 20000, 7)
    check(capi.probe_shading(gs, probes), ru.run_ref_probe(path, probes), composed_transforms=name == "nested")


def test_probe_argument_errors(workdir):
    _, hs, gs = gpu_scene("diffuse", workdir)
    with pytest.raises(capi.SlrError):
        capi._gpu_check(capi.gpu.slrgpu_probe_shading(gs.handle, None, 4, None), "slrgpu_probe_shading")
