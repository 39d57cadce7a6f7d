"""Instancing nested in instancing (CPU): the reference walks a chain of TransformedSurfaceObjects recursively
(libSLR/Core/SurfaceObject.cpp:307-336, libSLRSceneGraph/nodes.cpp:174-184); the host library expands the chain when it
flattens a scene -- the triangles a referenced subtree owns become one aggregate, every instance found inside it is placed
again under (outer transform x its own) -- so the device traversal stays at one level for any depth (host/scene.h
PlacedSubtree). Checked without a GPU: the flattened tables, the light-selection probabilities of the expansion, and the
closest hits of the CPU restatement on the expanded tables against the reference's own recursion on the same scene file
(golden: tests/golden/probe_nested.npz, made by ref_probe; live when ref_probe is present). GPU side: `nested` in the probe
and image tests."""
import os

import numpy as np
import pytest

import oracle_util as ou
import render_util as ru
from slr_b200 import capi


@pytest.fixture(scope="module")
def nested(tmp_path_factory):
    path = ru.scene_file("nested", str(tmp_path_factory.mktemp("nested")), 64, 64, 1)
    with capi.stdout_to_stderr():
        return path, capi.read_scene(path)


def test_expansion_tables(nested):
    _, hs = nested
    d = hs.desc
    # cluster = 3 ball references + own triangles; referenced 3 times: 3 templates + 3 x (1 + 3) placed instances
    assert d.num_instances == 15
    # no leaf record of a nested BVH is an instance record (one level on the device): only the top-level tree holds them
    leaves = hs.leaves_array()
    stats = hs.stats()
    top_leaves = leaves[stats[0]["leafBase"]:stats[1]["leafBase"], 3]
    nested_leaves = leaves[stats[1]["leafBase"]:, 3]
    assert int((top_leaves >> 31).sum()) == 12 and int((nested_leaves >> 31).sum()) == 0
    # lights: the ceiling quad's two triangles + the three placed triangle aggregates (each holds the panel's two triangles)
    assert d.num_top_lights == 5


def test_light_selection_probabilities_equal_the_chain_products(nested):
    """Reference: P(panel triangle) = pmf_top(cluster instance) x pmf_cluster(triangle); an aggregate's importance is the sum of
    its lights' importances (1 per emitting triangle), so with 2 + 3 x 2 emitting triangles every one is chosen with 1 / 8."""
    _, hs = nested
    lights = hs.lights_array() if hasattr(hs, "lights_array") else None
    if lights is None:
        pytest.skip("light table accessor not exported")
    top = lights[:hs.desc.num_top_lights]
    pmf = top["pmf"]
    inst = (top["object"] >> 31) == 1
    assert np.allclose(pmf[~inst], 1.0 / 8.0) and np.allclose(pmf[inst], 2.0 / 8.0)
    assert np.allclose(lights[hs.desc.num_top_lights:]["pmf"], 0.5)


def _rays_from_probes(probes):
    n = probes.shape[0]
    return {"ox": probes[:, 0], "oy": probes[:, 1], "oz": probes[:, 2], "dx": probes[:, 3], "dy": probes[:, 4], "dz": probes[:, 5],
            "tmin": np.zeros(n, np.float32), "tmax": np.full(n, np.inf, np.float32)}


def _check_hits(hs, probes, want):
    got = ou.restate_intersect(hs, _rays_from_probes(probes), counters=False)
    hit_ref = want[:, 0] == 1
    hit_got = got["prim"] != 0xFFFFFFFF
    # the same rays hit; composed transforms move a distance by rounding only
    assert float(np.mean(hit_ref != hit_got)) <= 0.001
    both = hit_ref & hit_got
    # 2e-5 of the distance + 1e-6 of the scene's radius (origins a fraction of a millimetre from a surface see the rounding
    # of the origin's own transform, not of the distance)
    excess = np.abs(got["t"][both] - want[both, 1]) - (2e-5 * np.abs(want[both, 1]) + 1e-6 * hs.desc.world_radius)
    assert both.sum() >= 500 and float(excess.max()) <= 0.0, float(excess.max())


def test_closest_hits_match_the_reference_recursion_golden(nested):
    _, hs = nested
    g = np.load(os.path.join(ru.GOLDEN, "probe_nested.npz"))
    _check_hits(hs, g["probes"], g["reference"])


def test_closest_hits_match_the_live_reference(nested):
    if not ru.have_ref_probe():
        pytest.skip("oracle/_ref/ref_probe not built")
    path, hs = nested
    center = [hs.desc.world_center[i] for i in range(3)]
    probes = ru.make_probes(center, hs.desc.world_radius, 6000, 3)
    _check_hits(hs, probes, ru.run_ref_probe(path, probes))
