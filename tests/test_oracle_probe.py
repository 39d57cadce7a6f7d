"""CPU: the committed probe / render goldens are what the compiled reference produces (where oracle/_ref is
built, i.e. in the container that has /root/reference), and are well-formed everywhere."""
import os

import numpy as np
import pytest

import render_util as ru
from slr_b200 import capi

SCENES = ["diffuse", "spheres", "materials", "ibl", "instanced"]


@pytest.mark.parametrize("name", SCENES)
def test_probe_golden_is_well_formed(name):
    g = np.load(os.path.join(ru.GOLDEN, f"probe_{name}.npz"))
    assert g["probes"].shape == (2048, 14) and g["reference"].shape == (2048, 64)
    ref = g["reference"]
    assert set(np.unique(ref[:, 0]).tolist()) <= {0.0, 1.0, 2.0}
    hit = ref[:, 0] == 1
    assert hit.sum() >= 500 and np.isfinite(ref[hit][:, :12]).all()
    # shading normals of the reference are unit vectors
    assert np.allclose(np.linalg.norm(ref[hit][:, 5:8], axis=1), 1.0, atol=1e-4)


@pytest.mark.parametrize("name", SCENES)
def test_render_golden_is_well_formed(name):
    g = np.load(os.path.join(ru.GOLDEN, f"render_{name}.npz"))
    assert g["block_mean"].shape == (8, 8, 3) and np.isfinite(g["block_mean"]).all() and np.isfinite(g["block_sigma"]).all()
    assert (g["block_mean"] > 0).all() and int(g["ref_spp"]) == 16384


@pytest.mark.parametrize("name", ["spheres", "materials"])
def test_probe_golden_regenerates_from_the_reference(name, tmp_path):
    """ref_probe is single-threaded and deterministic: the committed golden must come back bit for bit."""
    if not ru.have_ref_probe():
        pytest.skip("oracle/_ref/ref_probe not built (no /root/reference on this machine)")
    g = np.load(os.path.join(ru.GOLDEN, f"probe_{name}.npz"))
    path = ru.scene_file(name, str(tmp_path), 64, 64, 1)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    center = [hs.desc.world_center[i] for i in range(3)]
    probes = ru.make_probes(center, hs.desc.world_radius, 2048, 20261018)
    assert np.array_equal(probes, g["probes"]), "probe generator changed: regenerate the goldens"
    again = ru.run_ref_probe(path, probes)
    assert np.array_equal(again.view(np.uint32), g["reference"].view(np.uint32))
