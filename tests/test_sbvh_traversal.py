"""SBVH::intersect (Accelerator/SBVH.h:417-442) with BoundingBox3D::intersect
(Core/geometry.h:112-126) -- the accelerator the reference's SHIPPED build traverses (SurfaceObject.cpp:226-230). It
visits leaves in a different order than the QBVH, so rays that meet two primitives at bit-equal distances can report the
other primitive (the goldens record how many). slrgpu_intersect_batch_sbvh reproduces the SBVH's answers, bit for bit, on
the golden batches -- including exactly those tie rays; the CPU restatement of the same algorithm is pinned first.
"""
import numpy as np
import pytest

import oracle_util as ou
from slr_b200 import capi


def golden(name):
    return np.load(f"{ou.GOLDEN}/intersect_{name}.npz")


def assert_same(got, g):
    assert np.array_equal(got["prim"], g["prim_sbvh"]), f"{(got['prim'] != g['prim_sbvh']).sum()} hit-id mismatches"
    assert np.array_equal(got["inst"], g["inst_sbvh"])
    hit = g["prim_sbvh"] != 0xFFFFFFFF
    assert np.array_equal(got["t"].view(np.uint32)[hit], g["t_sbvh_bits"][hit])
    assert np.array_equal(got["u"].view(np.uint32)[hit], g["u_sbvh_bits"][hit])
    assert np.array_equal(got["v"].view(np.uint32)[hit], g["v_sbvh_bits"][hit])
    assert np.all(np.isinf(got["t"][~hit]))


@pytest.mark.parametrize("name", list(ou.CASES))
def test_restatement_matches_the_reference_sbvh(name):
    meshes, placements, rays = ou.CASES[name]()
    hs = ou.build_host_scene(meshes, placements, with_sbvh=True)
    assert hs.desc.num_sbvh_nodes > 0 and hs.desc.num_sbvh_leaf_records >= hs.desc.num_triangles
    r = ou.restate_intersect_sbvh(hs, rays)
    assert r["overflow"] == 0
    assert_same(r, golden(name))


@pytest.mark.parametrize("name", list(ou.CASES))
def test_golden_records_where_the_two_accelerators_disagree(name):
    """The two reference accelerators disagree exactly where the goldens say so: rays through shared edges / vertices,
    where the visiting order decides between two primitives at (nearly) the same distance, and axis-parallel rays in a box
    face, where 0 x inf = NaN falls differently through the SSE min / max of the QBVH and the comparisons of
    BoundingBox3D::intersect -- there one accelerator can even miss what the other hits."""
    g = golden(name)
    differ = g["prim"] != g["prim_sbvh"]
    assert int(differ.sum()) == int(g["qbvh_vs_sbvh_mismatches"])
    # (no closeness bar on those rays: in the `objects` batch a ray lying in a cube face is hit at t = 1 by the QBVH and at
    # t = 3, the far wall, by the SBVH -- each entry point must give ITS accelerator's answer, which is what is pinned)
    same = ~differ & (g["prim"] != 0xFFFFFFFF)
    assert np.array_equal(g["t_bits"][same], g["t_sbvh_bits"][same])


def test_scene_without_sbvh_tables_is_refused_for_the_sbvh_query():
    meshes, placements, rays = ou.CASES["objects"]()
    hs = ou.build_host_scene(meshes, placements)
    assert hs.desc.num_sbvh_nodes == 0
    with pytest.raises(AssertionError):
        ou.restate_intersect_sbvh(hs, rays)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(ou.CASES))
def test_gpu_sbvh_hits_bit_exact(name):
    assert capi.gpu.slrgpu_device_count() > 0, "needs a CUDA device"
    meshes, placements, rays = ou.CASES[name]()
    hs = ou.build_host_scene(meshes, placements, with_sbvh=True)
    gs = capi.GpuScene(hs)
    g = golden(name)
    got = gs.intersect_sbvh(rays)
    assert_same(got, g)
    # the QBVH entry point of the same scene still gives the QBVH's answers
    q = gs.intersect(rays)
    assert np.array_equal(q["prim"], g["prim"])


@pytest.mark.gpu
def test_gpu_sbvh_against_restatement_at_size():
    from slr_b200 import synth
    pos, idx = synth.heightfield(200)
    rays = synth.concat_rays(synth.random_rays(200_000, pos.min(0), pos.max(0), seed=5), ou.special_rays(pos, idx, 512))
    hs = ou.build_host_scene([(pos, idx)], [(0, 0, None)], with_sbvh=True)
    gs = capi.GpuScene(hs)
    got, want = gs.intersect_sbvh(rays), ou.restate_intersect_sbvh(hs, rays)
    assert np.array_equal(got["prim"], want["prim"])
    hit = want["prim"] != 0xFFFFFFFF
    assert np.array_equal(got["t"].view(np.uint32)[hit], want["t"].view(np.uint32)[hit])


@pytest.mark.gpu
def test_gpu_sbvh_query_needs_the_tables():
    meshes, placements, rays = ou.CASES["objects"]()
    gs = capi.GpuScene(ou.build_host_scene(meshes, placements))
    with pytest.raises(capi.SlrError, match="sbvh_nodes"):
        gs.intersect_sbvh(rays)
