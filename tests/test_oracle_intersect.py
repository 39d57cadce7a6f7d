"""CPU: pins the restatement (oracle/restate) and the host SBVH/QBVH builder against the golden
outputs of the compiled reference, and against the live reference when oracle/_ref is present."""
import hashlib

import numpy as np
import pytest

import oracle_util as ou


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_golden(name):
    return np.load(f"{ou.GOLDEN}/intersect_{name}.npz")


@pytest.mark.parametrize("name", list(ou.CASES))
def test_inputs_regenerate_identically(name):
    """The deterministic generators must reproduce the ray batch the goldens were made from."""
    _, _, rays = ou.CASES[name]()
    g = load_golden(name)
    d = sha(np.stack([rays[k] for k in ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")]))
    assert d == str(g["ray_digest"])


@pytest.mark.parametrize("name", list(ou.CASES))
def test_host_trees_match_reference(name):
    """SBVH (spatial splits included) -> QBVH built by libslrhost is bit-identical to the reference's:
    same nodes, same leaf reference order, same SAH costs."""
    meshes, placements, _ = ou.CASES[name]()
    hs = ou.build_host_scene(meshes, placements)
    g = load_golden(name)
    trees = ou.split_trees(hs)
    assert len(trees) == int(g["num_trees"])
    for i, t in enumerate(trees):
        assert tuple(g[f"tree{i}_shape"]) == (t["nodes"].shape[0], t["refs"].shape[0])
        assert sha(t["nodes"]) == str(g[f"tree{i}_nodes_sha"]), f"tree {i}: QBVH nodes differ from the reference"
        assert sha(t["refs"]) == str(g[f"tree{i}_refs_sha"]), f"tree {i}: leaf reference order differs"
        costs = g[f"tree{i}_costs"]
        assert np.float32(t["sbvh_cost"]) == costs[0] and np.float32(t["qbvh_cost"]) == costs[1]


@pytest.mark.parametrize("name", list(ou.CASES))
def test_restatement_matches_reference_hits(name):
    """The CPU restatement reproduces the reference's QBVH hits bit for bit (ids, t, u, v), including
    the tie cases where QBVH and SBVH themselves disagree."""
    meshes, placements, rays = ou.CASES[name]()
    hs = ou.build_host_scene(meshes, placements)
    g = load_golden(name)
    r = ou.restate_intersect(hs, rays)
    assert r["overflow"] == 0
    assert np.array_equal(r["prim"], g["prim"])
    assert np.array_equal(r["inst"], g["inst"])
    hit = g["prim"] != 0xFFFFFFFF
    assert np.array_equal(r["t"].view(np.uint32)[hit], g["t_bits"][hit])
    assert np.array_equal(r["u"].view(np.uint32)[hit], g["u_bits"][hit])
    assert np.array_equal(r["v"].view(np.uint32)[hit], g["v_bits"][hit])
    assert np.all(np.isinf(r["t"][~hit]))
    assert r["total_nodes"] == int(r["nodes"].sum()) and r["total_tris"] == int(r["tris"].sum())


def test_golden_contains_ties():
    """The special-ray batch really exercises the visiting order: the reference's own SBVH and QBVH
    disagree on some of these rays."""
    assert int(load_golden("heightfield")["qbvh_vs_sbvh_mismatches"]) > 0


@pytest.mark.skipif(not ou.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_live_reference_larger_scene():
    """With the reference binary available: a larger scene (spatial splits, 80k triangles), live."""
    import slr_b200.synth as synth
    pos, idx = synth.heightfield(200)
    rays = synth.concat_rays(synth.random_rays(20000, pos.min(0), pos.max(0), seed=1),
                             synth.aimed_rays(20000, pos.min(0), pos.max(0), seed=2))
    hits, trees, info = ou.run_ref_intersect([(pos, idx)], [(0, 0, None)], rays)
    hs = ou.build_host_scene([(pos, idx)], [(0, 0, None)])
    mine = ou.split_trees(hs)
    assert np.array_equal(mine[0]["nodes"], ou.normalise_ref_tree(trees[0]))
    assert np.array_equal(mine[0]["refs"], trees[0]["refs"])
    r = ou.restate_intersect(hs, rays)
    assert np.array_equal(r["prim"], hits["prim"])
    assert np.array_equal(r["t"].view(np.uint32), hits["t"].view(np.uint32))


def test_empty_and_degenerate_inputs():
    import slr_b200.synth as synth
    from slr_b200 import capi
    # a single triangle: root is a leaf, wrapped into a one-lane node (QBVH.h:261-277)
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    idx = np.array([[0, 2, 1]], np.uint32)
    hs = ou.build_host_scene([(pos, idx)], [(0, 0, None)])
    assert hs.desc.num_bvh_nodes == 1 and hs.desc.num_leaf_records == 1
    rays = {"ox": np.array([0.25, 5.0], np.float32), "oy": np.array([1.0, 1.0], np.float32), "oz": np.array([0.25, 5.0], np.float32),
            "dx": np.zeros(2, np.float32), "dy": -np.ones(2, np.float32), "dz": np.zeros(2, np.float32),
            "tmin": np.zeros(2, np.float32), "tmax": np.full(2, np.inf, np.float32)}
    r = ou.restate_intersect(hs, rays)
    assert r["prim"][0] == 0 and r["prim"][1] == 0xFFFFFFFF and r["t"][0] == 1.0
    # zero rays
    empty = {k: np.zeros(0, np.float32) for k in rays}
    r = ou.restate_intersect(hs, empty)
    assert r["prim"].shape == (0,)
    # an empty scene is rejected by the host
    b = capi.SceneBuilder()
    with pytest.raises(capi.SlrError):
        b.finish()


def test_parallel_tree_build_equals_serial_build(monkeypatch):
    """The host SBVH builder hands subtrees above 32 k fragments to other threads and splices the results in pre-order
    (host/bvh.cpp). Whatever the thread count, the tree must be the same tree: SLRHOST_BUILD_THREADS=1 (plain
    recursion) against the default, on a mesh large enough for several levels of parallel nodes, spatial splits included."""
    import slr_b200.synth as synth
    pos, idx = synth.heightfield(300)            # 180 000 triangles
    monkeypatch.setenv("SLRHOST_BUILD_THREADS", "1")
    serial = ou.split_trees(ou.build_host_scene([(pos, idx)], [(0, 0, None)]))
    monkeypatch.setenv("SLRHOST_BUILD_THREADS", "8")
    threaded = ou.split_trees(ou.build_host_scene([(pos, idx)], [(0, 0, None)]))
    monkeypatch.delenv("SLRHOST_BUILD_THREADS")
    default = ou.split_trees(ou.build_host_scene([(pos, idx)], [(0, 0, None)]))
    for other in (threaded, default):
        assert len(other) == len(serial) == 1
        assert np.array_equal(other[0]["nodes"], serial[0]["nodes"])
        assert np.array_equal(other[0]["refs"], serial[0]["refs"])
        assert other[0]["sbvh_cost"] == serial[0]["sbvh_cost"] and other[0]["qbvh_cost"] == serial[0]["qbvh_cost"]
