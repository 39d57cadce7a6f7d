"""Test-side access to the oracle: the CPU restatement (oracle/lib/liboracle_restate.so) and, when
present, the compiled reference drivers under oracle/_ref/. Test infrastructure only."""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

from slr_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RESTATE_SO = os.path.join(ROOT, "oracle", "lib", "liboracle_restate.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_restate = None


def restate_lib():
    global _restate
    if _restate is None:
        if not os.path.exists(RESTATE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "restate"])
        _restate = C.CDLL(RESTATE_SO)
        _restate.slr_restate_intersect.restype = C.c_int
        _restate.slr_restate_intersect.argtypes = [C.POINTER(capi.SceneDesc), C.POINTER(capi.RayBatch), C.c_uint64,
                                                   C.POINTER(capi.HitBatch), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return _restate


def restate_intersect(host_scene, rays, counters=True):
    """Closest hits of the CPU restatement on a flattened host scene."""
    lib = restate_lib()
    comps = [np.ascontiguousarray(rays[k], np.float32) for k in ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")]
    n = comps[0].shape[0]
    rb = capi.RayBatch(*[c.ctypes.data_as(capi.PF) for c in comps])
    out = {"prim": np.empty(n, np.uint32), "inst": np.empty(n, np.uint32), "t": np.empty(n, np.float32),
           "u": np.empty(n, np.float32), "v": np.empty(n, np.float32),
           "nodes": np.empty(n, np.uint32), "tris": np.empty(n, np.uint32)}
    hb = capi.HitBatch(out["prim"].ctypes.data_as(capi.PU32), out["inst"].ctypes.data_as(capi.PU32),
                       out["t"].ctypes.data_as(capi.PF), out["u"].ctypes.data_as(capi.PF), out["v"].ctypes.data_as(capi.PF),
                       out["nodes"].ctypes.data_as(capi.PU32), out["tris"].ctypes.data_as(capi.PU32))
    tn, tl = C.c_uint64(0), C.c_uint64(0)
    rc = lib.slr_restate_intersect(C.byref(host_scene.desc), C.byref(rb), n, C.byref(hb), C.byref(tn), C.byref(tl))
    out["overflow"] = rc
    out["total_nodes"] = tn.value
    out["total_tris"] = tl.value
    return out


def restate_intersect_sbvh(host_scene, rays):
    """Closest hits of the CPU restatement of SBVH::intersect (the scene must carry the SBVH tables)."""
    lib = restate_lib()
    lib.slr_restate_intersect_sbvh.restype = C.c_int
    lib.slr_restate_intersect_sbvh.argtypes = [C.POINTER(capi.SceneDesc), C.POINTER(capi.RayBatch), C.c_uint64, C.POINTER(capi.HitBatch)]
    comps = [np.ascontiguousarray(rays[k], np.float32) for k in ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")]
    n = comps[0].shape[0]
    rb = capi.RayBatch(*[c.ctypes.data_as(capi.PF) for c in comps])
    out = {"prim": np.empty(n, np.uint32), "inst": np.empty(n, np.uint32), "t": np.empty(n, np.float32),
           "u": np.empty(n, np.float32), "v": np.empty(n, np.float32)}
    hb = capi.HitBatch(out["prim"].ctypes.data_as(capi.PU32), out["inst"].ctypes.data_as(capi.PU32),
                       out["t"].ctypes.data_as(capi.PF), out["u"].ctypes.data_as(capi.PF), out["v"].ctypes.data_as(capi.PF), None, None)
    rc = lib.slr_restate_intersect_sbvh(C.byref(host_scene.desc), C.byref(rb), n, C.byref(hb))
    assert rc >= 0, "the scene carries no SBVH tables (capi.set_option('export_sbvh', 1) before building it)"
    out["overflow"] = rc
    return out


def have_ref(binary="ref_intersect"):
    return os.path.exists(os.path.join(REF_DIR, binary))


def run_ref_intersect(meshes, placements, rays, want_trees=True, threads=None):
    """Runs the compiled reference on a geometry spec + ray batch; returns (hits, trees, timing json)."""
    with tempfile.TemporaryDirectory() as d:
        spec, rf, hf, tf = (os.path.join(d, n) for n in ("spec.bin", "rays.bin", "hits.bin", "trees.bin"))
        synth.write_geom_spec(spec, meshes, placements)
        synth.write_rays(rf, rays)
        cmd = [os.path.join(REF_DIR, "ref_intersect"), spec, rf, hf, tf if want_trees else "-"]
        if threads is not None:
            cmd.append(str(threads))
        r = subprocess.run(cmd, capture_output=True, text=True, check=True)
        info = json.loads(r.stderr.strip().splitlines()[-1])
        hits = synth.read_hits(hf)
        trees = synth.read_trees(tf) if want_trees else None
    return hits, trees, info


# ------------------------------------------------------------------------------------------------
# the deterministic geometry cases shared by the golden generator and the tests
# ------------------------------------------------------------------------------------------------
def special_rays(pos, idx, count=512):
    """Rays that stress ties and degenerate arithmetic: straight down onto grid vertices/edges,
    axis-aligned with zero direction components, aimed exactly at mesh vertices and edge midpoints."""
    pos = np.asarray(pos, np.float32)
    lo, hi = pos.min(0), pos.max(0)
    top = np.float32(hi[1] + 1.0)
    sel = np.linspace(0, pos.shape[0] - 1, count).astype(np.int64)
    v = pos[sel]
    n = v.shape[0]
    down = {"ox": v[:, 0].copy(), "oy": np.full(n, top, np.float32), "oz": v[:, 2].copy(),
            "dx": np.zeros(n, np.float32), "dy": np.full(n, -1, np.float32), "dz": np.zeros(n, np.float32),
            "tmin": np.zeros(n, np.float32), "tmax": np.full(n, np.inf, np.float32)}
    negz = {k: a.copy() for k, a in down.items()}
    negz["dx"] = np.full(n, -0.0, np.float32)          # -0.0 components select the other slab side
    negz["dz"] = np.full(n, -0.0, np.float32)
    tri = idx[np.linspace(0, idx.shape[0] - 1, count).astype(np.int64)]
    mid = (pos[tri[:, 0]].astype(np.float64) + pos[tri[:, 1]].astype(np.float64)) * 0.5
    org = np.array([(lo[0] + hi[0]) * 0.5, top, (lo[2] + hi[2]) * 0.5], np.float64)
    d = mid - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    edge = {"ox": np.full(n, org[0], np.float32), "oy": np.full(n, org[1], np.float32), "oz": np.full(n, org[2], np.float32),
            "dx": d[:, 0].astype(np.float32), "dy": d[:, 1].astype(np.float32), "dz": d[:, 2].astype(np.float32),
            "tmin": np.zeros(n, np.float32), "tmax": np.full(n, np.inf, np.float32)}
    dv = v.astype(np.float64) - org
    dv /= np.linalg.norm(dv, axis=1, keepdims=True)
    vert = {k: a.copy() for k, a in edge.items()}
    vert["dx"], vert["dy"], vert["dz"] = (dv[:, i].astype(np.float32) for i in range(3))
    along = {"ox": np.full(n, lo[0] - 0.5, np.float32), "oy": v[:, 1].copy(), "oz": v[:, 2].copy(),
             "dx": np.ones(n, np.float32), "dy": np.zeros(n, np.float32), "dz": np.zeros(n, np.float32),
             "tmin": np.full(n, 1e-4, np.float32), "tmax": np.full(n, 1.25, np.float32)}   # finite tmax, may end inside
    return synth.concat_rays(down, negz, edge, vert, along)


def case_heightfield(n=64, rays_each=3000):
    pos, idx = synth.heightfield(n)
    rays = synth.concat_rays(synth.random_rays(rays_each, pos.min(0), pos.max(0), seed=12345),
                             synth.aimed_rays(rays_each, pos.min(0), pos.max(0), seed=777),
                             special_rays(pos, idx, 256))
    return [(pos, idx)], [(0, 0, None)], rays


def case_objects(rays_each=3000):
    c_pos, c_idx = synth.cube()
    s_pos, s_idx, _, _, _ = synth.uv_sphere(32, 16)
    meshes = [(c_pos, c_idx), (s_pos, s_idx)]
    placements = [(0, 0, synth.translate(0.0, 0.0, 0.0) @ synth.scale(3.0, 2.0, 3.0)),
                  (1, 0, synth.translate(-0.7, -1.0, -1.05) @ synth.scale(0.5) @ synth.translate(0, 1, 0))]
    lo, hi = np.array([-3, -2, -3], np.float32), np.array([3, 2, 3], np.float32)
    rays = synth.concat_rays(synth.random_rays(rays_each, lo, hi, seed=99, inflate=0.0),
                             synth.aimed_rays(rays_each, lo * 0.5, hi * 0.5, seed=5, height=0.2),
                             special_rays(s_pos * 0.5 + np.array([-0.7, -0.5, -1.05], np.float32), s_idx, 128))
    return meshes, placements, rays


def case_instanced(rays_each=3000):
    """Two-level scene: a ground quad baked at the top level plus 9 instances of one heightfield."""
    g_pos = np.array([[-3, 0, -3], [3, 0, -3], [3, 0, 3], [-3, 0, 3]], np.float32)
    g_idx = np.array([[0, 2, 1], [0, 3, 2]], np.uint32)
    h_pos, h_idx = synth.heightfield(24, amplitude=2.0)
    meshes = [(g_pos, g_idx), (h_pos, h_idx)]
    placements = [(0, 0, None)]
    k = 0
    for gx in range(3):
        for gz in range(3):
            m = synth.translate(-2.2 + 1.7 * gx, 0.3 + 0.05 * k, -2.2 + 1.7 * gz) @ synth.rotate_y(0.37 * k) @ synth.scale(1.1 + 0.07 * k, 0.9, 1.0 + 0.03 * k)
            placements.append((1, 1, m))
            k += 1
    lo, hi = np.array([-3, 0, -3], np.float32), np.array([3, 1.2, 3], np.float32)
    rays = synth.concat_rays(synth.random_rays(rays_each, lo, hi, seed=4242),
                             synth.aimed_rays(rays_each, lo, hi, seed=31337, height=1.0))
    return meshes, placements, rays


CASES = {"heightfield": case_heightfield, "objects": case_objects, "instanced": case_instanced}


def build_host_scene(meshes, placements, with_sbvh=False):
    if with_sbvh:
        capi.set_option("export_sbvh", 1)
    try:
        return _build_host_scene(meshes, placements)
    finally:
        if with_sbvh:
            capi.set_option("export_sbvh", 0)


def _build_host_scene(meshes, placements):
    b = capi.SceneBuilder()
    ids = [b.add_mesh(p, i) for p, i in meshes]
    for mesh, mode, m in placements:
        (b.place_mesh if mode == 0 else b.instance_mesh)(ids[mesh], m)
    hs = b.finish()
    b.close()
    return hs


def split_trees(host_scene):
    """Un-concatenates the host scene's node / leaf arrays into per-aggregate local-index trees,
    in the order ref_intersect dumps them (top level first, nested aggregates in order of discovery)."""
    nodes = host_scene.nodes_array()
    leaves = host_scene.leaves_array()
    stats = host_scene.stats()
    out = []
    for i, st in enumerate(stats):
        n0, n1 = st["nodeBase"], st["nodeBase"] + st["qbvhNodes"]
        l0 = st["leafBase"]
        l1 = stats[i + 1]["leafBase"] if i + 1 < len(stats) else leaves.shape[0]
        nd = nodes[n0:n1].copy()
        ch = nd[:, 24:28]
        valid = ch != 0xFFFFFFFF
        leaf = (ch >> 31) == 1
        idx = ch & 0x07FFFFFF
        idx = np.where(leaf, idx - l0, idx - n0)
        nd[:, 24:28] = np.where(valid, (ch & 0xF8000000) | (idx & 0x07FFFFFF), ch)
        nd[:, 28] &= 0x00FFFFFF
        nd[:, 29:] = 0
        out.append({"nodes": nd, "refs": leaves[l0:l1, 3].copy(), "sbvh_cost": st["sbvhCost"], "qbvh_cost": st["qbvhCost"]})
    return out


def normalise_ref_tree(tree):
    nd = tree["nodes"].copy()
    nd[:, 28] &= 0x00FFFFFF
    nd[:, 29:] = 0
    return nd
