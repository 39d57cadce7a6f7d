"""slrgpu_scene_create range-checks every index of the caller's tables on the host before anything is uploaded and
refuses what the device code cannot represent (CPU tests: validation runs before the device is touched, so a valid
scene gets as far as SLRGPU_ERR_NO_DEVICE here and an invalid one is rejected with the error that names the entry).
"""
import copy
import ctypes as C

import numpy as np
import pytest

import render_util as ru
from slr_b200 import capi

OK, INVALID, NO_DEVICE, UNSUPPORTED = 0, -1, -2, -5


def create(desc):
    out = C.c_void_p()
    rc = capi.gpu.slrgpu_scene_create(C.byref(desc), 0, C.byref(out))
    msg = capi.gpu.slrgpu_last_error().decode()
    if rc == OK:
        capi.gpu.slrgpu_scene_destroy(out)
    return rc, msg


def expected_ok():
    return OK if capi.gpu.slrgpu_device_count() > 0 else NO_DEVICE


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("validation"))
    path = ru.scene_file("materials", d, 32, 32, 1)
    with capi.stdout_to_stderr():
        return capi.read_scene(path)


def clone(desc):
    c = capi.SceneDesc()
    C.memmove(C.byref(c), C.byref(desc), C.sizeof(desc))
    return c


def patched(desc, field, count_field, elem_type, mutate):
    """a copy of `desc` whose table `field` is a mutated private copy; returns (desc, keepalive)"""
    n = getattr(desc, count_field)
    arr = (elem_type * n)()
    C.memmove(arr, getattr(desc, field), C.sizeof(elem_type) * n)
    mutate(arr)
    c = clone(desc)
    setattr(c, field, C.cast(arr, C.POINTER(elem_type)))
    return c, arr


@pytest.mark.parametrize("name", ["diffuse", "spheres", "materials", "ibl", "instanced", "scatter", "lamps", "cutout", "textured", "motion"])
def test_every_test_scene_passes_validation(name, tmp_path):
    path = ru.scene_file(name, str(tmp_path), 32, 32, 1)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    rc, msg = create(hs.desc)
    assert rc == expected_ok(), msg


def test_geometry_only_scene_passes_validation():
    from slr_b200 import synth
    b = capi.SceneBuilder()
    pos, idx = synth.cube()
    b.place_mesh(b.add_mesh(pos, idx))
    hs = b.finish()           # owns the buffers the description points into
    rc, msg = create(hs.desc)
    assert rc == expected_ok(), msg


def test_out_of_range_indices_are_rejected(scene):
    d = scene.desc

    def bad_child(a): a[0].child[0] = 0x07FFFFF0            # inner child far beyond the node count
    def bad_leaf_child(a): a[0].child[0] = 0x80000000 | (15 << 27) | (d.num_leaf_records - 1)
    def bad_vertex(a): a[3].v[1] = d.num_vertices
    def bad_material(a): a[0].material = d.num_materials + 7
    def bad_alpha(a): a[0].alpha_map = d.num_textures
    def bad_tex(a): a[0].tex[0] = d.num_textures
    def bad_spectrum_id(a):
        for t in a:
            if t.kind == 0:
                t.i0 = d.num_spectra
                return
    def bad_spectrum_data(a):
        for sp in a:
            if sp.kind in (0, 1):
                sp.data_offset = d.num_spectrum_floats
                return
        a[0].kind = 99
    def bad_light(a): a[0].object = d.num_triangles

    cases = [("bvh_nodes", "num_bvh_nodes", capi.BvhNode, bad_child, "node 0 child 0"),
             ("bvh_nodes", "num_bvh_nodes", capi.BvhNode, bad_leaf_child, "leaf records"),
             ("triangles", "num_triangles", capi.Triangle, bad_vertex, "vertex index"),
             ("triangles", "num_triangles", capi.Triangle, bad_material, "material"),
             ("triangles", "num_triangles", capi.Triangle, bad_alpha, "alpha map"),
             ("materials", "num_materials", capi.Material, bad_tex, "material 0"),
             ("textures", "num_textures", capi.Texture, bad_spectrum_id, "texture"),
             ("spectra", "num_spectra", capi.Spectrum, bad_spectrum_data, "spectrum"),
             ("lights", "num_lights", capi.Light, bad_light, "light 0")]
    for field, count, typ, mutate, needle in cases:
        c, keep = patched(d, field, count, typ, mutate)
        rc, msg = create(c)
        assert rc == INVALID, (field, needle, rc, msg)
        assert needle in msg, (needle, msg)
    # the untouched description is still fine
    assert create(d)[0] == expected_ok()


def test_leaf_record_checks(scene):
    d = scene.desc

    def bad_instance(a):
        bits = np.array([0x80000000 | 12345], np.uint32).view(np.float32)[0]
        a[0].a[3] = bits
    c, keep = patched(d, "leaf_records", "num_leaf_records", capi.LeafRecord, bad_instance)
    rc, msg = create(c)
    assert rc == INVALID and "instance" in msg

    def bad_triangle(a):
        a[0].a[3] = np.array([d.num_triangles + 1], np.uint32).view(np.float32)[0]
    c, keep = patched(d, "leaf_records", "num_leaf_records", capi.LeafRecord, bad_triangle)
    rc, msg = create(c)
    assert rc == INVALID and "triangle" in msg


def test_unsupported_material_trees_are_refused(scene):
    """More than four leaf lobes (MultiBSDF.h:17 holds four) is SLRGPU_ERR_UNSUPPORTED, not a silently thinner BSDF."""
    d = scene.desc
    n = d.num_materials
    # append a chain of four `sum` nodes over five diffuse leaves and point triangle 0 at its root
    extra = 9
    mats = (capi.Material * (n + extra))()
    C.memmove(mats, d.materials, C.sizeof(capi.Material) * n)
    diffuse = None
    for i in range(n):
        if mats[i].kind == 0:
            diffuse = i
            break
    assert diffuse is not None
    for k in range(5):
        C.memmove(C.byref(mats[n + k]), C.byref(mats[diffuse]), C.sizeof(capi.Material))
    inv = 0xFFFFFFFF
    # sums: n+5 = (leaf0, leaf1), n+6 = (n+5, leaf2), n+7 = (n+6, leaf3), n+8 = (n+7, leaf4)
    prev = n
    for k in range(4):
        m = mats[n + 5 + k]
        m.kind = 8
        m.tex[0] = m.tex[1] = m.tex[2] = m.tex[3] = inv
        m.sub[0] = prev
        m.sub[1] = n + 1 + k
        prev = n + 5 + k
    c = clone(d)
    c.materials = C.cast(mats, C.POINTER(capi.Material))
    c.num_materials = n + extra
    tris = (capi.Triangle * d.num_triangles)()
    C.memmove(tris, d.triangles, C.sizeof(capi.Triangle) * d.num_triangles)
    tris[0].material = n + 8
    c.triangles = C.cast(tris, C.POINTER(capi.Triangle))
    rc, msg = create(c)
    assert rc == UNSUPPORTED and "5 leaf lobes" in msg, (rc, msg)
    # four lobes are fine
    tris[0].material = n + 7
    assert create(c)[0] == expected_ok()
