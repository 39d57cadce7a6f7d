"""The reference-side binding, compiled and run: integration/GPUPathTracingRenderer.{h,cpp} is a SLR::Renderer subclass
built INSIDE the reference build (oracle/Makefile `dropin` -> oracle/_ref/slr_gpu: the reference's own HostProgram flow --
its scene interpreter, its scene graph, its SBVH builder and QBVH collapse, its ImageSensor -- with the Renderer swapped
for the GPU one, which flattens the SLR::Scene it receives into include/slrgpu.h's tables and calls the C ABI).
Nothing of libslrhost.so is involved: this is the drop-in of SURVEY.md section 8b proven end to end, and an independent
producer of SlrGpuSceneDesc next to the repo's own host library.
"""
import os
import subprocess

import numpy as np
import pytest

import render_util as ru
from slr_b200 import capi

SLR_GPU = os.path.join(ru.ROOT, "oracle", "_ref", "slr_gpu")
needs_binary = pytest.mark.skipif(not os.access(SLR_GPU, os.X_OK), reason="oracle/_ref/slr_gpu not built on this machine")


def run_slr_gpu(scene_path, dump=True):
    d = os.path.dirname(os.path.abspath(scene_path))
    out = os.path.join(d, "dropin_sensor.bin")
    p = subprocess.run([SLR_GPU, os.path.basename(scene_path)] + ([out] if dump else []), capture_output=True, text=True, cwd=d, timeout=1200)
    return p, out


def read_sensor(path):
    with open(path, "rb") as f:
        w, h, c = np.frombuffer(f.read(12), np.uint32)
        return np.frombuffer(f.read(), np.float32).reshape(h, w, c).copy()


@needs_binary
@pytest.mark.parametrize("name", ["spheres", "materials", "instanced", "lamps", "cutout", "ibl", "textured", "motion", "nested"])
def test_exported_scene_passes_the_library_validation(name, tmp_path):
    """CPU: the exporter's tables are accepted by slrgpu_scene_create's validation pass -- without a device the call gets as
    far as SLRGPU_ERR_NO_DEVICE (validation runs first), with one the render succeeds."""
    path = ru.scene_file(name, str(tmp_path), 24, 24, 2)
    p, _ = run_slr_gpu(path)
    if capi.gpu.slrgpu_device_count() == 0:
        assert p.returncode == 1 and "no CUDA device" in p.stderr, p.stderr[-500:]
    else:
        assert p.returncode == 0, p.stderr[-500:]


@needs_binary
def test_unsupported_content_fails_loudly(tmp_path):
    """What the exporter cannot express in the tables of include/slrgpu.h is an error, not a different image: here an
    animated node inside an animated node (two moving transforms in one chain; the repo's own host library refuses it too)."""
    path = ru.scene_file("motion", str(tmp_path), 24, 24, 2)
    text = open(path).read()
    placed = "addChild(root, flyer);"
    assert placed in text, "the motion scene's layout changed: adapt this test"
    text = text.replace(placed, """outer = createNode();
addChild(outer, flyer);
setTransform(outer, AnimatedTransform(translate(0.0, 0.0, 0.0), translate(0.1, 0.0, 0.0), 0.0, 1.0));
addChild(root, outer);""")
    with open(path, "w") as f:
        f.write(text)
    p, _ = run_slr_gpu(path)
    assert p.returncode == 1 and "chain of two animated transforms" in p.stderr, p.stderr[-500:]


@needs_binary
@pytest.mark.gpu
@pytest.mark.parametrize("name,size,spp", [("spheres", 96, 256), ("materials", 96, 256), ("instanced", 96, 128), ("cutout", 96, 256),
                                           ("ibl", 96, 256), ("textured", 96, 256), ("motion", 96, 256), ("nested", 96, 256)])
def test_dropin_renders_like_the_reference(name, size, spp, tmp_path):
    """The sensor the GPU renderer leaves behind (read through the reference's own ImageSensor::pixel) against the
    reference's PathTracingRenderer on the same file: the image-parity bars of tests/test_gpu_render.py. And against the
    repo's own host library on the same file: same RNG keys, same trees -> the same image up to fp32 summation order."""
    assert capi.gpu.slrgpu_device_count() > 0
    path = ru.scene_file(name, str(tmp_path), size, size, spp)
    p, out = run_slr_gpu(path)
    assert p.returncode == 0, p.stderr[-800:]
    sensor = read_sensor(out)
    assert sensor.shape == (size, size, 16) and np.isfinite(sensor).all()
    gpu = capi.accum_to_rgb(sensor, 1.0 / spp)
    ref1 = capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=1509761209)[0], 1.0 / spp)
    ref2 = capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=20240229)[0], 1.0 / spp)
    (ref1, gpu, ref2), d1 = ru.sanitize_reference(ref1, gpu, ref2)
    (ref2, gpu, ref1), d2 = ru.sanitize_reference(ref2, gpu, ref1)
    floor = ru.rel_rmse(ref2, ref1, trim=0.005)
    got = ru.rel_rmse(gpu, ref1, trim=0.005)
    assert got <= 1.25 * floor, f"relRMSE {got:.4f} vs noise floor {floor:.4f}"
    clip = float(np.percentile(ref1, 99.8))
    ratio = np.minimum(gpu, clip).reshape(-1, 3).mean(0) / np.minimum(ref1, clip).reshape(-1, 3).mean(0)
    assert np.all(np.abs(ratio - 1.0) < 0.01), f"image mean ratio {ratio}"
    # the same file through the repo's own host library (seed = the file's default, like slr_gpu)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    mine, _ = capi.host_render(hs, size, size, spp, device=0)
    mine = capi.accum_to_rgb(mine, 1.0 / spp)
    if name == "nested":
        # the two producers expand the nesting on their own (different object order in the rebuilt aggregates, so different
        # trees and light lists): the same estimator, not the same samples -- held to the reference's noise floor instead
        got2 = ru.rel_rmse(mine, ref1, trim=0.005)
        assert got2 <= 1.25 * floor, f"host library: relRMSE {got2:.4f} vs noise floor {floor:.4f}"
    else:
        np.testing.assert_allclose(ru.block_means(capi.accum_to_rgb(sensor, 1.0 / spp), 8), ru.block_means(mine, 8), rtol=2e-3, atol=1e-6)
    # the progressive BMPs the renderer wrote through the reference's own ImageSensor::saveImage
    assert os.path.exists(os.path.join(os.path.dirname(path), "000.bmp"))


@needs_binary
@pytest.mark.gpu
def test_dropin_bidirectional_renderer(tmp_path):
    """`slr_gpu scene.txt out.bin bpt`: a scene file that selects "BPT" (the reference's unchanged Cornell_Box_Spheres.txt) gets
    SLR::GPUBidirectionalPathTracingRenderer through the reference's own program; the sensor it leaves meets the reference
    BPT's two-seed noise floor."""
    assert capi.gpu.slrgpu_device_count() > 0
    size, spp = 96, 32
    path = ru.reference_scene_file("Cornell_Box_Spheres.txt", str(tmp_path), size, size, spp, method="BPT")
    if path is None or not ru.have_ref_render():
        pytest.skip("the reference's scene files / ref_render did not travel to this machine")
    d = os.path.dirname(os.path.abspath(path))
    out = os.path.join(d, "dropin_sensor.bin")
    p = subprocess.run([SLR_GPU, os.path.basename(path), out, "bpt"], capture_output=True, text=True, cwd=d, timeout=1200)
    assert p.returncode == 0, p.stderr[-800:]
    gpu = capi.accum_to_rgb(read_sensor(out), 1.0 / spp)
    ref1 = capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=1509761209, bpt=True)[0], 1.0 / spp)
    ref2 = capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=20240229, bpt=True)[0], 1.0 / spp)
    (ref1, gpu, ref2), _ = ru.sanitize_reference(ref1, gpu, ref2)
    floor = ru.rel_rmse(ref2, ref1, trim=0.005)
    got = ru.rel_rmse(gpu, ref1, trim=0.005)
    assert got <= 1.25 * floor, f"relRMSE {got:.4f} vs noise floor {floor:.4f}"


def _load_table_dump(path):
    """The file SLRGPU_DROPIN_DUMP writes: u32 magic, nodes, leaf records, instances, lights, top-level lights, then the four tables."""
    import ctypes as C
    raw = open(path, "rb").read()
    magic, nn, nl, ni, nlt, ntop = np.frombuffer(raw[:24], np.uint32)
    assert magic == 0x44524F50
    off = 24
    tables = []
    for cls, n in ((capi.BvhNode, int(nn)), (capi.LeafRecord, int(nl)), (capi.Instance, int(ni)), (capi.Light, int(nlt))):
        size = C.sizeof(cls) * n
        arr = (cls * max(n, 1))()
        C.memmove(arr, raw[off:off + size], size)
        off += size
        tables.append(arr)
    assert off == len(raw)
    d = capi.SceneDesc()
    d.struct_size = C.sizeof(capi.SceneDesc)
    d.bvh_nodes, d.num_bvh_nodes = C.cast(tables[0], C.POINTER(capi.BvhNode)), int(nn)
    d.leaf_records, d.num_leaf_records = C.cast(tables[1], C.POINTER(capi.LeafRecord)), int(nl)
    d.instances, d.num_instances = C.cast(tables[2], C.POINTER(capi.Instance)), int(ni)
    d.lights, d.num_lights, d.num_top_lights = C.cast(tables[3], C.POINTER(capi.Light)), int(nlt), int(ntop)
    return d, tables


@needs_binary
@pytest.mark.parametrize("name", ["instanced", "lamps", "nested", "motion"])
def test_exported_tables_give_the_reference_hits(name, tmp_path, monkeypatch):
    """CPU: the geometry tables the exporter hands to slrgpu_scene_create -- for `nested` the scene it expanded to one level,
    for `motion` the begin key frames -- walked by the CPU restatement of the traversal, against the reference's own
    Scene::intersect on the same file and rays (oracle/_ref/ref_probe, shutter closed at time 0): the same rays hit, at
    the same distance up to the rounding of the composed transforms."""
    import ctypes as C
    import oracle_util as ou
    if not ru.have_ref_probe():
        pytest.skip("oracle/_ref/ref_probe not built")
    path = ru.scene_file(name, str(tmp_path), 24, 24, 2)
    dump = str(tmp_path / "tables.bin")
    monkeypatch.setenv("SLRGPU_DROPIN_DUMP", dump)
    run_slr_gpu(path)
    monkeypatch.delenv("SLRGPU_DROPIN_DUMP")
    desc, keep = _load_table_dump(dump)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)          # only for the scene's bounding sphere
    center = [hs.desc.world_center[i] for i in range(3)]
    radius = hs.desc.world_radius
    probes = ru.make_probes(center, radius, 6000, 5)
    want = ru.run_ref_probe(path, probes)
    n = probes.shape[0]
    comps = [np.ascontiguousarray(probes[:, k], np.float32) for k in range(6)] + [np.zeros(n, np.float32), np.full(n, np.inf, np.float32)]
    rb = capi.RayBatch(*[c.ctypes.data_as(capi.PF) for c in comps])
    prim, inst, t = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.float32)
    u, v = np.empty(n, np.float32), np.empty(n, np.float32)
    hb = capi.HitBatch(prim.ctypes.data_as(capi.PU32), inst.ctypes.data_as(capi.PU32), t.ctypes.data_as(capi.PF),
                       u.ctypes.data_as(capi.PF), v.ctypes.data_as(capi.PF), None, None)
    tn, tl = C.c_uint64(0), C.c_uint64(0)
    rc = ou.restate_lib().slr_restate_intersect(C.byref(desc), C.byref(rb), n, C.byref(hb), C.byref(tn), C.byref(tl))
    assert rc == 0
    hit_ref, hit_got = want[:, 0] == 1, prim != 0xFFFFFFFF
    assert float(np.mean(hit_ref != hit_got)) <= 0.001
    both = hit_ref & hit_got
    excess = np.abs(t[both] - want[both, 1]) - (2e-5 * np.abs(want[both, 1]) + 1e-6 * radius)
    assert both.sum() >= 500 and float(excess.max()) <= 0.0, float(excess.max())
    # light selection: every emitting triangle has importance 1 and an aggregate's importance is the sum of its entries'
    # (SurfaceObject.cpp:232-252), so whatever the nesting, each emitting triangle placement must be chosen with the same
    # probability: the product of the pmfs along its chain
    probs = []
    for k in range(desc.num_top_lights):
        l = desc.lights[k]
        if l.object >> 31:
            ins = desc.instances[l.object & 0x7FFFFFFF]
            assert ins.light_index == k and ins.num_lights > 0
            for j in range(ins.num_lights):
                lj = desc.lights[ins.light_base + j]
                assert not (lj.object >> 31), "one level on the device: a nested light list holds triangles only"
                probs.append(l.pmf * lj.pmf)
        else:
            probs.append(l.pmf)
    assert probs and abs(sum(probs) - 1.0) < 1e-5 and np.allclose(probs, 1.0 / len(probs), rtol=1e-5)
    if name == "nested":
        assert desc.num_instances > 0 and len(probs) == 8          # 2 + 3 x 2 emitting triangle placements, as tests/test_nested_instancing.py
