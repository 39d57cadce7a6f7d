"""CPU: the reference's own TestScenes/*.txt are read UNCHANGED by the host scene language (north_star: "TestScenes/*.txt
render unchanged"). Runs only where /root/reference is mounted (the files are copied to a temp directory at test time and
never into this repository); the models / environment maps they load are replaced by synthetic assets because the
reference does not ship them (README.md:69-72). All seven shipped scene files are covered."""
import os
import re
import shutil

import pytest

from slr_b200 import capi, scenes, synth

SRC = "/root/reference/TestScenes"
# expected (width, height, samples, brightness) as written in the files
EXPECT = {
    "Cornell_Box_Spheres.txt": (1024, 768, 16384, 1.0),
    "Cornell_Box_Boxes.txt": (1024, 1024, 16384, 1.0),
    "Cornell_Box_ColorChecker.txt": (1024, 1024, 16384, 4.0),
    "Cornell_Box_ColorChecker_OverrideMaterial.txt": (1024, 1024, 512, 4.0),
    "IBL_Test.txt": (1024, 1024, 16384, 4.0),
    # environment-lit, 60 x 60 grass instances scattered by scanXZFromYPlus (a host ray cast while the file is read)
    "RTC3.txt": (1920, 1080, 16384, 2.0),
    "RTC3_pika.txt": (1920, 1080, 16384, 2.0),
}
ENV_LIT = {"IBL_Test.txt", "RTC3.txt", "RTC3_pika.txt"}


@pytest.mark.skipif(not os.path.isdir(SRC), reason="the reference is not mounted on this machine")
@pytest.mark.parametrize("name", sorted(EXPECT))
def test_reference_scene_file_reads_unchanged(name, tmp_path):
    d = str(tmp_path)
    text = open(os.path.join(SRC, name)).read()
    shutil.copy(os.path.join(SRC, name), os.path.join(d, name))
    pos, idx, nrm, tng, uv = synth.uv_sphere(32, 16)
    for asset in set(re.findall(r'"([^"]*\.(?:assbin|exr))"', text)):
        p = os.path.join(d, asset)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        if asset.endswith(".exr"):
            capi.write_exr(p, synth.sky_environment(256, 128))
        elif "Cornell_box_RB" in asset:
            scenes.write_cornell_box_rb_asset(d)
        else:
            capi.write_assbin(p, pos, idx, nrm, tng, uv, material_name="m", diffuse=(0.7, 0.6, 0.5))
    with capi.stdout_to_stderr():
        hs = capi.read_scene(os.path.join(d, name))
    w, h, spp, brightness = EXPECT[name]
    c = hs.context
    assert (c["width"], c["height"], c["samples"]) == (w, h, spp) and c["brightness"] == brightness and c["hasRenderer"]
    assert hs.desc.num_triangles >= 60 and hs.desc.num_materials >= 8 and hs.desc.num_lights >= (0 if name in ENV_LIT else 2)
    assert bool(hs.desc.environment.present) == (name in ENV_LIT)
    if name.startswith("RTC3"):
        assert hs.desc.num_instances > 1000        # the scan's hits, one grass instance each
    assert hs.desc.camera.obj_plane_dist > 0


def test_unknown_builtin_fails_loudly(tmp_path):
    """A scene file that calls a function the language does not have must fail with a message, not build a wrong scene."""
    path = scenes.SCENES["diffuse"](str(tmp_path), width=16, height=16, spp=1)
    bad = os.path.join(str(tmp_path), "bad.txt")
    with open(bad, "w") as f:
        f.write(open(path).read() + "\nscatterOnSurface(root, 10);\n")
    with pytest.raises(capi.SlrError, match="scatterOnSurface|not defined"):
        with capi.stdout_to_stderr():
            capi.read_scene(bad)


def test_debug_renderer_is_selected_by_the_scene_language(tmp_path):
    """setRenderer("method": "debug", ("outputs": (...),)) (API.cpp:1037-1062) builds the GPU debug renderer; an unknown
    method still fails like the reference."""
    from slr_b200 import scenes
    path = scenes.SCENES["diffuse"](str(tmp_path), width=32, height=32, spp=1)
    text = open(path).read()
    dbg = os.path.join(str(tmp_path), "debug_scene.txt")
    with open(dbg, "w") as f:
        f.write(text + '\nsetRenderer("method": "debug", ("outputs": ("geometric normal", "shading normal"),));\n')
    hs = capi.read_scene(dbg)
    assert hs.context["hasRenderer"]
    bad = os.path.join(str(tmp_path), "bad_scene.txt")
    with open(bad, "w") as f:
        f.write(text + '\nsetRenderer("method": "photon mapping");\n')
    with pytest.raises(capi.SlrError, match="Unknown method"):
        capi.read_scene(bad)


# every name libSLRSceneGraph/API.cpp puts on the global stack (stack["..."] = ..., API.cpp:246-1090)
REFERENCE_BUILTINS = """root print addItem numElements Point Vector getX getY getZ random min clamp sqrt pow sin cos tan asin acos atan
dot cross distance translate rotate rotateX rotateY rotateZ scale lookAt AnimatedTransform Texture2DMapping Texture3DMapping
SpectrumTexture NormalTexture FloatTexture createVertex Spectrum Image2D createSurfaceMaterial createEmitterSurfaceProperty
createMesh createNode copyNode createReferenceNode setTransform addChild load3DModel scanXZFromYPlus createPerspectiveCamera
setRenderer setRenderSettings setEnvironment""".split()


def test_every_reference_builtin_is_defined(tmp_path):
    """The scene language knows every global of the reference's (functions are first-class values, so naming one is enough)."""
    path = scenes.SCENES["diffuse"](str(tmp_path), width=16, height=16, spp=1)
    probe = os.path.join(str(tmp_path), "probe.txt")
    with open(probe, "w") as f:
        f.write(open(path).read() + "\n" + "".join(f"probe_{i} = {name};\n" for i, name in enumerate(REFERENCE_BUILTINS)))
    with capi.stdout_to_stderr():
        capi.read_scene(probe)
    if os.path.isdir(SRC):      # the list above is the reference's own, where it can be checked
        api = open(os.path.join(os.path.dirname(SRC), "libSLRSceneGraph", "API.cpp")).read()
        assert sorted(set(re.findall(r'stack\["(\w+)"\]', api))) == sorted(REFERENCE_BUILTINS)


SHOW = """
function slrB200ProbeShow(p, t, b, n) {
    print(getX(p)); print(getY(p)); print(getZ(p));
    print(getX(t)); print(getY(t)); print(getZ(t));
    print(getX(b)); print(getY(b)); print(getZ(b));
    print(getX(n)); print(getY(n)); print(getZ(n));
}
scanXZFromYPlus(root, 14, 14, 0.37, slrB200ProbeShow);
"""


def _numbers(text):
    out = []
    for line in text.splitlines():
        try:
            out.append(float(line.strip()))
        except ValueError:
            pass
    return out


@pytest.mark.skipif(not os.path.isdir(SRC) or not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_render")),
                    reason="needs the mounted reference and its compiled oracle")
@pytest.mark.parametrize("name", sorted(EXPECT))
def test_reference_scene_geometry_matches_live_reference(name, tmp_path):
    """The WHOLE scene a shipped scene file describes, as seen by a 14 x 14 grid of rays from above: a scan appended to a
    temporary copy of the file makes both interpreters print every hit's position and shading frame after the scene graph is
    complete (loops, integer arithmetic, tuples, createMesh, load3DModel with material callbacks, transforms, reference nodes,
    RTC3's own 60 x 60 scan with random() rotations). Compared number for number with the compiled reference, live."""
    import subprocess
    import sys
    import numpy as np
    d = str(tmp_path)
    text = open(os.path.join(SRC, name)).read()
    with open(os.path.join(d, name), "w") as f:
        f.write(text + SHOW)
    pos, idx, nrm, tng, uv = synth.uv_sphere(32, 16)
    for asset in set(re.findall(r'"([^"]*\.(?:assbin|exr))"', text)):
        p = os.path.join(d, asset)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        if asset.endswith(".exr"):
            capi.write_exr(p, synth.sky_environment(64, 32))
        elif "Cornell_box_RB" in asset:
            scenes.write_cornell_box_rb_asset(d)
        else:
            capi.write_assbin(p, pos, idx, nrm, tng, uv, material_name="m", diffuse=(0.7, 0.6, 0.5))
    ref = subprocess.run([os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_render"), name, "out.bin", "1", "8", "8"],
                         cwd=d, capture_output=True, text=True, timeout=600)
    code = ("import sys; sys.path.insert(0, %r)\nfrom slr_b200 import capi\ncapi.read_scene(%r)\n" %
            (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(d, name)))
    host = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert host.returncode == 0, host.stderr[-1500:]
    want, got = np.array(_numbers(ref.stdout)), np.array(_numbers(host.stdout))
    assert want.size >= 12 * 20, f"the reference printed only {want.size} numbers: {ref.stderr[-300:]}"
    assert got.shape == want.shape, f"{got.size} numbers against the reference's {want.size}"
    assert np.allclose(got, want, rtol=3e-5, atol=3e-5), float(np.abs(got - want).max())
