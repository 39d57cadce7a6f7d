"""scanXZFromYPlus and random() of the scene language against the reference itself (CPU, no GPU).

scanXZFromYPlus (libSLRSceneGraph/API.cpp:926-983) casts a grid of rays straight down onto a node WHILE the scene file is
being read and calls a script function with the hit's position and shading frame (RTC3*.txt scatter instances with it).
The host interpreter flattens the subtree, builds its SBVH -> QBVH and casts the rays on the host (host/raycast.cpp).
Parity: the same scene file is read by the compiled reference (oracle/_ref/ref_render prints while it parses) and by
libslrhost; the callback prints every component it receives, and the two transcripts must agree number for number
(the reference prints 6 significant digits). Fixtures: tests/golden/scan_builtin.txt and scan_copy.txt (a second scene
that also exercises copyNode), the reference's transcripts, made
by `python tests/test_scan_builtin.py --make-golden` in the container that has /root/reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_util as ou  # noqa: E402
from slr_b200 import capi, scenes, synth  # noqa: E402

def golden_path(which):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", which + ".txt")

SCRIPT = """
function terrainMat(name, attrs) {
    return createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.5, 0.5, 0.5)),));
}
terrain = load3DModel("models/terrain.assbin", terrainMat);
setTransform(terrain, translate(0.25, 0.5, -0.125) * rotateY(0.3) * scale(2.0, 1.5, 2.0));
ball = load3DModel("models/ball.assbin", terrainMat);
ballRef = createReferenceNode(ball);
inst = createNode();
addChild(inst, ballRef);
setTransform(inst, translate(0.6, 1.2, 0.4) * scale(0.35));
group = createNode();
setTransform(group, translate(0, 0, 0));
addChild(group, terrain);
addChild(group, inst);
addChild(root, group);

function show(p, t, b, n) {
    print(getX(p)); print(getY(p)); print(getZ(p));
    print(getX(t)); print(getY(t)); print(getZ(t));
    print(getX(b)); print(getY(b)); print(getZ(b));
    print(getX(n)); print(getY(n)); print(getZ(n));
    print(random());
}
scanXZFromYPlus(group, 7, 5, 0.5, show);
scanXZFromYPlus(terrain, 3, 3, show);

%(light)s
cameraNode = createNode();
camera = createPerspectiveCamera("aspect": 1.0, "fovY": 0.7, "radius": 0.01, "imgDist": 1.0, "objDist": 4.0);
addChild(cameraNode, camera);
setTransform(cameraNode, translate(0.0, 3.0, 6.0) * rotateY(3.1415926536) * rotateX(0.4));
addChild(root, cameraNode);
setRenderer("method": "PT", ("samples": 1,));
setRenderSettings("width": 8, "height": 8);
"""


# copyNode (API.cpp:736-744, nodes.cpp:69-77): a deep copy that can be placed on its own; scanned together with the original.
# Its transform goes through lookAt / rotateX / rotateZ / rotate(angle, axis) (builtin_transform.cpp), seen through the hits.
SCRIPT_COPY = SCRIPT.replace("scanXZFromYPlus(group, 7, 5, 0.5, show);\nscanXZFromYPlus(terrain, 3, 3, show);",
                             "terrain2 = copyNode(terrain);\n"
                             "setTransform(terrain2, lookAt((2.5, 0.9, 0.3), (2.1, 0.7, -5.0), (0.05, 1, 0)) * rotateX(0.15) * rotateZ(-0.1) *\n"
                             "                        rotate(0.2, Vector(1, 1, 0)) * scale(1.2));\n"
                             "pair = createNode();\nsetTransform(pair, translate(0, 0, 0));\n"
                             "addChild(pair, terrain2);\naddChild(root, pair);\n"
                             "scanXZFromYPlus(pair, 5, 4, 0.3, show);\nscanXZFromYPlus(root, 6, 6, show);")
assert SCRIPT_COPY != SCRIPT
SCRIPTS = {"scan_builtin": SCRIPT, "scan_copy": SCRIPT_COPY}


def write_scene(directory, which="scan_builtin"):
    os.makedirs(os.path.join(directory, "models"), exist_ok=True)
    pos, idx = synth.heightfield(24)
    nrm = np.zeros_like(pos)
    # smooth vertex normals of the heightfield (area-weighted), tangents along +x
    p = pos.astype(np.float64)
    fn = np.cross(p[idx[:, 1]] - p[idx[:, 0]], p[idx[:, 2]] - p[idx[:, 0]])
    acc = np.zeros_like(p)
    for k in range(3):
        np.add.at(acc, idx[:, k], fn)
    acc /= np.linalg.norm(acc, axis=1, keepdims=True)
    nrm = acc.astype(np.float32)
    tng = np.tile(np.array([1, 0, 0], np.float32), (pos.shape[0], 1))
    uv = np.ascontiguousarray(pos[:, [0, 2]], np.float32)
    capi.write_assbin(os.path.join(directory, "models", "terrain.assbin"), pos, idx, nrm, tng, uv, material_name="terrain", diffuse=(0.5, 0.5, 0.5))
    bp, bi, bn, bt, buv = synth.displaced_sphere(24, 12)
    capi.write_assbin(os.path.join(directory, "models", "ball.assbin"), bp, bi, bn, bt, buv, material_name="ball", diffuse=(0.7, 0.7, 0.7))
    # an area light above the terrain: the reference's renderer (run after parsing by ref_render) needs one
    light = "lightNode = createNode();\nsetTransform(lightNode, translate(0, 0, 0));\n" + scenes._quad(
        "lightMesh", [(-1, 5, -1), (1, 5, -1), (1, 5, 1), (-1, 5, 1)], (0, -1, 0), (1, 0, 0),
        ['scatterMat = createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.9, 0.9, 0.9)),));',
         'emitterMat = createEmitterSurfaceProperty("diffuse", (SpectrumTexture(Spectrum("ID": "D65") * 6),));',
         'surfMat = createSurfaceMaterial("emitter", (scatterMat, emitterMat));']).replace("CBNode", "lightNode") + "addChild(root, lightNode);\n"
    path = os.path.join(directory, "scan_scene.txt")
    with open(path, "w") as f:
        f.write(SCRIPTS[which] % {"light": light})
    return path


def numbers(text):
    out = []
    for line in text.splitlines():
        try:
            out.append(float(line.strip()))
        except ValueError:
            pass
    return np.array(out)


def reference_transcript(path):
    p = subprocess.run([os.path.join(ou.REF_DIR, "ref_render"), os.path.basename(path), "out.bin", "1", "8", "8"],
                       cwd=os.path.dirname(path), capture_output=True, text=True)
    return p.stdout


def host_transcript(path):
    code = ("import sys; sys.path.insert(0, %r); from slr_b200 import capi; capi.read_scene(%r)" %
            (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), path))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_random_builtin_matches_reference_stream(tmp_path):
    """random() = XORShiftRNG(2112984105).getFloat0cTo1o() (API.cpp:238-244); the seeding loop works on a SIGNED seed
    (XORShiftRNG.cpp:21-27). First eight values printed by the reference (oracle/_ref/ref_render on eight print(random()))."""
    want = [0.207131, 0.321073, 0.347968, 0.7741, 0.778341, 0.107025, 0.879004, 0.486558]
    path = os.path.join(str(tmp_path), "r.txt")
    with open(path, "w") as f:
        f.write("print(random());\n" * 8)
    code = ("import sys; sys.path.insert(0, %r); from slr_b200 import capi\ntry:\n    capi.read_scene(%r)\nexcept capi.SlrError:\n    pass\n" %
            (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), path))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    got = numbers(out)
    assert got.shape == (8,) and np.allclose(got, want, rtol=0, atol=5e-7), out


@pytest.mark.parametrize("which", sorted(SCRIPTS))
def test_scan_builtin_matches_golden(which, tmp_path):
    path = write_scene(str(tmp_path), which)
    got = numbers(host_transcript(path))
    want = numbers(open(golden_path(which)).read())
    assert want.size >= 13 * 20, "the golden holds a useful number of hits"
    assert got.shape == want.shape, f"{got.size} numbers printed, the reference printed {want.size}"
    # 6 significant digits in the transcript; positions within 2e-5 relative, unit vectors within 2e-5 absolute
    assert np.allclose(got, want, rtol=2e-5, atol=2e-5), np.abs(got - want).max()


@pytest.mark.skipif(not ou.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("which", sorted(SCRIPTS))
def test_scan_builtin_matches_live_reference(which, tmp_path):
    path = write_scene(str(tmp_path), which)
    want = numbers(reference_transcript(path))
    got = numbers(host_transcript(path))
    assert got.shape == want.shape and want.size > 0
    assert np.allclose(got, want, rtol=2e-5, atol=2e-5), np.abs(got - want).max()


if __name__ == "__main__" and "--make-golden" in sys.argv:
    import tempfile
    for which in sorted(SCRIPTS):
        with tempfile.TemporaryDirectory() as d:
            text = reference_transcript(write_scene(d, which))
        keep = [l for l in text.splitlines() if l.strip() and numbers(l).size == 1]
        with open(golden_path(which), "w") as f:
            f.write("\n".join(keep) + "\n")
        print(len(keep), "numbers written to", golden_path(which))
