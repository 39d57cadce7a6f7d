"""Generators for the benchmark / test scenes, written out as SLR scene-description files plus the
synthetic assets they load (the reference's TestScenes need models/*.assbin and images/*.exr that are
not distributed with it). The scene text is emitted by this module; geometry and material values
follow the configurations named in BASELINE.json / SURVEY.md section 8d.
"""
import os

import numpy as np

from . import capi, synth


def _quad(name, verts, normal, tangent, mat_setup, indent=""):
    """One createMesh statement for a quad (two triangles), with its material statements before it."""
    lines = list(mat_setup)
    uv = [(0, 0), (1, 0), (1, 1), (0, 1)]
    vs = ",\n".join(f"    (({v[0]}, {v[1]}, {v[2]}), ({normal[0]}, {normal[1]}, {normal[2]}), "
                    f"({tangent[0]}, {tangent[1]}, {tangent[2]}), ({uv[i][0]}, {uv[i][1]}))" for i, v in enumerate(verts))
    lines.append(f"{name} = createMesh(\n  (\n{vs}\n  ),\n  (\n    (surfMat, ((0, 1, 2), (0, 2, 3))),\n  )\n);")
    lines.append(f"addChild(CBNode, {name});")
    return "\n".join(lines) + "\n"


def _matte(r, g, b):
    return [f"diffuseTex = SpectrumTexture(Spectrum({r}, {g}, {b}));", 'surfMat = createSurfaceMaterial("matte", (diffuseTex,));']


def cornell_box_shell():
    """The five walls and the area light of the 3 x 2.5 x 5.1 Cornell box (config C1 geometry)."""
    t = 'CBNode = createNode();\nsetTransform(CBNode, translate(0, 0, 0));\n\n'
    t += _quad("leftWall", [(-1.5, 0, 2.55), (-1.5, 0, -2.55), (-1.5, 2.5, -2.55), (-1.5, 2.5, 2.55)], (1, 0, 0), (0, 0, -1), _matte(0.75, 0.25, 0.25))
    t += _quad("rightWall", [(1.5, 0, -2.55), (1.5, 0, 2.55), (1.5, 2.5, 2.55), (1.5, 2.5, -2.55)], (-1, 0, 0), (0, 0, 1), _matte(0.25, 0.25, 0.75))
    t += _quad("floor", [(-1.5, 0, 2.55), (1.5, 0, 2.55), (1.5, 0, -2.55), (-1.5, 0, -2.55)], (0, 1, 0), (1, 0, 0), _matte(0.75, 0.75, 0.75))
    t += _quad("innerWall", [(-1.5, 0, -2.55), (1.5, 0, -2.55), (1.5, 2.5, -2.55), (-1.5, 2.5, -2.55)], (0, 0, 1), (1, 0, 0), _matte(0.75, 0.75, 0.75))
    t += _quad("ceiling", [(-1.5, 2.5, -2.55), (1.5, 2.5, -2.55), (1.5, 2.5, 2.55), (-1.5, 2.5, 2.55)], (0, -1, 0), (1, 0, 0), _matte(0.75, 0.75, 0.75))
    light = ["diffuseTex = SpectrumTexture(Spectrum(0.9, 0.9, 0.9));",
             'scatterMat = createSurfaceMaterial("matte", (diffuseTex,));',
             'difLightTex = SpectrumTexture(Spectrum("ID": "D65") * 4);',
             'emitterMat = createEmitterSurfaceProperty("diffuse", (difLightTex,));',
             'surfMat = createSurfaceMaterial("emitter", (scatterMat, emitterMat));']
    t += _quad("lightMesh", [(-0.5, 2.499, -0.5), (0.5, 2.499, -0.5), (0.5, 2.499, 0.5), (-0.5, 2.499, 0.5)], (0, -1, 0), (1, 0, 0), light)
    t += "addChild(root, CBNode);\n\n"
    return t


CORNELL_CAMERA = """cameraNode = createNode();
camera = createPerspectiveCamera("aspect": 4.0 / 3.0, "fovY": 0.4807705238,
                                 "radius": 0.025, "imgDist": 1.0, "objDist": 6.3);
addChild(cameraNode, camera);
setTransform(cameraNode, translate(0.0, 1.689714, 6.70284) * rotateY(3.1415926536) * rotateX(0.0563936));
addChild(root, cameraNode);
"""


def write_sphere_asset(directory, segments_u=64, segments_v=32):
    os.makedirs(os.path.join(directory, "models"), exist_ok=True)
    pos, idx, nrm, tng, uv = synth.uv_sphere(segments_u, segments_v)
    path = os.path.join(directory, "models", "sphere.assbin")
    capi.write_assbin(path, pos, idx, nrm, tng, uv, material_name="sphere", diffuse=(0.8, 0.8, 0.8))
    return path


def write_cornell_spheres(directory, width=512, height=512, spp=64, sphere_segments=(64, 32), method="PT"):
    """Config C1: Cornell box with an aluminium mirror sphere and a BK7 glass sphere, thin-lens camera."""
    write_sphere_asset(directory, *sphere_segments)
    t = f'setRenderer("method": "{method}", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height});\n\n'
    t += cornell_box_shell()
    t += """function leftSphereMat(name, attrs) {
    eta = SpectrumTexture(Spectrum("ID": "Aluminium", 0));
    k = SpectrumTexture(Spectrum("ID": "Aluminium", 1));
    return createSurfaceMaterial("metal", (SpectrumTexture(Spectrum("Reflectance", 1.0)), eta, k));
}
leftSphereNode = load3DModel("models/sphere.assbin", leftSphereMat);
setTransform(leftSphereNode, translate(-0.7, 0, -1.05) * scale(0.5) * translate(0, 1, 0));
addChild(CBNode, leftSphereNode);

function rightSphereMat(name, attrs) {
    etaExt = SpectrumTexture(Spectrum("ID": "Air", 0));
    etaInt = SpectrumTexture(Spectrum("ID": "Glass_BK7", 0));
    coeff = SpectrumTexture(Spectrum("Reflectance", 0.999));
    return createSurfaceMaterial("glass", (coeff, etaExt, etaInt));
}
rightSphereNode = load3DModel("models/sphere.assbin", rightSphereMat);
setTransform(rightSphereNode, translate(0.7, 0, 0) * scale(0.5) * translate(0, 1, 0));
addChild(CBNode, rightSphereNode);

"""
    t += CORNELL_CAMERA
    path = os.path.join(directory, "Cornell_Box_Spheres.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


def write_cornell_diffuse(directory, width=128, height=128, spp=16):
    """Smallest end-to-end scene: the empty Cornell box (Lambert walls + D65 area light)."""
    os.makedirs(directory, exist_ok=True)
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height});\n\n'
    t += cornell_box_shell() + CORNELL_CAMERA
    path = os.path.join(directory, "Cornell_Box_Diffuse.txt")
    with open(path, "w") as f:
        f.write(t)
    return path



def write_cornell_box_rb_asset(directory):
    """models/Cornell_box_RB.assbin: the [-1, 1]^3 box of Cornell_Box_ColorChecker.txt, open towards +z,
    white floor / ceiling / back wall, red left and blue right wall (one mesh + material per colour)."""
    os.makedirs(os.path.join(directory, "models"), exist_ok=True)
    q = synth.quad_mesh
    white = [q([(-1, -1, 1), (1, -1, 1), (1, -1, -1), (-1, -1, -1)], (0, 1, 0), (1, 0, 0)),
             q([(-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1)], (0, -1, 0), (1, 0, 0)),
             q([(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1)], (0, 0, 1), (1, 0, 0))]
    pos = np.concatenate([m["positions"] for m in white])
    idx = np.concatenate([m["indices"] + 4 * k for k, m in enumerate(white)])
    wmesh = {"name": "white", "positions": pos, "indices": idx, "normals": np.concatenate([m["normals"] for m in white]),
             "tangents": np.concatenate([m["tangents"] for m in white]), "uvs": np.concatenate([m["uvs"] for m in white]), "material": 0}
    red = q([(-1, -1, 1), (-1, -1, -1), (-1, 1, -1), (-1, 1, 1)], (1, 0, 0), (0, 0, -1))
    red.update(name="red", material=1)
    blue = q([(1, -1, -1), (1, -1, 1), (1, 1, 1), (1, 1, -1)], (-1, 0, 0), (0, 0, 1))
    blue.update(name="blue", material=2)
    path = os.path.join(directory, "models", "Cornell_box_RB.assbin")
    synth.write_assbin_scene(path, [wmesh, red, blue],
                             [{"name": "white", "diffuse": (0.75, 0.75, 0.75)}, {"name": "red", "diffuse": (0.75, 0.25, 0.25)},
                              {"name": "blue", "diffuse": (0.25, 0.25, 0.75)}])
    return path


_RB_LIGHT = """    lightNode = createNode();
    setTransform(lightNode, translate(0.0, 0.999, 0.0));
        diffuseCol = Spectrum(0.9, 0.9, 0.9);
        diffuseTex = SpectrumTexture(diffuseCol);
        scatterMat = createSurfaceMaterial("matte", (diffuseTex,));
        difLightCol = Spectrum("ID": "D65");
        difLightTex = SpectrumTexture(difLightCol);
        emitterMat = createEmitterSurfaceProperty("diffuse", (difLightTex,));
        surfMat = createSurfaceMaterial("emitter", (scatterMat, emitterMat));

        lightMesh = createMesh(
            (
            ((-0.25, 0, -0.25), (0, -1, 0), (1, 0, 0), (0, 0)),
            (( 0.25, 0, -0.25), (0, -1, 0), (1, 0, 0), (1, 0)),
            (( 0.25, 0,  0.25), (0, -1, 0), (1, 0, 0), (1, 1)),
            ((-0.25, 0,  0.25), (0, -1, 0), (1, 0, 0), (0, 1))
            ),
            (
            (surfMat, ((0, 1, 2), (0, 2, 3))),
            )
            );
        addChild(lightNode, lightMesh);
    addChild(CBNode, lightNode);
addChild(root, CBNode);
"""

_COLOR_CHECKER = """// Color Checker Materials
mats = (,);
for (i = 0; i < 24; ++i) {
    sp = Spectrum("ID": "ColorChecker", i);
    difTex = SpectrumTexture(sp);
    scatterMat = createSurfaceMaterial("matte", (difTex,));
    addItem(mats, scatterMat);
}

// Create Color Checker
vertices = (,);
matGroups = (,);
for (i = 0; i < 24; ++i) {
    transform = %s;

    addItem(vertices, transform * createVertex((0.0, 0.0, 0.0), %s, (1, 0, 0), (0, 0)));
    addItem(vertices, transform * createVertex((1.0, 0.0, 0.0), %s, (1, 0, 0), (1, 0)));
    addItem(vertices, transform * createVertex(%s, %s, (1, 0, 0), (1, 1)));
    addItem(vertices, transform * createVertex(%s, %s, (1, 0, 0), (0, 1)));

    idxBase = 4 * i;
    addItem(matGroups, (mats[i], ((idxBase + 0, idxBase + 1, idxBase + 2), (idxBase + 0, idxBase + 2, idxBase + 3))));
}
colorCheckerMesh = createMesh(vertices, matGroups);
"""


def write_cornell_materials(directory, width=1024, height=1024, spp=256):
    """Config C2: the Cornell_Box_ColorChecker.txt layout (Cornell_box_RB model, D65 light quad, colour
    checker on the back wall) with the material override script of SURVEY.md section 8d: GGX conductor
    with a checker roughness texture, Ward, Oren-Nayar with checker / Voronoi colour and normal textures,
    Ashikhmin-Shirley, a mixed material and rough glass."""
    write_cornell_box_rb_asset(directory)
    write_sphere_asset(directory)
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height}, "brightness": 4.0);\n\n'
    t += """function CornellBoxMaterial(name, attrs) {
    difCol = attrs["diffuse color"];
    if (name == "white") {
        c0 = Spectrum(difCol[0], difCol[1], difCol[2]);
        c1 = Spectrum(0.35, 0.45, 0.35);
        difTex = SpectrumTexture("checker board", (c0, c1));
        sigma = FloatTexture(0.6);
        normalTex = NormalTexture("checker board", (0.05, false));
        return (createSurfaceMaterial("matte", (difTex, sigma)), normalTex);
    }
    if (name == "red") {
        difTex = SpectrumTexture("voronoi", (0.35, 0.8));
        return createSurfaceMaterial("matte", (difTex, FloatTexture(0.3)));
    }
    difTex = SpectrumTexture(Spectrum(difCol[0], difCol[1], difCol[2]));
    ax = FloatTexture(0.1);
    ay = FloatTexture(0.3);
    return createSurfaceMaterial("Ward", (difTex, ax, ay));
}

CBNode = load3DModel("models/Cornell_box_RB.assbin", CornellBoxMaterial);
""" + _RB_LIGHT + "\n" + (_COLOR_CHECKER % (
        "translate(0, 0, -0.999) * scale(0.9 / 3.0) *\n                translate(-3.0 + (i % 6), 1.0 - (i / 6), 0.0)",
        "(0, 0, 1)", "(0, 0, 1)", "(1.0, 1.0, 0.0)", "(0, 0, 1)", "(0.0, 1.0, 0.0)", "(0, 0, 1)")) + """addChild(CBNode, colorCheckerMesh);

function ggxMetal(name, attrs) {
    eta = SpectrumTexture(Spectrum("ID": "Gold", 0));
    k = SpectrumTexture(Spectrum("ID": "Gold", 1));
    alpha = FloatTexture("checker board", (0.05, 0.3));
    return createSurfaceMaterial("microfacet metal", (eta, k, alpha));
}
s0 = load3DModel("models/sphere.assbin", ggxMetal);
setTransform(s0, translate(-0.55, -0.65, -0.3) * scale(0.35));
addChild(CBNode, s0);

function ashikhmin(name, attrs) {
    Rd = SpectrumTexture(Spectrum(0.1, 0.5, 0.7));
    Rs = SpectrumTexture(Spectrum(0.1, 0.1, 0.1));
    return (createSurfaceMaterial("Ashikhmin", (Rd, Rs, FloatTexture(1000), FloatTexture(100))), NormalTexture("voronoi", (0.08, 0.4)));
}
s1 = load3DModel("models/sphere.assbin", ashikhmin);
setTransform(s1, translate(0.45, -0.7, 0.1) * scale(0.3));
addChild(CBNode, s1);

function mixedMat(name, attrs) {
    m0 = createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.7, 0.6, 0.2)),));
    eta = SpectrumTexture(Spectrum("ID": "Copper", 0));
    k = SpectrumTexture(Spectrum("ID": "Copper", 1));
    m1 = createSurfaceMaterial("metal", (SpectrumTexture(Spectrum("Reflectance", 0.95)), eta, k));
    return createSurfaceMaterial("mix", (m0, m1, FloatTexture("voronoi", (0.15, 1.0, true))));
}
s2 = load3DModel("models/sphere.assbin", mixedMat);
setTransform(s2, translate(-0.1, -0.75, 0.55) * scale(0.25));
addChild(CBNode, s2);

function roughGlass(name, attrs) {
    etaExt = SpectrumTexture(Spectrum("ID": "Air", 0));
    etaInt = SpectrumTexture(Spectrum("ID": "Glass_BK7", 0));
    return createSurfaceMaterial("microfacet glass", (etaExt, etaInt, FloatTexture(0.15)));
}
s3 = load3DModel("models/sphere.assbin", roughGlass);
setTransform(s3, translate(0.55, 0.2, -0.4) * scale(0.28));
addChild(CBNode, s3);

cameraNode = createNode();
    camera = createPerspectiveCamera("aspect": 1.0, "fovY": 0.5235987756, "radius": 0.025,
                                     "imgDist": 1.0, "objDist": 5);
    addChild(cameraNode, camera);
setTransform(cameraNode, translate(0, 0, 5) * rotateY(-3.1415926536));

addChild(root, cameraNode);
"""
    path = os.path.join(directory, "Cornell_Box_ColorChecker.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


def write_ibl_test(directory, width=1024, height=1024, spp=256, env_size=(2048, 1024)):
    """Config C3: the IBL_Test.txt layout -- HDR environment with importance sampling, colour checker on
    the ground, aluminium sphere, an Ashikhmin-Shirley object (a synthetic bumpy ball stands in for the
    Kirby model), thin-lens camera, whole scene rotated."""
    write_sphere_asset(directory)
    os.makedirs(os.path.join(directory, "images"), exist_ok=True)
    os.makedirs(os.path.join(directory, "models", "Kirby_Pikachu_Hat"), exist_ok=True)
    capi.write_exr(os.path.join(directory, "images", "Malibu_Overlook_3k_corrected.exr"), synth.sky_environment(*env_size))
    pos, idx, nrm, tng, uv = synth.displaced_sphere(96, 48)
    capi.write_assbin(os.path.join(directory, "models", "Kirby_Pikachu_Hat", "pikachu_hat_corrected.assbin"),
                      pos * 0.8 + np.array([1.2, 0.8, 0.6], np.float32), idx, nrm, tng, uv, material_name="hat", diffuse=(0.8, 0.7, 0.1))
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height}, "brightness": 4.0);\n\n'
    t += 'setEnvironment("images/Malibu_Overlook_3k_corrected.exr");\n\n'
    t += _COLOR_CHECKER % ("translate(0, -1, 0) * scale(0.9 / 3.0) *\n                translate(-3.0 + (i % 6), 0.0, (i / 6) - 1.0)",
                           "(0, 1, 0)", "(0, 1, 0)", "(1.0, 0.0, -1.0)", "(0, 1, 0)", "(0.0, 0.0, -1.0)", "(0, 1, 0)")
    t += """addChild(root, colorCheckerMesh);

function sphereMaterial(name, attrs) {
    coeffR = SpectrumTexture(Spectrum("type": "Reflectance", 0.99));
    eta = SpectrumTexture(Spectrum("ID": "Aluminium", 0));
    k = SpectrumTexture(Spectrum("ID": "Aluminium", 1));
    return createSurfaceMaterial("metal", (coeffR, eta, k));
}

sphereNode = load3DModel("models/sphere.assbin", sphereMaterial);
setTransform(sphereNode, translate(0, 0.5, 0) * scale(0.4));
addChild(root, sphereNode);

function PikachuMaterial(name, attrs) {
    difTex = 0;
    if (numElements(attrs["diffuse textures"])) {
        difTexPaths = attrs["diffuse textures"];
        image = Image2D(difTexPaths[0]);
        difTex = SpectrumTexture(image);
    }
    else {
        difCol = attrs["diffuse color"];
        difTex = SpectrumTexture(Spectrum(difCol[0], difCol[1], difCol[2]));
    }
    speTex = SpectrumTexture(Spectrum(0.1, 0.1, 0.1));
    nxTex = nyTex = FloatTexture(1000);
    return createSurfaceMaterial("Ashikhmin", (difTex, speTex, nxTex, nyTex));
}

kirbyNode = load3DModel("models/Kirby_Pikachu_Hat/pikachu_hat_corrected.assbin", PikachuMaterial);
setTransform(kirbyNode, translate(0, -1, 0) * scale(0.5));
addChild(root, kirbyNode);

cameraNode = createNode();
    camera = createPerspectiveCamera("aspect": 1.0, "fovY": 0.5235987756, "radius": 0.025,
                                     "imgDist": 1.0, "objDist": 4.75);
    addChild(cameraNode, camera);
setTransform(cameraNode, translate(0, 0, 5) * rotateY(3.1415926536));

addChild(root, cameraNode);
setTransform(root, rotateY(1.2));
"""
    path = os.path.join(directory, "IBL_Test.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


def write_instanced(directory, width=1920, height=1080, spp=1024, base_segments=(224, 112), grid=10, env=False):
    """Config C4 shape: one procedural base mesh (a bumpy ball, 2 * su * (sv - 1) triangles) instanced
    grid x grid times through createReferenceNode on a jittered grid (LCG seed 12345) over a ground quad,
    one D65 area light. The default base mesh has 49,728 triangles; base_segments=(318, 159) gives
    100,488 (x 100 instances = 10 M). Two-level BVH: top-level over the instances, one shared nested BVH."""
    os.makedirs(os.path.join(directory, "models"), exist_ok=True)
    pos, idx, nrm, tng, uv = synth.displaced_sphere(*base_segments)
    capi.write_assbin(os.path.join(directory, "models", "bumpy_ball.assbin"), pos, idx, nrm, tng, uv, material_name="ball", diffuse=(0.7, 0.7, 0.7))
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height});\n\n'
    ext = 1.6 * grid
    t += "groundNode = createNode();\nsetTransform(groundNode, translate(0, 0, 0));\n"
    t += _quad("ground", [(-ext, 0, ext), (ext, 0, ext), (ext, 0, -ext), (-ext, 0, -ext)], (0, 1, 0), (1, 0, 0),
               ['diffuseTex = SpectrumTexture("checker board", (Spectrum(0.7, 0.7, 0.7), Spectrum(0.3, 0.3, 0.35)));',
                'surfMat = createSurfaceMaterial("matte", (diffuseTex,));']).replace("CBNode", "groundNode")
    light = ['scatterMat = createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.9, 0.9, 0.9)),));',
             'emitterMat = createEmitterSurfaceProperty("diffuse", (SpectrumTexture(Spectrum("ID": "D65") * 6),));',
             'surfMat = createSurfaceMaterial("emitter", (scatterMat, emitterMat));']
    h = 9.0
    t += _quad("lightMesh", [(-ext / 2, h, -ext / 2), (ext / 2, h, -ext / 2), (ext / 2, h, ext / 2), (-ext / 2, h, ext / 2)], (0, -1, 0), (1, 0, 0), light).replace("CBNode", "groundNode")
    t += "addChild(root, groundNode);\n\n"
    t += """function ballMat(name, attrs) {
    difTex = SpectrumTexture(Spectrum(0.75, 0.55, 0.3));
    speTex = SpectrumTexture(Spectrum(0.08, 0.08, 0.08));
    return createSurfaceMaterial("Ashikhmin", (difTex, speTex, FloatTexture(200), FloatTexture(200)));
}
ballNode = load3DModel("models/bumpy_ball.assbin", ballMat);
ballRef = createReferenceNode(ballNode);
"""
    state = 12345
    def lcg():
        nonlocal state
        state = (state * 1103515245 + 12345) & 0x7FFFFFFF
        return state / float(0x80000000)
    for j in range(grid):
        for i in range(grid):
            x = (i - (grid - 1) / 2) * 2.6 + (lcg() - 0.5) * 0.6
            z = (j - (grid - 1) / 2) * 2.6 + (lcg() - 0.5) * 0.6
            sc = 0.8 + 0.4 * lcg()
            rot = 6.2831853 * lcg()
            t += (f"inst = createNode();\naddChild(inst, ballRef);\n"
                  f"setTransform(inst, translate({x:.5f}, {1.08 * sc:.5f}, {z:.5f}) * rotateY({rot:.5f}) * scale({sc:.5f}));\naddChild(root, inst);\n")
    dist = 1.9 * grid
    t += f"""
cameraNode = createNode();
camera = createPerspectiveCamera("aspect": {width / height:.6f}, "fovY": 0.7, "radius": 0.02, "imgDist": 1.0, "objDist": {dist:.3f});
addChild(cameraNode, camera);
setTransform(cameraNode, translate(0.0, {0.55 * dist:.3f}, {0.85 * dist:.3f}) * rotateY(3.1415926536) * rotateX(0.55));
addChild(root, cameraNode);
"""
    path = os.path.join(directory, "Instanced_Balls.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


def _small_instanced(directory, width=128, height=128, spp=64):
    return write_instanced(directory, width, height, spp, base_segments=(48, 24), grid=4)


def write_nested(directory, width=128, height=128, spp=64):
    """Instancing nested in instancing (TransformedSurfaceObject over an aggregate that itself holds TransformedSurfaceObjects,
    SurfaceObject.cpp:307-336): a `cluster` = three references to the bumpy ball (rotated / scaled) + a pedestal of its own
    triangles + a small emitting panel of its own, referenced three times from the root (rotated, one of them scaled), over a
    ground quad under a dim ceiling light. The panels are lights two instance levels below the root for the balls' shading and
    one level below for selection. No emitter sits under a scale (the reference's area pdfs are object-space)."""
    os.makedirs(os.path.join(directory, "models"), exist_ok=True)
    pos, idx, nrm, tng, uv = synth.displaced_sphere(40, 20)
    capi.write_assbin(os.path.join(directory, "models", "bumpy_ball.assbin"), pos, idx, nrm, tng, uv, material_name="ball", diffuse=(0.7, 0.7, 0.7))
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height});\n\n'
    ext = 9.0
    t += "groundNode = createNode();\nsetTransform(groundNode, translate(0, 0, 0));\n"
    t += _quad("ground", [(-ext, 0, ext), (ext, 0, ext), (ext, 0, -ext), (-ext, 0, -ext)], (0, 1, 0), (1, 0, 0),
               ['diffuseTex = SpectrumTexture("checker board", (Spectrum(0.7, 0.7, 0.7), Spectrum(0.3, 0.3, 0.35)));',
                'surfMat = createSurfaceMaterial("matte", (diffuseTex,));']).replace("CBNode", "groundNode")
    light = ['scatterMat = createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.9, 0.9, 0.9)),));',
             'emitterMat = createEmitterSurfaceProperty("diffuse", (SpectrumTexture(Spectrum("ID": "D65") * 1.5),));',
             'surfMat = createSurfaceMaterial("emitter", (scatterMat, emitterMat));']
    t += _quad("lightMesh", [(-3, 8, -3), (3, 8, -3), (3, 8, 3), (-3, 8, 3)], (0, -1, 0), (1, 0, 0), light).replace("CBNode", "groundNode")
    t += "addChild(root, groundNode);\n\n"
    t += """function ballMat(name, attrs) {
    difTex = SpectrumTexture(Spectrum(0.75, 0.55, 0.3));
    return createSurfaceMaterial("matte", (difTex,));
}
ballNode = load3DModel("models/bumpy_ball.assbin", ballMat);
ballRef = createReferenceNode(ballNode);
clusterNode = createNode();
setTransform(clusterNode, translate(0, 0, 0));
"""
    for k, (x, y, z, rot, sc) in enumerate([(-0.9, 0.85, 0.0, 0.3, 0.55), (0.9, 0.95, 0.2, 1.7, 0.65), (0.0, 0.8, -0.9, 4.0, 0.5)]):
        t += (f"b{k} = createNode();\naddChild(b{k}, ballRef);\n"
              f"setTransform(b{k}, translate({x}, {y}, {z}) * rotateY({rot}) * scale({sc}));\naddChild(clusterNode, b{k});\n")
    t += "pedestalNode = createNode();\nsetTransform(pedestalNode, translate(0, 0, 0));\n"
    t += _quad("pedestal", [(-1.6, 0.25, 1.6), (1.6, 0.25, 1.6), (1.6, 0.25, -1.6), (-1.6, 0.25, -1.6)], (0, 1, 0), (1, 0, 0),
               ['surfMat = createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.25, 0.5, 0.7)),));']).replace("CBNode", "pedestalNode")
    panel = ['scatterMat = createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.8, 0.8, 0.8)),));',
             'emitterMat = createEmitterSurfaceProperty("diffuse", (SpectrumTexture(Spectrum("ID": "D65") * 12),));',
             'surfMat = createSurfaceMaterial("emitter", (scatterMat, emitterMat));']
    t += _quad("panel", [(-0.35, 2.2, -0.35), (0.35, 2.2, -0.35), (0.35, 2.2, 0.35), (-0.35, 2.2, 0.35)], (0, -1, 0), (1, 0, 0), panel).replace("CBNode", "pedestalNode")
    t += "addChild(clusterNode, pedestalNode);\nclusterRef = createReferenceNode(clusterNode);\n"
    for k, (x, z, rot) in enumerate([(-3.4, 0.5, 0.4), (3.2, -0.6, 2.2), (0.2, -3.6, 5.1)]):
        t += (f"c{k} = createNode();\naddChild(c{k}, clusterRef);\n"
              f"setTransform(c{k}, translate({x}, 0.0, {z}) * rotateY({rot}));\naddChild(root, c{k});\n")
    t += f"""
cameraNode = createNode();
camera = createPerspectiveCamera("aspect": {width / height:.6f}, "fovY": 0.75, "radius": 0.01, "imgDist": 1.0, "objDist": 11.0);
addChild(cameraNode, camera);
setTransform(cameraNode, translate(0.0, 6.0, 9.5) * rotateY(3.1415926536) * rotateX(0.5));
addChild(root, cameraNode);
"""
    path = os.path.join(directory, "Nested_Instances.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


def write_scatter(directory, width=128, height=128, spp=64, grid=14):
    """The structure of TestScenes/RTC3.txt at a small scale: a terrain mesh, grid x grid instances of a tuft mesh scattered
    over it by scanXZFromYPlus (a ray cast while the file is read; the callback aligns each instance with the surface normal
    and turns it by random()), the reference's two-sided grass material (sum of a matte lobe and an inverse matte lobe),
    an Ashikhmin object, all lit by the synthetic HDR environment; the root rotated as in RTC3.txt."""
    os.makedirs(os.path.join(directory, "models"), exist_ok=True)
    os.makedirs(os.path.join(directory, "images"), exist_ok=True)
    capi.write_exr(os.path.join(directory, "images", "sky.exr"), synth.sky_environment(512, 256))
    pos, idx = synth.heightfield(24)
    p = pos.astype(np.float64)
    fn = np.cross(p[idx[:, 1]] - p[idx[:, 0]], p[idx[:, 2]] - p[idx[:, 0]])
    acc = np.zeros_like(p)
    for k in range(3):
        np.add.at(acc, idx[:, k], fn)
    nrm = (acc / np.linalg.norm(acc, axis=1, keepdims=True)).astype(np.float32)
    tng = np.tile(np.array([1, 0, 0], np.float32), (pos.shape[0], 1))
    capi.write_assbin(os.path.join(directory, "models", "plain.assbin"), pos, idx, nrm, tng, np.ascontiguousarray(pos[:, [0, 2]], np.float32),
                      material_name="plain", diffuse=(0.35, 0.3, 0.2))
    bp, bi, bn, bt, buv = synth.displaced_sphere(16, 8)
    capi.write_assbin(os.path.join(directory, "models", "tuft.assbin"), bp * np.array([0.35, 1.0, 0.35], np.float32) + np.array([0, 1.0, 0], np.float32),
                      bi, bn, bt, buv, material_name="tuft", diffuse=(0.15, 0.35, 0.1))
    cp, ci, cn, ct, cuv = synth.displaced_sphere(32, 16)
    capi.write_assbin(os.path.join(directory, "models", "toy.assbin"), cp, ci, cn, ct, cuv, material_name="toy", diffuse=(0.5, 0.1, 0.1))
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height}, "brightness": 2.0);\n'
    t += 'setEnvironment("images/sky.exr");\n\n'
    t += f"""function plainMaterial(name, attrs) {{
    difCol = attrs["diffuse color"];
    return createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(difCol[0], difCol[1], difCol[2])),));
}}
plain = load3DModel("models/plain.assbin", plainMaterial);
setTransform(plain, translate(-1.0, 0.0, -1.0) * scale(2.0, 1.5, 2.0));
addChild(root, plain);

function grassMaterial(name, attrs) {{
    difCol = attrs["diffuse color"];
    baseColor = Spectrum("Reflectance", "sRGB", 2 * difCol[0], 2 * difCol[1], 2 * difCol[2]);
    r = createSurfaceMaterial("matte", (SpectrumTexture(baseColor * 0.7),));
    tBase = createSurfaceMaterial("matte", (SpectrumTexture(baseColor * 0.3),));
    t = createSurfaceMaterial("inverse", (tBase,));
    return createSurfaceMaterial("sum", (r, t));
}}
grass = load3DModel("models/tuft.assbin", grassMaterial);
grassReference = createReferenceNode(grass);

function scanCallback(p, t, b, n) {{
    trans = translate(getX(p), getY(p), getZ(p));
    axis = cross(Vector(0, 1, 0), n);
    angle = acos(clamp(dot(Vector(0, 1, 0), n), -1, 1));
    if (angle < 0.0001)
        axis = Vector(1, 0, 0);
    rot = rotate(angle, axis);
    sc = scale(0.06);
    rotY = rotateY(2 * 3.1415926536 * random());
    instanceNode = createNode();
    addChild(instanceNode, grassReference);
    setTransform(instanceNode, trans * rot * sc * rotY);
    addChild(root, instanceNode);
}}
scanXZFromYPlus(plain, {grid}, {grid}, 0.6, scanCallback);

function toyMaterial(name, attrs) {{
    difCol = attrs["diffuse color"];
    RdTex = SpectrumTexture(Spectrum(1.5 * difCol[0], 1.5 * difCol[1], 1.5 * difCol[2]));
    RsTex = SpectrumTexture(Spectrum("type": "Reflectance", 0.025));
    nxTex = nyTex = FloatTexture(100.0);
    return createSurfaceMaterial("Ashikhmin", (RdTex, RsTex, nxTex, nyTex));
}}
toy = load3DModel("models/toy.assbin", toyMaterial);
setTransform(toy, translate(0.1, 0.45, 0.2) * rotateY(0.7) * scale(0.22));
addChild(root, toy);

cameraNode = createNode();
camera = createPerspectiveCamera("aspect": {width / height:.6f}, "fovY": 0.6, "radius": 0.0025, "imgDist": 1.0, "objDist": 3.0);
addChild(cameraNode, camera);
setTransform(cameraNode, translate(0.0, 1.6, 2.9) * rotateY(3.1415926536) * rotateX(0.45));
addChild(root, cameraNode);

setTransform(root, rotateY(-1.5707963268));
"""
    path = os.path.join(directory, "Scatter.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


def write_lamps(directory, width=96, height=96, spp=64):
    """Emitters INSIDE an instanced subtree: a lamp (a small emitting quad under a shade) referenced twice with different
    scales above a floor -- the top-level light list then holds the two instances (TransformedSurfaceObject over an aggregate
    that has lights, SurfaceObject.cpp:279-336), each with its nested list."""
    floor = "floorNode = createNode();\nsetTransform(floorNode, translate(0, 0, 0));\n" + \
        _quad("floor", [(-3, 0, 3), (3, 0, 3), (3, 0, -3), (-3, 0, -3)], (0, 1, 0), (1, 0, 0), _matte(0.6, 0.6, 0.6)).replace("CBNode", "floorNode") + \
        "addChild(root, floorNode);\n"
    lamp = "lampNode = createNode();\nsetTransform(lampNode, translate(0, 0, 0));\n" + \
        _quad("lampMesh", [(-0.3, 0, -0.3), (0.3, 0, -0.3), (0.3, 0, 0.3), (-0.3, 0, 0.3)], (0, -1, 0), (1, 0, 0),
              ['scatterMat = createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.9, 0.9, 0.9)),));',
               'emitterMat = createEmitterSurfaceProperty("diffuse", (SpectrumTexture(Spectrum("ID": "D65") * 5),));',
               'surfMat = createSurfaceMaterial("emitter", (scatterMat, emitterMat));']).replace("CBNode", "lampNode") + \
        _quad("shade", [(-0.4, 0.1, -0.4), (0.4, 0.1, -0.4), (0.4, 0.1, 0.4), (-0.4, 0.1, 0.4)], (0, 1, 0), (1, 0, 0), _matte(0.3, 0.3, 0.3)).replace("CBNode", "lampNode")
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height});\n' + floor + lamp
    t += "lampRef = createReferenceNode(lampNode);\n"
    for i, (x, y, z, sc) in enumerate([(-1, 2, 0, 1.0), (1.2, 2.5, 0.5, 2.0)]):
        t += (f"i{i} = createNode();\naddChild(i{i}, lampRef);\nsetTransform(i{i}, translate({x}, {y}, {z}) * rotateZ(0.2) * scale({sc}));\n"
              f"addChild(root, i{i});\n")
    t += ('cameraNode = createNode();\ncamera = createPerspectiveCamera("aspect": 1.0, "fovY": 0.7, "radius": 0.01, "imgDist": 1.0, "objDist": 4.0);\n'
          'addChild(cameraNode, camera);\nsetTransform(cameraNode, translate(0.0, 1.5, 6.0) * rotateY(3.1415926536));\naddChild(root, cameraNode);\n')
    path = os.path.join(directory, "Lamps.txt")
    os.makedirs(directory, exist_ok=True)
    with open(path, "w") as f:
        f.write(t)
    return path


def write_reference_scene(src_path, directory, width=0, height=0, spp=0, method="PT"):
    """One of the reference's shipped TestScenes/*.txt, byte for byte, with nothing but an override block APPENDED
    (renderer method / samples and image size, the way SURVEY.md section 7 hard part 7 prescribes -- the last
    setRenderer / setRenderSettings wins in both interpreters) and the synthetic stand-ins for the assets the file loads
    (the reference does not ship its models / environment maps, README.md:69-72). `src_path` is the caller's copy of
    the file; this module never looks for the reference itself."""
    import re
    os.makedirs(directory, exist_ok=True)
    with open(src_path) as f:
        text = f.read()
    for asset in sorted(set(re.findall(r'"([^"]*\.(?:assbin|exr))"', text))):
        p = os.path.join(directory, asset)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        if asset.endswith(".exr"):
            capi.write_exr(p, synth.sky_environment(2048, 1024))
        elif "Cornell_box_RB" in asset:
            write_cornell_box_rb_asset(directory)
        elif asset.endswith("sphere.assbin"):
            write_sphere_asset(directory)
        else:
            pos, idx, nrm, tng, uv = synth.displaced_sphere(96, 48)
            capi.write_assbin(p, pos, idx, nrm, tng, uv, material_name="m", diffuse=(0.7, 0.6, 0.5))
    override = "\n// ---- appended override (everything above is the reference's file, unchanged) ----\n"
    if spp:
        override += f'setRenderer("method": "{method}", ("samples": {spp},));\n'
    if width and height:
        override += f'setRenderSettings("width": {width}, "height": {height});\n'
    path = os.path.join(directory, os.path.basename(src_path))
    with open(path, "w") as f:
        f.write(text + override)
    return path


def _alpha_quad(name, parent, verts, normal, tangent, uv_scale, mat_setup, alpha_expr):
    """A quad whose material group carries an alpha texture ("alpha": FloatTexture, createMesh's matGroups signature,
    libSLRSceneGraph/API.cpp:672-679): hits where the texture evaluates to 0 are passed through (TriangleMesh.cpp:160-168)."""
    uv = [(0, 0), (uv_scale, 0), (uv_scale, uv_scale), (0, uv_scale)]
    vs = ",\n".join(f"    (({v[0]}, {v[1]}, {v[2]}), ({normal[0]}, {normal[1]}, {normal[2]}), "
                    f"({tangent[0]}, {tangent[1]}, {tangent[2]}), ({uv[i][0]}, {uv[i][1]}))" for i, v in enumerate(verts))
    lines = list(mat_setup) + [f"alphaTex = {alpha_expr};",
                               f'{name} = createMesh(\n  (\n{vs}\n  ),\n  (\n    (surfMat, "alpha": alphaTex, ((0, 1, 2), (0, 2, 3))),\n  )\n);',
                               f"addChild({parent}, {name});"]
    return "\n".join(lines) + "\n"


def write_cutout(directory, width=128, height=128, spp=64):
    """Alpha-mapped (cut-out) geometry: the Cornell box with a checker-perforated screen between camera and back wall (so
    camera rays, bounce rays and shadow rays all meet the alpha test) and a perforated leaf quad inside a subtree that is
    referenced twice (the alpha test inside instances)."""
    os.makedirs(directory, exist_ok=True)
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height});\n\n'
    t += cornell_box_shell()
    t += _alpha_quad("screen", "CBNode", [(-1.1, 0.2, -0.6), (1.1, 0.2, -0.6), (1.1, 2.1, -0.6), (-1.1, 2.1, -0.6)], (0, 0, 1), (1, 0, 0), 3.5,
                     _matte(0.2, 0.7, 0.3), 'FloatTexture("checker board", (0.0, 1.0))')
    t += "leafNode = createNode();\nsetTransform(leafNode, translate(0, 0, 0));\n"
    t += _alpha_quad("leaf", "leafNode", [(-0.5, 0, 0.5), (0.5, 0, 0.5), (0.5, 0, -0.5), (-0.5, 0, -0.5)], (0, 1, 0), (1, 0, 0), 2.5,
                     _matte(0.8, 0.6, 0.1), 'FloatTexture("checker board", (1.0, 0.0))')
    t += "leafRef = createReferenceNode(leafNode);\n"
    for i, (x, y, z, rot, sc) in enumerate([(-0.6, 0.9, 0.9, 0.5, 1.0), (0.7, 1.5, 0.4, -0.8, 0.7)]):
        t += (f"l{i} = createNode();\naddChild(l{i}, leafRef);\nsetTransform(l{i}, translate({x}, {y}, {z}) * rotateZ({rot}) * scale({sc}));\n"
              f"addChild(root, l{i});\n")
    t += CORNELL_CAMERA
    path = os.path.join(directory, "Cutout.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


def _tex_quad(name, parent, verts, normal, tangent, uv_scale, mat_setup, group):
    """A quad whose material group is written out by the caller (`group` = the tuple after the material: normal / alpha maps)."""
    uv = [(0, 0), (uv_scale, 0), (uv_scale, uv_scale), (0, uv_scale)]
    vs = ",\n".join(f"    (({v[0]}, {v[1]}, {v[2]}), ({normal[0]}, {normal[1]}, {normal[2]}), "
                    f"({tangent[0]}, {tangent[1]}, {tangent[2]}), ({uv[i][0]}, {uv[i][1]}))" for i, v in enumerate(verts))
    lines = list(mat_setup) + [f'{name} = createMesh(\n  (\n{vs}\n  ),\n  (\n    (surfMat, {group}((0, 1, 2), (0, 2, 3))),\n  )\n);',
                               f"addChild({parent}, {name});"]
    return "\n".join(lines) + "\n"


def write_textured(directory, width=128, height=128, spp=64):
    """Image textures ON SURFACES (Textures/image_textures.cpp:13-79,136-209): a PNG colour texture on a matte panel, an EXR
    (half float) colour texture on another, a PNG normal map on a Ward panel, and a PNG whose alpha channel cuts holes into
    a third (ImageStoreMode AsIs / NormalTexture / AlphaTexture, Image.h:121-335). Image2D takes paths as written, i.e.
    relative to the working directory (API.cpp:466): the scene file names the images by absolute path."""
    from PIL import Image as PILImage
    os.makedirs(os.path.join(directory, "images"), exist_ok=True)
    img = os.path.abspath(os.path.join(directory, "images"))
    yy, xx = np.mgrid[0:64, 0:64]
    colour = np.stack([(xx * 4) % 256, (yy * 4) % 256, ((xx // 8 + yy // 8) % 2) * 200 + 30], -1).astype(np.uint8)
    PILImage.fromarray(colour, "RGB").save(os.path.join(img, "colour.png"))
    # bumps: a grid of domes, encoded as (n * 0.5 + 0.5) * 255 (the loader's gamma table is applied by both sides)
    u, v = (xx % 16 - 7.5) / 8.0, (yy % 16 - 7.5) / 8.0
    r2 = np.minimum(u * u + v * v, 0.8)
    n = np.stack([-u, -v, np.sqrt(1.0 - r2) + 0.6], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    PILImage.fromarray(np.clip((n * 0.5 + 0.5) * 255 + 0.5, 0, 255).astype(np.uint8), "RGB").save(os.path.join(img, "normal.png"))
    holes = np.full((64, 64, 4), 255, np.uint8)
    holes[..., :3] = (90, 160, 220)
    holes[((xx % 16 - 8) ** 2 + (yy % 16 - 8) ** 2) < 25, 3] = 0
    PILImage.fromarray(holes, "RGBA").save(os.path.join(img, "holes.png"))
    rgba = np.zeros((32, 32, 4), np.float32)
    gy, gx = np.mgrid[0:32, 0:32]
    rgba[..., 0] = 0.15 + 0.7 * gx / 31.0; rgba[..., 1] = 0.15 + 0.7 * gy / 31.0; rgba[..., 2] = 0.3; rgba[..., 3] = 1.0
    capi.write_exr(os.path.join(img, "ramp.exr"), rgba)

    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height});\n\n'
    t += cornell_box_shell()
    t += _tex_quad("pngPanel", "CBNode", [(-1.3, 0.3, -2.2), (-0.2, 0.3, -2.2), (-0.2, 1.4, -2.2), (-1.3, 1.4, -2.2)], (0, 0, 1), (1, 0, 0), 1.0,
                   [f'diffuseTex = SpectrumTexture(Image2D("{img}/colour.png"));', 'surfMat = createSurfaceMaterial("matte", (diffuseTex,));'], "")
    t += _tex_quad("exrPanel", "CBNode", [(0.2, 0.3, -2.2), (1.3, 0.3, -2.2), (1.3, 1.4, -2.2), (0.2, 1.4, -2.2)], (0, 0, 1), (1, 0, 0), 2.0,
                   [f'diffuseTex = SpectrumTexture(Image2D("{img}/ramp.exr"));', 'surfMat = createSurfaceMaterial("matte", (diffuseTex,));'], "")
    t += _tex_quad("bumpPanel", "CBNode", [(-1.0, 0.02, 1.0), (1.0, 0.02, 1.0), (1.0, 0.02, -1.0), (-1.0, 0.02, -1.0)], (0, 1, 0), (1, 0, 0), 2.0,
                   ['wardTex = SpectrumTexture(Spectrum(0.6, 0.55, 0.5));',
                    'surfMat = createSurfaceMaterial("Ward", (wardTex, FloatTexture(0.15), FloatTexture(0.15)));',
                    f'bumpTex = NormalTexture(Image2D("{img}/normal.png", "Normal"));'], '"normal": bumpTex, ')
    t += _tex_quad("holePanel", "CBNode", [(-0.9, 1.2, 0.2), (0.9, 1.2, 0.2), (0.9, 2.2, -0.6), (-0.9, 2.2, -0.6)], (0, 0.6247, 0.7809), (1, 0, 0), 1.5,
                   [f'holeCol = SpectrumTexture(Image2D("{img}/holes.png"));', 'surfMat = createSurfaceMaterial("matte", (holeCol,));',
                    f'holeAlpha = FloatTexture(Image2D("{img}/holes.png", "Alpha"));'], '"alpha": holeAlpha, ')
    t += CORNELL_CAMERA
    path = os.path.join(directory, "Textured.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


def write_motion(directory, width=128, height=128, spp=64):
    """Motion blur (AnimatedTransform, libSLR/Core/Transform.h:89-144): the Cornell box over the shutter interval [0, 1]
    with a matte ball that flies and turns, an emitting panel that slides (a moving light: the transform is sampled at the
    ray's time in light sampling too, SurfaceObject.cpp:351-364) and a camera that dollies sideways."""
    write_sphere_asset(directory, 32, 16)
    t = f'setRenderer("method": "PT", ("samples": {spp},));\nsetRenderSettings("width": {width}, "height": {height}, "timeStart": 0.0, "timeEnd": 1.0);\n\n'
    t += cornell_box_shell()
    t += """function ballMat(name, attrs) {
    return createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.2, 0.6, 0.8)),));
}
ball = load3DModel("models/sphere.assbin", ballMat);
setTransform(ball, scale(0.35));
flyer = createNode();
addChild(flyer, ball);
setTransform(flyer, AnimatedTransform(translate(-0.9, 0.6, -0.6), translate(0.7, 1.3, -0.2) * rotateY(1.6) * scale(1.3), 0.0, 1.0));
addChild(root, flyer);

slider = createNode();
"""
    t += _quad("panel", [(-0.25, 0, -0.25), (0.25, 0, -0.25), (0.25, 0, 0.25), (-0.25, 0, 0.25)], (0, 1, 0), (1, 0, 0),
               ['scatterMat = createSurfaceMaterial("matte", (SpectrumTexture(Spectrum(0.8, 0.8, 0.8)),));',
                'emitterMat = createEmitterSurfaceProperty("diffuse", (SpectrumTexture(Spectrum("ID": "D65") * 3),));',
                'surfMat = createSurfaceMaterial("emitter", (scatterMat, emitterMat));']).replace("CBNode", "slider")
    t += _quad("panelBack", [(-0.3, -0.02, 0.3), (0.3, -0.02, 0.3), (0.3, -0.02, -0.3), (-0.3, -0.02, -0.3)], (0, -1, 0), (1, 0, 0),
               _matte(0.4, 0.4, 0.4)).replace("CBNode", "slider")
    t += """setTransform(slider, AnimatedTransform(translate(0.9, 0.3, 0.9), translate(-0.2, 0.5, 1.2) * rotateZ(0.5), 0.0, 1.0));
addChild(root, slider);

dolly = createNode();
cameraNode = createNode();
camera = createPerspectiveCamera("aspect": 4.0 / 3.0, "fovY": 0.4807705238, "radius": 0.025, "imgDist": 1.0, "objDist": 6.3);
addChild(cameraNode, camera);
setTransform(cameraNode, translate(0.0, 1.689714, 6.70284) * rotateY(3.1415926536) * rotateX(0.0563936));
addChild(dolly, cameraNode);
setTransform(dolly, AnimatedTransform(translate(-0.15, 0, 0), translate(0.15, 0.05, 0), 0.0, 1.0));
addChild(root, dolly);
"""
    path = os.path.join(directory, "Motion.txt")
    with open(path, "w") as f:
        f.write(t)
    return path


SCENES = {
    "motion": write_motion,
    "textured": write_textured,
    "cutout": write_cutout,
    "diffuse": write_cornell_diffuse,
    "spheres": write_cornell_spheres,
    "materials": write_cornell_materials,
    "ibl": lambda d, width=1024, height=1024, spp=256: write_ibl_test(d, width, height, spp, env_size=(512, 256)),
    "ibl_full": write_ibl_test,
    "instanced": _small_instanced,
    "scatter": write_scatter,
    "lamps": write_lamps,
    "nested": write_nested,
    "instanced_full": write_instanced,
    "instanced_10m": lambda d, width=1920, height=1080, spp=1024: write_instanced(d, width, height, spp, base_segments=(318, 159)),
}
