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


SCENES = {
    "diffuse": write_cornell_diffuse,
    "spheres": write_cornell_spheres,
}
