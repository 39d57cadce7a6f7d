"""Golden output of the reference's DebugRenderer (libSLR/Renderers/DebugRenderer.cpp) for the GPU debug renderer's
parity test: oracle/_ref/ref_render <scene> ... debug writes geometric_normal.bmp, shading_normal.bmp and
shading_tangent.bmp; their pixels (RGB8, top-down) are stored in tests/golden/debug_<scene>.npz together with
floor_<channel> = the fractions of pixels within 6/255 and within 1/255 between two reference runs with different seeds.
Run here (needs /root/reference compiled into oracle/_ref):   python tests/golden/make_debug_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402

CASES = [("spheres", 128, 128), ("materials", 128, 128), ("instanced", 160, 90)]


def read_bmp(path):
    """24-bit bottom-up BMP as written by the reference's saveBMP (row padding width % 4) -> [h, w, 3] RGB uint8 top-down."""
    raw = open(path, "rb").read()
    w, h = int.from_bytes(raw[18:22], "little"), int.from_bytes(raw[22:26], "little")
    off = int.from_bytes(raw[10:14], "little")
    row = 3 * w + w % 4
    px = np.frombuffer(raw, np.uint8, count=row * h, offset=off).reshape(h, row)[:, :3 * w].reshape(h, w, 3)
    return px[::-1, :, ::-1].copy()


def run_ref_debug(scene_path, width, height, seed=0):
    scene_path = os.path.abspath(scene_path)
    d = os.path.dirname(scene_path)
    subprocess.run([ru.REF_RENDER, os.path.basename(scene_path), "debug_out.bin", "1", str(width), str(height), str(seed), "0", "debug"],
                   check=True, cwd=d, capture_output=True)
    return {k: read_bmp(os.path.join(d, k + ".bmp")) for k in ("geometric_normal", "shading_normal", "shading_tangent")}


def main():
    for name, w, h in CASES:
        with tempfile.TemporaryDirectory() as d:
            path = ru.scene_file(name, d, w, h, 1)
            imgs = run_ref_debug(path, w, h)
            # the noise floor of the comparison: the reference against itself with another seed (another jitter)
            other = run_ref_debug(path, w, h, seed=4242)
        for k in list(imgs):
            diff = np.abs(imgs[k].astype(np.int32) - other[k].astype(np.int32)).max(-1)
            imgs["floor_" + k] = np.array([(diff <= 6).mean(), (diff <= 1).mean()], np.float64)
        np.savez_compressed(os.path.join(HERE, f"debug_{name}.npz"), **imgs)
        print(name, {k: (v.shape, int(v.mean())) if v.ndim == 3 else v.round(4).tolist() for k, v in imgs.items()})


if __name__ == "__main__":
    main()
