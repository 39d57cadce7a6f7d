"""Helpers of the image-parity tests: run the compiled reference renderer (oracle/_ref/ref_render,
TEST INFRASTRUCTURE ONLY), write the benchmark scenes, compare images."""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RENDER = os.path.join(ROOT, "oracle", "_ref", "ref_render")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def have_ref_render():
    return os.access(REF_RENDER, os.X_OK)


def run_ref_render(scene_path, spp, width, height, seed=0, qbvh=0, timeout=3600):
    """Returns (accum[h, w, 16] float32, timing dict) from the reference's PathTracingRenderer."""
    scene_path = os.path.abspath(scene_path)
    out = scene_path + f".ref_{spp}_{width}x{height}_{seed}.bin"
    # the reference resolves asset paths as <cwd>/<dirname(scene)>/<asset>: run it on the bare file name
    p = subprocess.run([REF_RENDER, os.path.basename(scene_path), out, str(spp), str(width), str(height), str(seed), str(qbvh)],
                       capture_output=True, text=True, timeout=timeout, cwd=os.path.dirname(scene_path))
    if p.returncode != 0:
        raise RuntimeError(f"ref_render failed: {p.stderr[-2000:]}")
    timing = {}
    for line in p.stderr.splitlines():
        if line.startswith("{"):
            timing = json.loads(line.replace(': inf', ': Infinity').replace(': nan', ': NaN'))
    with open(out, "rb") as f:
        w, h, c = np.frombuffer(f.read(12), np.uint32)
        accum = np.frombuffer(f.read(), np.float32).reshape(h, w, c).copy()
    os.remove(out)
    return accum, timing


def scene_file(name, directory, width, height, spp):
    from slr_b200 import scenes
    if os.path.exists(name):
        return name
    return scenes.SCENES[name](directory, width=width, height=height, spp=spp)


def rel_rmse(img, ref, trim=0.0):
    """sqrt(mean((img - ref)^2)) / mean(ref) over all pixels and channels (linear sRGB floats). With
    trim > 0 the largest `trim` fraction of the squared errors is left out: a few pixels that see a
    tiny very bright feature (a mirrored sun, caustic fireflies) have a heavy-tailed error that would
    otherwise decide the whole statistic -- both for the image under test and for the noise floor."""
    img = np.asarray(img, np.float64)
    ref = np.asarray(ref, np.float64)
    ok = np.isfinite(ref)              # the reference itself occasionally writes a NaN pixel (see sanitize_reference)
    err = ((img - ref) ** 2)[ok]
    ref = ref[ok]
    if trim > 0.0:
        keep = err.size - int(np.ceil(trim * err.size))
        err = np.partition(err, keep - 1)[:keep]
    return float(np.sqrt(np.mean(err)) / np.mean(ref))


def sanitize_reference(ref, *others):
    """The reference's PathTracingRenderer now and then accumulates a NaN into a pixel (its BSDF asserts
    are compiled out with NDEBUG; seen with the Ward / Ashikhmin materials at grazing angles, about one
    pixel per 10^7 paths). Such pixels carry no information: they are set to zero in the reference image
    AND in every image compared with it. Returns the cleaned copies and the number of pixels dropped."""
    bad = ~np.isfinite(ref).all(-1)
    out = []
    for img in (ref,) + others:
        img = np.array(img, copy=True)
        img[bad] = 0.0
        out.append(img)
    return out, int(bad.sum())


def block_means(img, block):
    h, w, c = img.shape
    hb, wb = h // block, w // block
    return np.asarray(img[:hb * block, :wb * block], np.float64).reshape(hb, block, wb, block, c).mean((1, 3))


def block_rel_rmse(img, ref, block, trim=0.0):
    """rel_rmse of block x block averages: Monte-Carlo noise shrinks by `block`, systematic
    differences (a wrong BSDF, a missing light path) do not."""
    return rel_rmse(block_means(img, block), block_means(ref, block), trim)
