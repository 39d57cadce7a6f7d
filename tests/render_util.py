"""Helpers of the image-parity tests: run the compiled reference renderer (oracle/_ref/ref_render,
TEST INFRASTRUCTURE ONLY), write the benchmark scenes, compare images."""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RENDER = os.path.join(ROOT, "oracle", "_ref", "ref_render")
REF_RENDER_RGB = os.path.join(ROOT, "oracle", "_ref", "ref_render_rgb")      # the reference built without Use_Spectral_Representation
GOLDEN = os.path.join(ROOT, "tests", "golden")


def have_ref_render(rgb=False):
    return os.access(REF_RENDER_RGB if rgb else REF_RENDER, os.X_OK)


def run_ref_render(scene_path, spp, width, height, seed=0, qbvh=0, timeout=3600, rgb=False, bpt=False):
    """Returns (accum[h, w, 16] float32, timing dict) from the reference's PathTracingRenderer (bpt=True: its
    BidirectionalPathTracingRenderer, separated light-tracing buffers added in); rgb=True runs the reference's RGB-mode
    build (accum[h, w, 3])."""
    scene_path = os.path.abspath(scene_path)
    out = scene_path + f".ref_{spp}_{width}x{height}_{seed}.bin"
    # the reference resolves asset paths as <cwd>/<dirname(scene)>/<asset>: run it on the bare file name
    p = subprocess.run([REF_RENDER_RGB if rgb else REF_RENDER, os.path.basename(scene_path), out, str(spp), str(width), str(height), str(seed), str(qbvh)]
                       + (["bpt"] if bpt else []), capture_output=True, text=True, timeout=timeout, cwd=os.path.dirname(scene_path))
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


def run_ref_render_with_bmps(scene_path, spp, width, height, seed=0):
    """run_ref_render plus the bytes of the NNN.bmp files the reference's renderer exported on the way (in order)."""
    import glob
    d = os.path.dirname(os.path.abspath(scene_path))
    for f in glob.glob(os.path.join(d, "[0-9][0-9][0-9].bmp")):
        os.remove(f)
    accum, timing = run_ref_render(scene_path, spp, width, height, seed=seed)
    bmps = []
    for f in sorted(glob.glob(os.path.join(d, "[0-9][0-9][0-9].bmp"))):
        with open(f, "rb") as fh:
            bmps.append(fh.read())
        os.remove(f)
    return accum, timing, bmps


def bmp_pixels(data, width, height):
    """(header bytes, pixel bytes [h, w, 3]) of a 24-bit BMP written by saveBMP (Helper/bmp_exporter.cpp:14-53): rows are
    3 * width + width % 4 bytes; the padding bytes are uninitialised heap in the reference and are left out."""
    data = np.frombuffer(data, np.uint8)
    row = 3 * width + width % 4
    assert data.size == 54 + row * height
    return data[:54].copy(), data[54:].reshape(height, row)[:, :3 * width].reshape(height, width, 3).copy()


REF_SCENES = os.path.join(ROOT, "oracle", "_ref", "TestScenes")     # the reference's shipped scene files (oracle/Makefile `scenes`)


def reference_scene_file(name, directory, width, height, spp, method="PT"):
    """The reference's TestScenes/<name>, unchanged, with the size / sample-count override appended and synthetic assets
    written next to it (slr_b200.scenes.write_reference_scene). None when the file did not travel to this machine."""
    from slr_b200 import scenes
    src = os.path.join(REF_SCENES, name)
    if not os.path.exists(src):
        return None
    return scenes.write_reference_scene(src, directory, width, height, spp, method=method)


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


# --------------------------------------------------------------------------------------------------
# shading probes (function-level parity of surface points, materials, textures, spectra and BSDFs)
# --------------------------------------------------------------------------------------------------
REF_PROBE = os.path.join(ROOT, "oracle", "_ref", "ref_probe")


def have_ref_probe():
    return os.access(REF_PROBE, os.X_OK)


def make_probes(center, radius, n, seed):
    """n probe rays from inside the scene's bounding sphere in random directions, with the random numbers a
    path would draw (wavelength offset, wavelength selection, BSDF component, BSDF direction) and a random
    world direction to evaluate the BSDF for. Deterministic in (center, radius, n, seed)."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = rng.normal(size=(n, 3))
    o *= (0.55 * radius * rng.random((n, 1)) ** (1 / 3)) / np.linalg.norm(o, axis=1, keepdims=True)
    o += np.asarray(center, np.float64)
    e = rng.normal(size=(n, 3))
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    u = rng.random((n, 5)) * (1.0 - 1e-6)
    return np.concatenate([o, d, u, e], 1).astype(np.float32)


def run_ref_probe(scene_path, probes, timeout=600, bpt=False):
    scene_path = os.path.abspath(scene_path)
    pin, pout = scene_path + ".probes.bin", scene_path + ".probes.out"
    with open(pin, "wb") as f:
        f.write(np.uint32(probes.shape[0]).tobytes())
        f.write(np.ascontiguousarray(probes, np.float32).tobytes())
    p = subprocess.run([REF_PROBE, os.path.basename(scene_path), pin, pout] + (["bpt"] if bpt else []), capture_output=True, text=True, timeout=timeout,
                       cwd=os.path.dirname(scene_path))
    if p.returncode != 0:
        raise RuntimeError(f"ref_probe failed: {p.stderr[-2000:]}")
    with open(pout, "rb") as f:
        n, stride = np.frombuffer(f.read(8), np.uint32)
        out = np.frombuffer(f.read(), np.float32).reshape(n, stride).copy()
    os.remove(pin)
    os.remove(pout)
    return out


def compare_probes(got, want, rel=2e-3):
    """Returns a dict of mismatch fractions / worst errors between two probe result arrays [n, 64]."""
    res = {}
    res["status_mismatch"] = float(np.mean(got[:, 0] != want[:, 0]))
    hit = (got[:, 0] == 1) & (want[:, 0] == 1)
    g, w = got[hit].astype(np.float64), want[hit].astype(np.float64)
    res["hits"] = int(hit.sum())

    def relerr(a, b, floor):
        return np.abs(a - b) / (np.abs(b) + floor)
    res["t_worst_rel"] = float(relerr(g[:, 1], w[:, 1], 1e-6).max())
    res["frame_worst_abs"] = float(np.abs(g[:, 2:11] - w[:, 2:11]).max())
    res["nondelta_mismatch"] = float(np.mean(g[:, 11] != w[:, 11]))
    same_type = g[:, 32] == w[:, 32]
    res["sample_type_mismatch"] = float(np.mean(~same_type))
    gs, ws = g[same_type], w[same_type]
    scale = np.abs(ws[:, 12:28]).max(1, keepdims=True) + 1e-6
    bad_fs = (np.abs(gs[:, 12:28] - ws[:, 12:28]) / scale).max(1) > rel
    bad_dir = np.abs(gs[:, 28:31] - ws[:, 28:31]).max(1) > 2e-3
    bad_pdf = relerr(gs[:, 31], ws[:, 31], 1e-6) > rel
    res["sample_value_mismatch"] = float(np.mean(bad_fs | bad_dir | bad_pdf))
    scale = np.abs(w[:, 33:49]).max(1, keepdims=True) + 1e-6
    bad_ev = (np.abs(g[:, 33:49] - w[:, 33:49]) / scale).max(1) > rel
    bad_evpdf = relerr(g[:, 49], w[:, 49], 1e-6) > rel
    res["eval_mismatch"] = float(np.mean(bad_ev | bad_evpdf))
    res["emitting_mismatch"] = float(np.mean(g[:, 50] != w[:, 50]))
    em = w[:, 50] == 1
    res["emittance_worst_rel"] = float(relerr(g[em, 51:64], w[em, 51:64], 1e-6).max()) if em.any() else 0.0
    return res


def compare_bpt_probes(got, want, rel=2e-3):
    """Mismatch fractions between two bidirectional probe result arrays [n, 64] (include/slrgpu.h slrgpu_probe_shading_bpt):
    sampled value / direction / pdf, the reverse value and pdf of the sample, evaluatePDF with its reverse pdf, evaluate."""
    res = {"status_mismatch": float(np.mean(got[:, 0] != want[:, 0]))}
    hit = (got[:, 0] == 1) & (want[:, 0] == 1)
    g, w = got[hit].astype(np.float64), want[hit].astype(np.float64)
    res["hits"] = int(hit.sum())

    def relerr(a, b, floor):
        return np.abs(a - b) / (np.abs(b) + floor)

    def spec_bad(a, b):
        scale = np.abs(b).max(1, keepdims=True) + 1e-6
        return (np.abs(a - b) / scale).max(1) > rel
    same_type = g[:, 22] == w[:, 22]
    res["sample_type_mismatch"] = float(np.mean(~same_type))
    gs, ws = g[same_type], w[same_type]
    bad = spec_bad(gs[:, 2:18], ws[:, 2:18]) | (np.abs(gs[:, 18:21] - ws[:, 18:21]).max(1) > 2e-3) | (relerr(gs[:, 21], ws[:, 21], 1e-6) > rel)
    res["sample_value_mismatch"] = float(np.mean(bad))
    res["sample_reverse_fs_mismatch"] = float(np.mean(spec_bad(gs[:, 23:39], ws[:, 23:39])))
    res["sample_reverse_pdf_mismatch"] = float(np.mean(relerr(gs[:, 39], ws[:, 39], 1e-5) > rel))
    res["pdf_mismatch"] = float(np.mean(relerr(g[:, 40], w[:, 40], 1e-6) > rel))
    res["reverse_pdf_mismatch"] = float(np.mean(relerr(g[:, 41], w[:, 41], 1e-5) > rel))
    res["eval_mismatch"] = float(np.mean(spec_bad(g[:, 42:58], w[:, 42:58])))
    return res
