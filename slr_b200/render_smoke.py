"""smoke() part 2: one small render of the Cornell box with spheres on cuda:0 through the renderer front
end (slrhost_render -> slrgpu_render), checked against the committed golden block means that
tests/golden/make_render_golden.py produced from the reference's PathTracingRenderer, and -- when the
compiled reference is present (oracle/_ref/ref_render) -- against a fresh reference render."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import render_util as ru
    from . import capi
    g = np.load(os.path.join(ru.GOLDEN, "render_spheres.npz"))
    size, block, spp = int(g["size"]), int(g["block"]), 2048
    work = tempfile.mkdtemp(prefix="slr_smoke_")
    path = ru.scene_file("spheres", work, size, size, spp)
    hs = capi.read_scene(path)
    accum, st = capi.host_render(hs, size, size, spp)
    assert np.isfinite(accum).all(), "non-finite values in the frame buffer"
    rgb = capi.accum_to_rgb(accum, 1.0 / spp)
    got = ru.block_means(rgb, block)
    want = g["block_mean"]
    sig = g["block_sigma"] * np.sqrt(float(g["ref_spp"]) / spp + 1.0)
    bad = np.abs(got - want) > 6.0 * sig + 0.03 * want + 1e-6
    ratio = got.mean((0, 1)) / want.mean((0, 1))
    assert bad.mean() <= 0.02 and np.all(np.abs(ratio - 1) < 0.02), f"render differs from the golden: {bad.sum()} blocks, mean ratio {ratio}"
    msg = (f"smoke: {size}x{size}x{spp} spp Cornell_Box_Spheres, {st['paths']} paths, {st['rays']} rays, "
           f"{st['paths'] / max(st['device_s'], 1e-9) / 1e6:.1f} Mpaths/s device; block means match the reference golden (mean ratio {ratio.round(4).tolist()})")
    if ru.have_ref_render():
        ref = capi.accum_to_rgb(ru.run_ref_render(path, 64, size, size)[0], 1.0 / 64)
        r2 = ru.block_means(ref, block).mean((0, 1)) / want.mean((0, 1))
        msg += f"; fresh reference render mean ratio {r2.round(4).tolist()}"
    print(msg)
