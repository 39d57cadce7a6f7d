"""Generates tests/golden/render_<scene>.npz from the REFERENCE renderer (oracle/_ref/ref_render, built by
oracle/Makefile from /root/reference): block means of the converged image and their standard error.

    python tests/golden/make_render_golden.py [--bpt] [scene ...]

--bpt: the reference's BidirectionalPathTracingRenderer instead (oracle/_ref/ref_render ... bpt) -> render_bpt_<scene>.npz,
the expectation of tests/test_gpu_bpt.py. It is NOT always the path tracer's: with emitters inside scaled instances (`lamps`)
the reference's bidirectional estimator converges to a darker image than its path tracer (object-space area pdfs and an
un-normalised emission direction), and the GPU twin is held to what the reference's BPT renders.

For every scene of slr_b200.scenes.SCENES used by tests/test_gpu_render.py the reference's
PathTracingRenderer renders SEEDS (8) independent images at SIZE x SIZE (64), SPP_EACH (2048) spp; the golden is the
per-block (BLOCK x BLOCK pixels, linear sRGB) mean over the seeds, `block_sigma` the standard error of
that mean estimated from the spread between seeds, `ref_spp` = SEEDS * SPP_EACH. Runs only where the
reference is available (this container); the .npz files are committed and travel to the GPU box.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402
from slr_b200 import capi  # noqa: E402

SIZE, BLOCK, SEEDS, SPP_EACH, GPU_SPP = 64, 8, 8, 2048, 16384
SCENES = ["diffuse", "spheres", "materials", "ibl", "instanced", "scatter", "lamps", "cutout", "textured", "motion", "nested"]


def main():
    args = sys.argv[1:]
    bpt = "--bpt" in args
    names = [a for a in args if a != "--bpt"] or SCENES
    work = tempfile.mkdtemp(prefix="slr_golden_")
    for name in names:
        path = ru.scene_file(name, work, SIZE, SIZE, SPP_EACH)
        means = []
        for k in range(SEEDS):
            accum, timing = ru.run_ref_render(path, SPP_EACH, SIZE, SIZE, seed=1000 + 7919 * k, bpt=bpt)
            means.append(ru.block_means(capi.accum_to_rgb(accum, 1.0 / SPP_EACH), BLOCK))
        # a block that contains a NaN pixel in one seed (the reference's own defect, see render_util.sanitize_reference)
        # is left out of that seed's contribution
        means = np.stack(means)
        count = np.isfinite(means).sum(0)
        assert count.min() >= SEEDS // 2, "too many NaN blocks in the reference renders"
        block_mean = np.nanmean(means, 0)
        block_sigma = np.nanstd(means, 0, ddof=1) / np.sqrt(count)
        out = os.path.join(ru.GOLDEN, f"render_bpt_{name}.npz" if bpt else f"render_{name}.npz")
        np.savez_compressed(out, block_mean=block_mean.astype(np.float32), block_sigma=block_sigma.astype(np.float32),
                            size=SIZE, block=BLOCK, ref_spp=SEEDS * SPP_EACH, gpu_spp=GPU_SPP,
                            reference_threads=timing.get("threads", 0))
        rel = (block_sigma / block_mean).mean()
        print(f"{name}: wrote {out}; mean relative standard error of a block mean {rel:.4f}")


if __name__ == "__main__":
    main()
