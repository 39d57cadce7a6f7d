"""GPU bidirectional path tracing (csrc/bpt.cu behind SLRGPU_RENDER_BPT / GPUBidirectionalPathTracingRenderer) against the
reference's BidirectionalPathTracingRenderer (libSLR/Renderers/BidirectionalPathTracingRenderer.cpp:25-414).

Statistical parity only, like the path tracer's (the reference's image depends on thread scheduling). Bars, linear sRGB:
  * EXPECTATION: against the committed golden block means of the reference's converged BIDIRECTIONAL image
    (tests/golden/render_bpt_*.npz, make_render_golden.py --bpt: 8 seeds x 2048 spp of oracle/_ref/ref_render ... bpt):
    >= 98 % of the 8x8 block means within 6 standard errors + 1 %, image means within 1 %. A bidirectional estimator whose
    MIS weights do not add up to one, whose reverse pdfs are wrong or that loses a strategy shows up here as a bias.
    The reference's BPT does not always converge to its path tracer's image: with emitters inside SCALED instances (`lamps`)
    it renders 0.47 x the path tracer's brightness (object-space area pdfs, an emission direction that keeps the instance's
    scale), and `scatter` comes out 3.5 % greener. The GPU twin reproduces both -- it is held to the reference's BPT, and
    for the scenes where the two reference renderers agree also to the path tracer's golden.
  * VARIANCE: rel_rmse(gpu_bpt, ref_bpt_1) <= 1.25 x rel_rmse(ref_bpt_2, ref_bpt_1) at equal sample counts, ref_bpt_k = the
    reference's own BPT with two seeds (oracle/_ref/ref_render ... bpt): weights that add up to one but are not the
    reference's power heuristic would pass the first bar with a noisier image and fail this one.
Size-independent properties: sample-range additivity (the multi-GPU partition) and determinism.
"""
import os

import numpy as np
import pytest

import render_util as ru
from slr_b200 import capi

pytestmark = pytest.mark.gpu

BPT = 0x2      # SLRGPU_RENDER_BPT


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert capi.gpu.slrgpu_device_count() > 0, "these tests need a CUDA device"


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("bpt_scenes"))


def _bpt_rgb(path, size, spp):
    hs = capi.read_scene(path)
    accum, st = capi.host_render(hs, size, size, spp, method="BPT")
    assert np.isfinite(accum).all()
    assert st["paths"] == size * size * spp
    return capi.accum_to_rgb(accum, 1.0 / spp), st


@pytest.mark.parametrize("name", ["diffuse", "spheres", "materials", "ibl", "instanced", "scatter", "lamps", "cutout", "textured", "motion", "nested"])
def test_bpt_image_matches_golden_block_means(name, workdir):
    g = np.load(os.path.join(ru.GOLDEN, f"render_bpt_{name}.npz"))
    size, block = int(g["size"]), int(g["block"])
    spp = 4096
    path = ru.scene_file(name, workdir, size, size, spp)
    gpu, _ = _bpt_rgb(path, size, spp)
    got = ru.block_means(gpu, block)

    def within(gold):
        want, sigma = gold["block_mean"], gold["block_sigma"]
        sig = sigma * np.sqrt(float(gold["ref_spp"]) / spp + 1.0)
        err = np.abs(got - want)
        tol = 6.0 * sig + 0.01 * want + 1e-7
        bad = err > tol
        # 2 % (one block of the 64, all three channels) instead of the path tracer's 1 %: the light-tracing splats are heavy
        # tailed -- `scatter` has one block whose mean is carried by rare bright t = 1 connections in the reference too (its
        # +3.7 % of green), and 8 seeds do not pin the spread of such a block
        assert bad.mean() <= 0.02, f"{bad.sum()} of {bad.size} block means outside 6 sigma + 1 %: worst {np.max(err / tol):.2f}x"
        ratio = got.mean((0, 1)) / want.mean((0, 1))
        # 1 % + six standard errors of the image mean itself (negligible except for `scatter`, whose heavy-tailed
        # light-tracing splats leave 0.5-0.8 % of noise in the mean of a 4096-spp image)
        mean_sigma = np.sqrt((sig ** 2).sum((0, 1))) / (sig.shape[0] * sig.shape[1]) / want.mean((0, 1))
        assert np.all(np.abs(ratio - 1.0) < 0.01 + 6.0 * mean_sigma), f"image mean ratio {ratio}"
    within(g)
    if name not in ("lamps", "scatter"):      # where the reference's two renderers agree: the path tracer's golden too
        within(np.load(os.path.join(ru.GOLDEN, f"render_{name}.npz")))


@pytest.mark.parametrize("name,size,spp", [("diffuse", 96, 64), ("spheres", 128, 64), ("materials", 128, 64), ("ibl", 128, 64), ("instanced", 96, 32)])
def test_bpt_noise_matches_reference_bpt(name, size, spp, workdir):
    if not ru.have_ref_render():
        pytest.skip("oracle/_ref/ref_render not built")
    path = ru.scene_file(name, workdir, size, size, spp)
    gpu, _ = _bpt_rgb(path, size, spp)
    ref1 = capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=1509761209, bpt=True)[0], 1.0 / spp)
    ref2 = capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=20240229, bpt=True)[0], 1.0 / spp)
    (ref1, gpu, ref2), dropped1 = ru.sanitize_reference(ref1, gpu, ref2)
    (ref2, gpu, ref1), dropped2 = ru.sanitize_reference(ref2, gpu, ref1)
    assert dropped1 + dropped2 <= 4, "the reference image is mostly NaN"
    floor = ru.rel_rmse(ref2, ref1, trim=0.005)
    got = ru.rel_rmse(gpu, ref1, trim=0.005)
    assert got <= 1.25 * floor, f"relRMSE {got:.4f} vs the reference BPT's own noise floor {floor:.4f}"
    clip = float(np.percentile(ref1, 99.8))
    gpu_c, ref1_c, ref2_c = np.minimum(gpu, clip), np.minimum(ref1, clip), np.minimum(ref2, clip)
    ratio = gpu_c.reshape(-1, 3).mean(0) / ref1_c.reshape(-1, 3).mean(0)
    assert np.all(np.abs(ratio - 1.0) < 0.015), f"image mean ratio {ratio}"
    bfloor = ru.block_rel_rmse(ref2_c, ref1_c, 16, trim=0.03)
    bgot = ru.block_rel_rmse(gpu_c, ref1_c, 16, trim=0.03)
    assert bgot <= 1.5 * bfloor + 0.002, f"16x16-block relRMSE {bgot:.4f} vs floor {bfloor:.4f}"


def test_bpt_is_less_noisy_than_pt_where_it_should_be(workdir):
    """The diffuse box is lit by a small ceiling light: connections to light-subpath vertices pay off (relRMSE against the
    converged golden below the path tracer's at the same sample count)."""
    g = np.load(os.path.join(ru.GOLDEN, "render_diffuse.npz"))
    size, block = int(g["size"]), int(g["block"])
    spp = 64
    path = ru.scene_file("diffuse", workdir, size, size, spp)
    hs = capi.read_scene(path)
    pt = capi.accum_to_rgb(capi.host_render(hs, size, size, spp)[0], 1.0 / spp)
    bpt = capi.accum_to_rgb(capi.host_render(hs, size, size, spp, method="BPT")[0], 1.0 / spp)
    want = g["block_mean"]
    e_pt = ru.rel_rmse(ru.block_means(pt, block), want)
    e_bpt = ru.rel_rmse(ru.block_means(bpt, block), want)
    assert e_bpt < 0.7 * e_pt, f"BPT block relRMSE {e_bpt:.4f} vs PT {e_pt:.4f}"


def test_bpt_sample_ranges_add_up_and_rerun_is_identical(workdir):
    path = ru.scene_file("spheres", workdir, 96, 96, 16)
    hs = capi.read_scene(path)
    gs = capi.GpuScene(hs)
    whole, st = capi.gpu_render(gs, 96, 96, 0, 16, flags=BPT)
    a, _ = capi.gpu_render(gs, 96, 96, 0, 8, flags=BPT)
    b, _ = capi.gpu_render(gs, 96, 96, 8, 16, flags=BPT)
    again, st2 = capi.gpu_render(gs, 96, 96, 0, 16, flags=BPT)
    assert st["paths"] == 96 * 96 * 16 and st["rays"] == st2["rays"] and st["rays"] > st["paths"]
    # identical sample sets; only the fp32 atomic summation order differs
    scale = float(whole.mean())
    np.testing.assert_allclose(a + b, whole, rtol=5e-4, atol=1e-4 * scale)
    np.testing.assert_allclose(again, whole, rtol=5e-4, atol=1e-4 * scale)
    assert not np.array_equal(a, b)


def test_scene_file_that_selects_bpt_renders_bidirectionally(workdir):
    """setRenderer("BPT") in a scene file creates the bidirectional GPU renderer (6 of the reference's 7 TestScenes ask for
    it); the unchanged Cornell_Box_Spheres.txt through it meets the reference BPT's noise floor."""
    size, spp = 96, 32
    path = ru.reference_scene_file("Cornell_Box_Spheres.txt", os.path.join(workdir, "ref_cbs"), size, size, spp, method="BPT")
    if path is None or not ru.have_ref_render():
        pytest.skip("the reference's scene files / ref_render did not travel to this machine")
    hs = capi.read_scene(path)
    assert hs.context["method"] == "BPT"
    gpu, _ = _bpt_rgb(path, size, spp)
    ref1 = capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=1509761209, bpt=True)[0], 1.0 / spp)
    ref2 = capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=20240229, bpt=True)[0], 1.0 / spp)
    (ref1, gpu, ref2), _ = ru.sanitize_reference(ref1, gpu, ref2)
    floor = ru.rel_rmse(ref2, ref1, trim=0.005)
    got = ru.rel_rmse(gpu, ref1, trim=0.005)
    assert got <= 1.25 * floor, f"relRMSE {got:.4f} vs noise floor {floor:.4f}"
