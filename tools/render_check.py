"""Render a scene with the reference (oracle/_ref/ref_render, CPU) and with the GPU path and print
image statistics: relative RMSE in linear sRGB against the reference, the reference's own two-seed
noise floor and the ratio of image means (bias). Development tool -- the committed parity tests are
in tests/test_gpu_render.py.

  python tools/render_check.py [scene ...] [--size N] [--spp N] [--out DIR]
scene: diffuse | spheres | a path to a scene file
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from slr_b200 import capi, scenes  # noqa: E402
import render_util as ru  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scenes", nargs="*", default=["diffuse", "spheres", "materials", "ibl", "instanced"])
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "render_check"))
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    work = tempfile.mkdtemp(prefix="slr_rc_")
    W = a.width or a.size
    H = a.height or a.size
    for name in a.scenes:
        path = ru.scene_file(name, work, W, H, a.spp)
        tag = os.path.splitext(os.path.basename(path))[0]
        res = {"scene": tag, "width": W, "height": H, "spp": a.spp}
        t0 = time.time()
        hs = capi.read_scene(path)
        res["host_read_build_s"] = round(time.time() - t0, 3)
        accum, st = capi.host_render(hs, W, H, a.spp)
        res["triangles"] = int(hs.desc.num_triangles); res["instances"] = int(hs.desc.num_instances); res["qbvh_nodes"] = int(hs.desc.num_bvh_nodes)
        res["gpu"] = st
        res["gpu_mpaths_s"] = st["paths"] / max(st["device_s"], 1e-9) / 1e6
        res["rays_per_path"] = st["rays"] / max(st["paths"], 1)
        gpu_rgb = capi.accum_to_rgb(accum, 1.0 / a.spp)
        res["gpu_mean_rgb"] = gpu_rgb.reshape(-1, 3).mean(0).tolist()
        res["gpu_nonfinite"] = int((~np.isfinite(accum)).sum())
        capi.save_bmp(os.path.join(a.out, f"{tag}_gpu.bmp"), accum, 1.0 / a.spp, 509.295807)
        np.save(os.path.join(a.out, f"{tag}_gpu_rgb.npy"), gpu_rgb.astype(np.float16))
        if not a.no_ref and ru.have_ref_render():
            ref1, j1 = ru.run_ref_render(path, a.spp, W, H, seed=1509761209)
            ref2, j2 = ru.run_ref_render(path, a.spp, W, H, seed=777)
            r1 = capi.accum_to_rgb(ref1, 1.0 / a.spp)
            r2 = capi.accum_to_rgb(ref2, 1.0 / a.spp)
            capi.save_bmp(os.path.join(a.out, f"{tag}_ref.bmp"), ref1, 1.0 / a.spp, 509.295807)
            (r1, gpu_rgb, r2), d1 = ru.sanitize_reference(r1, gpu_rgb, r2)
            (r2, gpu_rgb, r1), d2 = ru.sanitize_reference(r2, gpu_rgb, r1)
            res["reference_nan_pixels"] = d1 + d2
            res["ref"] = j1
            res["ref_mean_rgb"] = r1.reshape(-1, 3).mean(0).tolist()
            res["noise_floor_relrmse"] = ru.rel_rmse(r2, r1, trim=0.005)
            res["gpu_relrmse"] = ru.rel_rmse(gpu_rgb, r1, trim=0.005)
            res["gpu_relrmse_vs_ref2"] = ru.rel_rmse(gpu_rgb, r2, trim=0.005)
            res["speedup_device_vs_ref"] = res["gpu_mpaths_s"] / j1["mpaths_per_s"]
            res["speedup_wall_vs_ref"] = (st["paths"] / st["wall_s"] / 1e6) / j1["mpaths_per_s"]
            res["mean_ratio_rgb"] = (gpu_rgb.reshape(-1, 3).mean(0) / r1.reshape(-1, 3).mean(0)).tolist()
            res["blocks_relrmse_gpu"] = ru.block_rel_rmse(gpu_rgb, r1, 16, trim=0.03)
            res["blocks_relrmse_floor"] = ru.block_rel_rmse(r2, r1, 16, trim=0.03)
        print(json.dumps(res))
        with open(os.path.join(a.out, f"{tag}.json"), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
