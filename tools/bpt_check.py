"""Diagnostic run of the GPU bidirectional path tracer against the committed golden block means (and the path tracer at
the same sample count): prints per scene the image-mean ratios, the share of block means outside tolerance, device time and
rays per sample. python tools/bpt_check.py [spp] [scene ...]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402
from slr_b200 import capi  # noqa: E402


def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    names = sys.argv[2:] or ["diffuse", "spheres", "materials", "ibl", "instanced", "scatter", "lamps", "cutout", "textured", "motion"]
    work = tempfile.mkdtemp(prefix="bpt_check_")
    for name in names:
        g = np.load(os.path.join(ru.GOLDEN, f"render_{name}.npz"))
        size, block = int(g["size"]), int(g["block"])
        path = ru.scene_file(name, work, size, size, spp)
        hs = capi.read_scene(path)
        want, sigma = g["block_mean"], g["block_sigma"]
        direct = os.environ.get("BPT_DIRECT") == "1"      # slrgpu_render of the library SLRGPU_LIB names (kernel A/B runs)
        gs = capi.GpuScene(hs) if direct else None
        for method in ("PT", "BPT"):
            if direct:
                accum, st = capi.gpu_render(gs, size, size, 0, spp, flags=capi.RENDER_BPT if method == "BPT" else 0)
                st["device_s"] = st["device_ms"] * 1e-3
            else:
                accum, st = capi.host_render(hs, size, size, spp, method=method)
            rgb = capi.accum_to_rgb(accum, 1.0 / spp)
            got = ru.block_means(rgb, block)
            sig = sigma * np.sqrt(float(g["ref_spp"]) / spp + 1.0)
            err = np.abs(got - want)
            tol = 6.0 * sig + 0.01 * want + 1e-7
            ratio = got.mean((0, 1)) / want.mean((0, 1))
            print(f"{name:10s} {method:3s} spp {spp} finite {bool(np.isfinite(accum).all())} mean ratio {np.round(ratio, 4)} "
                  f"bad blocks {float((err > tol).mean()):.4f} worst {float(np.max(err / tol)):.2f} block relRMSE {ru.rel_rmse(got, want):.4f} "
                  f"device {st['device_s'] * 1e3:.1f} ms rays/path {st['rays'] / max(st['paths'], 1):.2f} "
                  f"Mpaths/s {st['paths'] / max(st['device_s'], 1e-9) / 1e6:.1f}", flush=True)


if __name__ == "__main__":
    main()
