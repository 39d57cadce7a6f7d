"""Device frame time of a render workload against the tail-kernel limit (SLRGPU_TAIL_PATHS; 0 = every bounce is a
wave, the round-1 pipeline) and the pool size, with a per-wave timeline (SLRGPU_WAVE_LOG) of one frame per setting
and the image difference against the setting without a tail kernel. Run under gpurun:
    python tools/tail_sweep.py [workload] [frames] > gpurun_out/tail_sweep.txt
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from slr_b200 import capi  # noqa: E402
from slr_b200 import render_bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="cornell_spheres")
    ap.add_argument("frames", nargs="?", type=int, default=5)
    ap.add_argument("--tails", default="0,2048,8192,37888")
    ap.add_argument("--pools", default="0")
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    a = ap.parse_args()
    args = argparse.Namespace(workload=a.workload, size=a.size, spp=a.spp, pool=0)
    path, w, h, spp, desc = render_bench._scene(args)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    gs = capi.GpuScene(hs)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    print(f"# {desc}: {w}x{h}, {spp} spp, {w * h * spp / 1e6:.1f} M paths per frame")
    ref_img = None
    for pool in [int(x) for x in a.pools.split(",")]:
        for tail in [int(x) for x in a.tails.split(",")]:
            os.environ["SLRGPU_TAIL_PATHS"] = str(tail)
            os.environ.pop("SLRGPU_WAVE_LOG", None)
            capi.gpu_render(gs, w, h, 0, spp, pool_size=pool)            # warm-up (allocates the pool)
            ms = []
            for _ in range(a.frames):
                img, st = capi.gpu_render(gs, w, h, 0, spp, pool_size=pool)
                ms.append(st["device_ms"])
            log = os.path.join(out_dir, f"wavelog_{a.workload}_pool{pool}_tail{tail}.csv")
            if os.path.exists(log):
                os.remove(log)
            os.environ["SLRGPU_WAVE_LOG"] = log
            capi.gpu_render(gs, w, h, 0, spp, pool_size=pool)
            os.environ.pop("SLRGPU_WAVE_LOG", None)
            _, stp = capi.gpu_render(gs, w, h, 0, spp, pool_size=pool, flags=capi.RENDER_PROFILE_STAGES)
            if ref_img is None:
                ref_img = img
            num = float(np.sqrt(np.mean((img - ref_img) ** 2)))
            den = float(np.sqrt(np.mean(ref_img ** 2)))
            med = float(np.median(ms))
            print(f"pool {pool:9d} tail {tail:6d}: device {med:7.2f} ms (min {min(ms):.2f} max {max(ms):.2f}) = "
                  f"{w * h * spp / med / 1e3:7.1f} Mpaths/s | waves {st['waves']} tail paths {st['tail_paths']} bounces {st['tail_waves']} "
                  f"rays {st['rays']} launches {st['kernel_launches']} | rms diff vs first {num / den:.2e} | profiled: "
                  + " ".join(f"{k[:-3]} {stp[k]:.2f}" for k in ("raygen_ms", "extend_ms", "surface_ms", "material_ms", "shadow_ms", "tail_ms", "other_ms")),
                  flush=True)


if __name__ == "__main__":
    main()
