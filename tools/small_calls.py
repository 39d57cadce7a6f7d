"""A few small calls through every CUDA entry point (render with a tiny pool = many waves + the tail kernel, default
render, debug renderer, ray batch with counters) on small scenes: a quick manual check after a kernel change.
(compute-sanitizer is closed on this GPU pool, so this is not a memcheck run.)
    python tools/small_calls.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from slr_b200 import capi, scenes, synth  # noqa: E402
import oracle_util as ou  # noqa: E402

d = tempfile.mkdtemp(prefix="slr_sanitize_")
for name, size, spp in (("spheres", 48, 4), ("materials", 40, 2), ("instanced", 48, 2), ("ibl", 40, 2)):
    if name not in scenes.SCENES:
        continue
    path = scenes.SCENES[name](d, width=size, height=size, spp=spp)
    hs = capi.read_scene(path)
    gs = capi.GpuScene(hs)
    img, st = capi.gpu_render(gs, size, size, 0, spp, pool_size=4096)          # small pool: many waves + the tail kernel
    img2, _ = capi.gpu_render(gs, size, size, 0, spp)
    out, _ = capi.host_render_debug(hs, size, size)
    print(name, "render ok", float(img.mean()), float(img2.mean()), "tail paths", st["tail_paths"], "debug hits", float(out[:, :, 0].mean()), flush=True)
pos, idx = synth.heightfield(32)
hs = ou.build_host_scene([(pos, idx)], [(0, 0, None)])
gs = capi.GpuScene(hs)
rays = synth.random_rays(20000, pos.min(0), pos.max(0), seed=5)
r = gs.intersect(rays, counters=True)
print("intersect ok", int((r["prim"] != 0xFFFFFFFF).sum()), flush=True)
