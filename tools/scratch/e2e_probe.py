import sys, time, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from slr_b200 import capi, scenes
d = tempfile.mkdtemp()
p = scenes.write_cornell_spheres(d, 512, 512, 64)
with capi.stdout_to_stderr():
    hs = capi.read_scene(p)
capi.host_render(hs, 512, 512, 64)
for k in range(3):
    t0 = time.perf_counter()
    img, st = capi.host_render(hs, 512, 512, 64)
    t1 = time.perf_counter()
    print("iter", k, "python wall %.1f ms" % (1e3 * (t1 - t0)), {a: round(1e3 * st[a], 2) for a in ("call_s", "wall_s", "upload_s", "device_s")})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); capi.host_render(hs, 512, 512, 64); pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(8)
