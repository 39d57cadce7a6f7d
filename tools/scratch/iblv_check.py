import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from slr_b200 import capi
d = "tools/scratch/iblv"
np.set_printoptions(linewidth=200, precision=3, suppress=True)
for k in ["base", "pinhole", "norot"]:
    ref = np.load(f"{d}/ref_{k}.npy")
    hs = capi.read_scene(f"{d}/ibl_{k}.txt")
    a, st = capi.host_render(hs, 128, 128, 4096)
    g = capi.accum_to_rgb(a, 1 / 4096)
    np.save(f"gpurun_out/iblv_gpu_{k}.npy", g.astype(np.float32))
    lum = ref.mean(-1)
    ys, xs = np.nonzero(lum > 0.003)
    print(k, "gpu/ref mean", g.mean() / ref.mean())
    for y, x in zip(ys, xs):
        print("  ", y, x, "ref", lum[y, x], "gpu", g[y, x].mean(), "ratio", g[y, x].mean() / lum[y, x])
