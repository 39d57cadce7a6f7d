import sys, os, argparse
sys.path.insert(0, os.getcwd())
import numpy as np
from slr_b200 import capi, render_bench
for wl, spp in (("materials", 32), ("cornell_spheres", 16), ("ibl", 16), ("instanced", 2)):
    args = argparse.Namespace(workload=wl, size=0, spp=spp, pool=0)
    path, w, h, spp, desc = render_bench._scene(args)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    gs = capi.GpuScene(hs)
    img, st = capi.gpu_render(gs, w, h, 0, spp)
    bad = ~np.isfinite(img)
    print(wl, "non-finite values:", int(bad.sum()), "pixels:", int(bad.any(-1).sum()), "mean", float(np.nanmean(img)), "device ms", round(st["device_ms"], 2), flush=True)
