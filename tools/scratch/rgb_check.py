import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, render_util as ru
from slr_b200 import capi
np.set_printoptions(linewidth=200, precision=5, suppress=True)
p = ru.scene_file("diffuse", "/tmp/rgbc", 64, 64, 64)
for mode in [False, True]:
    hs = capi.read_scene(p, rgb_mode=mode)
    gs = capi.GpuScene(hs)
    for mpl in [1, 2, 100]:
        a, st = capi.gpu_render(gs, 64, 64, 0, 64, max_path_length=mpl)
        print("rgb" if mode else "spec", "maxlen", mpl, a.shape, "mean per channel", a.reshape(-1, a.shape[-1]).mean(0) / 64, "rays", st["rays"])
        if mode:
            print("   light px", a[6, 32] / 64, " floor px", a[55, 32] / 64, " left wall", a[32, 6] / 64)
        else:
            rgb = capi.accum_to_rgb(a, 1 / 64)
            print("   light px", rgb[6, 32], " floor px", rgb[55, 32], " left wall", rgb[32, 6], "mean", rgb.reshape(-1, 3).mean(0))
