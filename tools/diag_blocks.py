#!/usr/bin/env python
"""Diagnostic: per-block mean ratios GPU / reference for one scene (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru
from slr_b200 import capi

name, size, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
d = "/tmp/diag_" + name.replace(".", "_")
path = ru.reference_scene_file(name, d, size, size, spp) if name.endswith(".txt") else ru.scene_file(name, d, size, size, spp)
with capi.stdout_to_stderr():
    hs = capi.read_scene(path)
acc, st = capi.host_render(hs, size, size, spp, device=0)
gpu = capi.accum_to_rgb(acc, 1.0 / spp)
refs = [capi.accum_to_rgb(ru.run_ref_render(path, spp, size, size, seed=s)[0], 1.0 / spp) for s in (1509761209, 20240229, 777)]
ref = np.mean(refs, 0)
b = 16
bm = lambda x: ru.block_means(x, b).mean(-1)
np.set_printoptions(precision=3, linewidth=200, suppress=True)
print("gpu / mean(ref) per block"); print(bm(gpu) / bm(ref))
print("ref0 / mean(ref1, ref2) per block"); print(bm(refs[0]) / bm((refs[1] + refs[2]) / 2))
print("image mean ratio", gpu.mean((0, 1)) / ref.mean((0, 1)))
