"""Prints the mismatch fractions of the bidirectional shading probe against the committed goldens (diagnostic twin of
tests/test_gpu_bpt_probe.py). python tools/bpt_probe_check.py [scene ...]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402
from slr_b200 import capi  # noqa: E402

work = tempfile.mkdtemp(prefix="bpt_probe_")
for name in sys.argv[1:] or ["diffuse", "spheres", "materials", "ibl", "instanced", "cutout", "textured", "motion"]:
    g = np.load(os.path.join(ru.GOLDEN, f"probe_bpt_{name}.npz"))
    path = ru.scene_file(name, work, 64, 64, 1)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    gs = capi.GpuScene(hs)
    got = capi.probe_shading_bpt(gs, g["probes"])
    for rel in (1e-4, 2e-3):
        print(name, rel, ru.compare_bpt_probes(got, g["reference"], rel=rel), flush=True)
    # where the reverse pdf of the sample is off: sampled type and adjoint flag of the offenders
    want = g["reference"]
    hit = (got[:, 0] == 1) & (want[:, 0] == 1)
    idx = np.nonzero(hit)[0]
    for col, label in ((39, "sample rev pdf"), (41, "pdf rev"), (40, "pdf")):
        bad = np.abs(got[idx, col] - want[idx, col]) / (np.abs(want[idx, col]) + 1e-5) > 2e-3
        if bad.any():
            b = idx[bad][:6]
            print("   ", label, "offenders:", [(int(i), int(want[i, 22]), float(got[i, col]), float(want[i, col])) for i in b], flush=True)
