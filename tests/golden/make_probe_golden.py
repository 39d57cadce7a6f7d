"""Generates tests/golden/probe_<scene>.npz from the REFERENCE (oracle/_ref/ref_probe, built by oracle/Makefile
from /root/reference): for 2048 probe rays per scene the reference's Scene::intersect, getSurfacePoint,
createBSDF, BSDF::sample / evaluate / evaluatePDF and emittance outputs (layout: include/slrgpu.h,
slrgpu_probe_shading). The probe rays are stored with the results.

    python tests/golden/make_probe_golden.py [--bpt] [scene ...]

--bpt: the bidirectional path tracer's queries (ref_probe ... bpt: reverse values, adjoint on odd probes; layout:
slrgpu_probe_shading_bpt) -> probe_bpt_<scene>.npz.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402
from slr_b200 import capi  # noqa: E402

SCENES = ["diffuse", "spheres", "materials", "ibl", "instanced", "cutout", "textured", "motion", "nested"]
N, SEED = 2048, 20261018


def main():
    work = tempfile.mkdtemp(prefix="slr_probe_golden_")
    bpt = "--bpt" in sys.argv[1:]
    for name in [a for a in sys.argv[1:] if a != "--bpt"] or SCENES:
        path = ru.scene_file(name, work, 64, 64, 1)
        with capi.stdout_to_stderr():
            hs = capi.read_scene(path)
        center = [hs.desc.world_center[i] for i in range(3)]
        probes = ru.make_probes(center, hs.desc.world_radius, N, SEED)
        want = ru.run_ref_probe(path, probes, bpt=bpt)
        out = os.path.join(ru.GOLDEN, f"probe_bpt_{name}.npz" if bpt else f"probe_{name}.npz")
        np.savez_compressed(out, probes=probes, reference=want)
        hit = want[:, 0] == 1
        types, counts = np.unique(want[hit, 22 if bpt else 32].astype(int), return_counts=True)
        print(f"{name}: {int(hit.sum())} of {N} probes hit a surface; sampled direction types {dict(zip(types.tolist(), counts.tolist()))}; wrote {out}")


if __name__ == "__main__":
    main()
