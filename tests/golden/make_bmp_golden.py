"""Generates tests/golden/save_image.npz from the REFERENCE (oracle/_ref/ref_render): the sensor's float sums after an
8-spp PathTracingRenderer run of the `materials` scene at 50 x 37 (a width that needs row padding), the camera's
sensitivity, and the bytes of every NNN.bmp the reference wrote along the way (ImageSensor::saveImage,
libSLR/Core/ImageSensor.cpp:138-186; export cadence PathTracingRenderer.cpp:63-65,83-94).

    python tests/golden/make_bmp_golden.py
"""
import glob
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402

W, H, SPP, SEED = 50, 37, 8, 7


def main():
    work = tempfile.mkdtemp(prefix="slr_bmp_golden_")
    path = ru.scene_file("materials", work, W, H, SPP)
    accum, timing, bmps = ru.run_ref_render_with_bmps(path, SPP, W, H, seed=SEED)
    out = os.path.join(ru.GOLDEN, "save_image.npz")
    np.savez_compressed(out, accum=accum, sensitivity=np.float32(timing["sensitivity"]), brightness=np.float32(4.0), spp=SPP,
                        **{f"bmp_{k:03d}": np.frombuffer(b, np.uint8) for k, b in enumerate(bmps)})
    print(f"wrote {out}: {len(bmps)} BMPs of {len(bmps[-1])} bytes")


if __name__ == "__main__":
    main()
