"""Motion blur, host side: the host library's AnimatedTransform -- polar decomposition of the key
matrices, lerp / Slerp / lerp sampling, motion bounds (libSLR/Core/Transform.h:89-144, BasicTypes/Quaternion.cpp:15-43) --
against the reference's own classes (oracle/_ref/ref_motion) and against a committed golden of the same outputs; and the
scene language's AnimatedTransform builtin reaching the flattened scene as a moving instance / a moving camera."""
import os
import struct
import subprocess

import numpy as np
import pytest

import render_util as ru
from slr_b200 import capi, synth

REF_MOTION = os.path.join(ru.ROOT, "oracle", "_ref", "ref_motion")
GOLDEN = os.path.join(ru.GOLDEN, "motion_transforms.npz")


def cases():
    rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0, 0], [np.sin(a), np.cos(a), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    out = []
    out.append(("translate", synth.translate(0.0, 0.5, 0.0), synth.translate(1.5, 0.7, -0.4)))
    out.append(("spin", synth.translate(0.2, 1.0, 0.0) @ synth.rotate_y(0.1) @ synth.scale(0.5), synth.translate(0.8, 1.2, 0.3) @ synth.rotate_y(1.9) @ synth.scale(0.8)))
    out.append(("big_rotation", synth.rotate_y(0.2), synth.rotate_y(2.9) @ rz(0.7).astype(np.float32)))
    out.append(("nonuniform", synth.scale(1.0, 2.0, 0.5), synth.translate(1, 0, 0) @ rz(0.4).astype(np.float32) @ synth.scale(2.0, 0.7, 1.3)))
    return [(n, np.asarray(a, np.float32), np.asarray(b, np.float32)) for n, a, b in out]


TIMES = np.array([-0.5, 0.0, 0.1, 0.25, 0.5, 0.77, 0.999, 1.0, 1.7], np.float32)
BOX = np.array([-0.5, -0.2, -1.0, 0.7, 0.9, 0.4], np.float32)
T_BEGIN, T_END = 0.0, 1.0


def run_reference(mb, me):
    d = os.path.dirname(GOLDEN)
    pin, pout = os.path.join("/tmp", f"motion_{os.getpid()}.in"), os.path.join("/tmp", f"motion_{os.getpid()}.out")
    with open(pin, "wb") as f:
        f.write(np.ascontiguousarray(mb.T, np.float32).tobytes() + np.ascontiguousarray(me.T, np.float32).tobytes())
        f.write(struct.pack("<ff", T_BEGIN, T_END) + BOX.tobytes() + struct.pack("<I", len(TIMES)) + TIMES.tobytes())
    subprocess.run([REF_MOTION, pin, pout], check=True)
    a = np.fromfile(pout, np.float32)
    os.remove(pin); os.remove(pout)
    return a[:46], a[46:52], a[52:].reshape(len(TIMES), 32)


def check(mb, me, dec_w, bounds_w, sampled_w):
    dec, bounds, sampled = capi.sample_animated(mb, me, T_BEGIN, T_END, BOX, TIMES)
    np.testing.assert_allclose(dec, dec_w, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(sampled, sampled_w, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(bounds, bounds_w, rtol=2e-5, atol=2e-6)
    # outside the key times the key frames are returned bit for bit (Transform.h:105-112)
    assert np.array_equal(sampled[0, :16], np.ascontiguousarray(mb.T).reshape(-1)) and np.array_equal(sampled[-1, :16], np.ascontiguousarray(me.T).reshape(-1))


def test_animated_transform_matches_golden():
    g = np.load(GOLDEN)
    for name, mb, me in cases():
        check(mb, me, g[f"{name}_dec"], g[f"{name}_bounds"], g[f"{name}_sampled"])


@pytest.mark.skipif(not os.access(REF_MOTION, os.X_OK), reason="oracle/_ref/ref_motion not built")
def test_animated_transform_matches_live_reference():
    for name, mb, me in cases():
        check(mb, me, *run_reference(mb, me))


def test_scene_language_animated_node_becomes_a_moving_instance(tmp_path):
    path = ru.scene_file("motion", str(tmp_path), 32, 32, 1)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    d = hs.desc
    assert d.num_motions >= 2 and d.camera_motion != 0
    moving = [d.instances[i] for i in range(d.num_instances) if d.instances[i].motion]
    assert len(moving) >= 1
    m = d.motions[moving[0].motion - 1]
    assert (m.t_begin, m.t_end) == (0.0, 1.0)
    assert list(m.mat_end) != list(moving[0].mat)
    assert hs.context["timeStart"] == 0.0 and hs.context["timeEnd"] == 1.0


if __name__ == "__main__":       # python tests/test_motion_host.py: regenerates the golden from the reference
    out = {}
    for name, mb, me in cases():
        out[f"{name}_dec"], out[f"{name}_bounds"], out[f"{name}_sampled"] = run_reference(mb, me)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN)
