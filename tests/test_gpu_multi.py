"""Several GPUs of one box behind the renderer API: slrgpu_render_multi partitions the frame's
sample range over scene replicas (one host thread per replica), sums the accumulation buffers onto the first replica's
device with one kernel over peer memory and downloads once; GPUPathTracingRenderer uses every visible device.
The counter-based RNG is keyed by (pixel, global sample index): the N-way frame is the 1-GPU frame to fp32 summation order.
On a 1-GPU box the replicas share the device (same threads, same exchange kernel); with >= 2 GPUs they sit on different ones.
"""
import numpy as np
import pytest

import render_util as ru
from slr_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    assert capi.gpu.slrgpu_device_count() > 0, "these tests need a CUDA device"
    d = str(tmp_path_factory.mktemp("multi"))
    path = ru.scene_file("spheres", d, 96, 96, 24)
    return capi.read_scene(path)


def _devices(n):
    have = capi.gpu.slrgpu_device_count()
    return [g % have for g in range(n)]


@pytest.mark.parametrize("replicas", [2, 3])
def test_partitioned_frame_equals_single_gpu_frame(scene, replicas):
    gss = [capi.GpuScene(scene, device=d) for d in _devices(replicas)]
    whole, st1 = capi.gpu_render(gss[0], 96, 96, 0, 24)
    multi, stn = capi.gpu_render_multi(gss, 96, 96, 0, 24)
    for k in ("paths", "rays", "extend_rays", "shadow_rays", "class_hits"):
        assert stn[k] == st1[k], k
    np.testing.assert_allclose(multi, whole, rtol=2e-4, atol=1e-5 * float(whole.mean()))
    assert stn["other_ms"] > 0.0                     # the exchange step ran and was timed


def test_partitioned_bidirectional_frame_equals_single_gpu_frame(scene):
    """SLRGPU_RENDER_BPT through slrgpu_render_multi: the same sample partition and exchange step; a bidirectional sample's
    random numbers are keyed by (pixel, global sample index) too."""
    gss = [capi.GpuScene(scene, device=d) for d in _devices(2)]
    whole, st1 = capi.gpu_render(gss[0], 96, 96, 0, 8, flags=capi.RENDER_BPT)
    multi, stn = capi.gpu_render_multi(gss, 96, 96, 0, 8, flags=capi.RENDER_BPT)
    for k in ("paths", "rays", "extend_rays", "shadow_rays"):
        assert stn[k] == st1[k], k
    np.testing.assert_allclose(multi, whole, rtol=5e-4, atol=1e-4 * float(whole.mean()))


def test_more_replicas_than_samples(scene):
    """A replica whose share of the sample range is empty renders nothing; the frame is still complete."""
    gss = [capi.GpuScene(scene, device=d) for d in _devices(4)]
    whole, st1 = capi.gpu_render(gss[0], 96, 96, 5, 7)
    multi, stn = capi.gpu_render_multi(gss, 96, 96, 5, 7)
    assert stn["paths"] == st1["paths"] == 96 * 96 * 2
    np.testing.assert_allclose(multi, whole, rtol=2e-4, atol=1e-5 * float(whole.mean()))


def test_renderer_front_end_uses_every_visible_device(scene):
    """slrhost_render with device < 0 = GPUPathTracingRenderer over all visible devices (what a scene file's
    setRenderer("PT") gets): same image as the single-device renderer."""
    one, st1 = capi.host_render(scene, 96, 96, 24, device=0)
    every, stn = capi.host_render(scene, 96, 96, 24, device=-1)
    assert st1["devices"] == 1 and stn["devices"] == min(capi.gpu.slrgpu_device_count(), 24)
    assert stn["paths"] == st1["paths"]
    np.testing.assert_allclose(every, one, rtol=2e-4, atol=1e-5 * float(one.mean()))


def test_multi_argument_errors(scene):
    gs = capi.GpuScene(scene)
    with pytest.raises(capi.SlrError):
        capi.gpu_render_multi([gs], 96, 96, 4, 4)
