"""CPU (gloo, world_size 2 and 3): the sample partition + single frame reduce that bench.py / a
multi-GPU renderer front end use. The per-sample contribution is a deterministic function of
(pixel, global sample index) -- the property the counter-based path RNG gives the real renderer -- so
the reduced frame of N ranks must equal the single-process frame."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slr_b200.distributed import reduce_frame, sample_range

W, H, C, SPP = 16, 12, 16, 13


def fake_pass(sample):
    """Stand-in for one sample per pixel: depends only on (pixel, channel, global sample index)."""
    p = np.arange(W * H * C, dtype=np.uint64).reshape(H, W, C)
    h = (p * np.uint64(2654435761) + np.uint64(sample) * np.uint64(40503)) % np.uint64(1 << 20)
    return h.astype(np.float32) / np.float32(1 << 20)


def render_range(begin, end):
    acc = np.zeros((H, W, C), np.float32)
    for s in range(begin, end):
        acc += fake_pass(s)
    return acc


def _worker(rank, world, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sample_range(rank, world, SPP, mode)
    acc = torch.from_numpy(render_range(b, e))
    reduce_frame(acc, dist, dst=0)
    if rank == 0:
        np.save(out, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,mode", [(2, "strong"), (3, "strong"), (2, "weak")])
def test_partition_plus_reduce_equals_single_process(world, mode, tmp_path):
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    total = SPP if mode == "strong" else SPP * world
    np.testing.assert_allclose(np.load(out), render_range(0, total), rtol=1e-6)


def test_sample_ranges_tile_the_frame():
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 7, 64, 1024):
            r = [sample_range(g, world, spp) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    assert sample_range(3, 8, 64, "weak") == (192, 256)
    with pytest.raises(ValueError):
        sample_range(2, 2, 8)
