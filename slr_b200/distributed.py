"""Sample-parallel multi-GPU rendering: one process per GPU, the frame's samples-per-pixel partitioned
over the ranks, scene replicated, ONE sum-reduce of the float accumulation buffers per frame (SURVEY.md
section 8e). The renderer itself needs no communication: a path's random numbers are keyed by (pixel,
global sample index), so any partition of the sample range renders the same set of paths."""


def sample_range(rank, world, spp, mode="strong"):
    """[begin, end) of global sample indices rank `rank` renders.
    strong: `spp` is the frame total, split as evenly as possible (the first spp % world ranks get one more);
    weak:   every rank renders `spp` samples of its own (frame total = spp * world)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank/world out of range")
    if mode == "weak":
        return rank * spp, (rank + 1) * spp
    base, extra = divmod(spp, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reduce_frame(accum, dist, dst=0):
    """Sum the per-rank accumulation buffers onto rank `dst` (torch.distributed: NCCL over NVLink on
    GPUs, gloo in the CPU tests). No-op for a single process."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst)
    return accum
