"""Path-tracing workloads of bench.py (see bench.py's docstring for the command line contract).

Workload `cornell_spheres` = BASELINE.json configs[0] (the configuration north_star's target is quoted
on): Cornell_Box_Spheres, 512x512, 64 spp, unidirectional PT, spectral. One step = one frame.

  value   frame throughput with the scene resident in HBM: per step every rank clears its device
          accumulation buffer, renders its samples (slrgpu_render_device) and -- N > 1 -- the buffers are
          summed onto rank 0 with one NCCL reduce. CUDA events on the launch stream, max over ranks.
          Weak scaling: every rank renders the full 64 spp of its own sample range [64 r, 64 (r + 1)),
          i.e. the N-GPU frame has 64 N spp (BASELINE.json configs[3] splits 1024 spp as 128 per GPU).
  e2e     the same frame through the renderer front end with HOST buffers: every rank calls
          slrhost_render_range (= GPUPathTracingRenderer::render, the drop-in for
          PathTracingRenderer::render, on its sample range): scene upload from the host SoA buffers,
          render, download of the frame buffer into (pinned) host memory. N > 1: the host frame buffers
          are staged back to the device, summed with one NCCL reduce and downloaded by rank 0. Wall
          clock between barriers, median over the frames.
"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRIC = "Mpaths/s (unidirectional path tracing, camera samples per second)"

WORKLOADS = {
    # name: (scene key of slr_b200.scenes.SCENES, width, height, spp, description)
    "cornell_spheres": ("spheres", 512, 512, 64,
                        "Cornell_Box_Spheres 512x512 64spp unidirectional PT spectral (BASELINE configs[0])"),
    "cornell_diffuse": ("diffuse", 512, 512, 64, "empty Cornell box 512x512 64spp unidirectional PT spectral"),
    "materials": ("materials", 1024, 1024, 256,
                  "Cornell_Box_ColorChecker 1024x1024 256spp, GGX/Ward/Oren-Nayar/Ashikhmin/mix/rough-glass with procedural textures (BASELINE configs[1])"),
    "ibl": ("ibl_full", 1024, 1024, 256,
            "IBL_Test 1024x1024 256spp, synthetic 2048x1024 HDR environment, importance sampling + thin lens (BASELINE configs[2])"),
    "instanced": ("instanced_10m", 1920, 1080, 128,
                  "100 instances of a 100,488-triangle procedural mesh (10 M instanced triangles, two-level SBVH->QBVH), 1920x1080, "
                  "128 spp per GPU = 1024 spp on 8 GPUs (BASELINE configs[3])"),
}


def _scene(args):
    from . import scenes
    key, w, h, spp, desc = WORKLOADS[args.workload]
    if getattr(args, "size", 0):
        w = h = args.size
    if getattr(args, "spp", 0):
        spp = args.spp
    d = tempfile.mkdtemp(prefix="slr_bench_")
    path = scenes.SCENES[key](d, width=w, height=h, spp=spp)
    return path, w, h, spp, desc


def _ref_step(path, w, h, spp):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import render_util as ru
    t0 = time.perf_counter()
    _, j = ru.run_ref_render(path, spp, w, h)
    wall = time.perf_counter() - t0
    return j, wall


def bounded_spp(w, h, spp, max_paths=40e6):
    """Samples per pixel of one reference step: the whole workload when it is at most ~40 M paths (C1: all 64 spp),
    else the largest count that keeps a step near 10-15 s of host time. Mpaths/s does not depend on it (the
    reference's per-pass time is constant)."""
    return int(max(1, min(spp, max_paths // (w * h))))


def cpu_baseline(path, w, h, spp):
    """The reference's own PathTracingRenderer (oracle/_ref/ref_render, built from /root/reference) on
    this box's host cores: std::thread::hardware_concurrency() workers, shipped default accelerator (SBVH)."""
    spp = bounded_spp(w, h, spp)
    j, wall = _ref_step(path, w, h, spp)
    return {"value": j["mpaths_per_s"], "unit": "Mpaths/s", "cores": j["threads"], "kind": "reference",
            "sample": f"the same scene at {w}x{h}, {spp} spp ({w * h * spp} paths) through the reference's "
                      f"PathTracingRenderer::render, {j['threads']} threads, {j['accelerator']}; render {j['render_s']:.2f} s "
                      f"(scene read+build {j['read_s'] + j['build_s']:.2f} s not counted)"}


def run_reference(args, rank):
    if rank != 0:
        return
    path, w, h, spp, desc = _scene(args)
    total = args.steps + args.warmup
    step_spp = bounded_spp(w, h, spp, 40e6 if total <= 20 else 10e6)     # keep the whole run within a few minutes
    vals = []
    t0 = time.perf_counter()
    for i in range(total):
        j, wall = _ref_step(path, w, h, step_spp)
        if i >= args.warmup:
            vals.append(j)
    wall = time.perf_counter() - t0
    mp = float(np.mean([j["mpaths_per_s"] for j in vals]))
    render_s = float(np.mean([j["render_s"] for j in vals]))
    cb = {"value": mp, "unit": "Mpaths/s", "cores": vals[0]["threads"], "kind": "reference",
          "sample": f"{w}x{h}, {step_spp} spp per step through the reference's PathTracingRenderer::render "
                    f"({vals[0]['threads']} threads, {vals[0]['accelerator']}); mean of {len(vals)} steps"}
    line = {"impl": "reference", "metric": METRIC, "value": mp, "unit": "Mpaths/s", "n_gpus": 0, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * render_s, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": w, "height": h, "spp": spp, "spp_per_step": step_spp},
            "cpu_baseline": cb,
            "e2e": {"value": mp, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main(args, rank, world):
    if args.impl == "reference":
        return run_reference(args, rank)
    import torch
    import bench
    from . import capi
    from .distributed import reduce_frame, sample_range

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    path, w, h, spp, desc = _scene(args)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    gs = capi.GpuScene(hs, device=dev)
    chan = capi.gpu.slrgpu_scene_channels(gs.handle)
    accum = torch.zeros((h, w, chan), dtype=torch.float32, device="cuda")
    # a non-default stream: slrgpu_render_device replays a captured CUDA graph, which the legacy stream cannot do
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    seed = 1509761209
    spp_begin, spp_end = sample_range(rank, world, spp, "weak")
    params = capi.RenderParams(C.sizeof(capi.RenderParams), w, h, spp_begin, spp_end, 0.0, 0.0, seed, 0,
                               getattr(args, "pool", 0) or 0, 0)

    def frame(flags=0):
        params.flags = flags
        st = capi.RenderStats()
        accum.zero_()
        rc = capi.gpu.slrgpu_render_device(gs.handle, C.byref(params), C.c_void_p(accum.data_ptr()),
                                           C.c_void_p(stream.cuda_stream), C.byref(st))
        if rc != 0:
            raise RuntimeError(capi.gpu.slrgpu_last_error().decode())
        reduce_frame(accum, dist)
        return st

    # one plain frame (allocates the pooled queues), then one profiled, untimed frame: per-kernel-family device
    # times and the traversal counts of the algorithmic-bytes model
    frame()
    torch.cuda.synchronize()
    prof = frame(capi.RENDER_PROFILE_STAGES)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        frame()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    rays = 0
    with bench.ClockSampler(dev) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            st = frame()
            launches += st.kernel_launches + (1 if dist is not None else 0) + 1     # + reduce, + clear
            rays += st.rays
        e1.record(stream)
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    paths_per_step = w * h * spp

    # ---- end to end with host buffers
    # every rank: Renderer::render of its sample range through the host front end (scene upload + render +
    # frame-buffer download into pinned host memory); N > 1: the host frame buffers go back to the device,
    # one NCCL reduce, and rank 0 downloads the sum. Median of the per-frame wall times.
    e2e_steps = max(3, min(args.steps, 10))
    pinned = torch.empty((h, w, chan), dtype=torch.float32).pin_memory()
    staged = torch.empty((h, w, chan), dtype=torch.float32, device="cuda") if dist is not None else None
    hst = None

    def e2e_frame():
        nonlocal hst
        _, hst = capi.host_render(hs, w, h, spp, seed, dev, spp_begin=spp_begin, out=pinned.numpy())
        if dist is not None:
            staged.copy_(pinned, non_blocking=True)
            reduce_frame(staged, dist)
            if rank == 0:
                pinned.copy_(staged, non_blocking=True)
            torch.cuda.synchronize()

    e2e_frame()
    times = []
    for _ in range(e2e_steps):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_frame()
        if dist is not None:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    e2e_s = float(np.median(times))
    e2e_detail = {"frames": e2e_steps, "wall_ms_median": round(1e3 * e2e_s, 2), "wall_ms_min": round(1e3 * min(times), 2),
                  "wall_ms_max": round(1e3 * max(times), 2), "renderer_wall_ms": round(1e3 * hst["wall_s"], 2),
                  "scene_upload_ms": round(1e3 * hst["upload_s"], 2), "device_ms": round(1e3 * hst["device_s"], 2)}
    scene_bytes = int(gs.device_bytes)
    accum_bytes = w * h * chan * 4

    if dist is not None:
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        tr = torch.tensor([float(rays)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tr)
        rays = float(tr[0])
    else:
        rays = float(rays)
    if rank != 0:
        dist.destroy_process_group()
        return

    peak, peak_src = bench.measured_peaks()
    value = paths_per_step * world * args.steps / (ms * 1e-3) / 1e6
    # dominant kernel by the profiled frame's stage times
    stages = {"extendKernel": prof.extend_ms, "surfaceKernel": prof.surface_ms, "materialKernel": prof.material_ms,
              "shadowKernel": prof.shadow_ms, "raygenKernel": prof.raygen_ms}
    dom = max(stages, key=stages.get)
    S_PATH, S_HIT, S_SHADOW = (120, 24, 108) if chan == 16 else (72, 24, 60)
    n_ext, n_sh = prof.extend_rays, prof.shadow_rays
    algo = {
        # 32 B ray in + 128 B per node popped + 48 B per leaf record tested + 24 B hit record out
        "extendKernel": 32 * n_ext + 128 * prof.extend_nodes + 48 * prof.extend_leaf_records + S_HIT * n_ext,
        # shadow entry in + nodes + leaf records (+ the splat, counted as 64 B read-modify-write of unoccluded entries; upper bound: all)
        "shadowKernel": S_SHADOW * n_sh + 128 * prof.shadow_nodes + 48 * prof.shadow_leaf_records,
        # hit id + meta + triangle record + roulette slot in; roulette slot + class-queue entry out (upper bound: every ray survives)
        "surfaceKernel": (8 + 16 + 32 + 4) * n_ext + (4 + 8) * n_ext,
        # (all class kernels of a frame) path state + hit + class entry + triangle + 3 vertices in per surviving hit (bounded by
        # the extend rays), surviving path state out (= the non-camera extend rays), shadow entry out
        "materialKernel": (S_PATH + S_HIT + 8 + 176) * n_ext + S_PATH * (n_ext - paths_per_step) + S_SHADOW * n_sh,
        "raygenKernel": S_PATH * paths_per_step,
    }
    # share of the step the dominant kernel takes in the profiled frame, applied to the timed steps
    share = stages[dom] / max(prof.device_ms, 1e-9)
    dom_ms_per_step = share * ms / args.steps
    ach = algo[dom] / (dom_ms_per_step * 1e-3) / 1e9
    # measured DRAM traffic of that kernel family over one frame, from the committed ncu capture of this workload
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(args.workload, {}).get(dom)
        if t is not None:
            traffic = t["dram_bytes_per_frame"]
    except (OSError, ValueError, KeyError):
        traffic = None
    cpu = cpu_baseline(path, w, h, spp) if world == 1 else None      # the CPU leg runs at N = 1 only
    line = {"metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": w, "height": h, "spp_per_gpu": spp, "frame_spp": spp * world,
                       "paths_per_gpu_per_step": paths_per_step, "rays_per_path": rays / (paths_per_step * world * args.steps),
                       "mrays_per_s": rays / (ms * 1e-3) / 1e6, "waves_per_frame": int(prof.waves),
                       "extend_nodes_per_ray": round(prof.extend_nodes / max(n_ext, 1), 3), "extend_leaf_records_per_ray": round(prof.extend_leaf_records / max(n_ext, 1), 3),
                       "shadow_nodes_per_ray": round(prof.shadow_nodes / max(n_sh, 1), 3), "shadow_leaf_records_per_ray": round(prof.shadow_leaf_records / max(n_sh, 1), 3),
                       "class_hits_per_frame": [int(x) for x in prof.class_hits], "extend_rays_per_frame": int(n_ext), "shadow_rays_per_frame": int(n_sh), "pool": int(getattr(args, "pool", 0) or (1 << 24)),
                       "triangles": int(hs.desc.num_triangles), "qbvh_nodes": int(hs.desc.num_bvh_nodes),
                       "scene_bytes": scene_bytes, "host_scene_build_s": round(hs.build_seconds, 3),
                       "l2_policy": "per-step working set (wavefront queues + accumulation buffer, > 500 MB) is larger than L2; "
                                    "the scene itself (QBVH + leaf records) is L2-resident by design",
                       "stage_ms_profiled_frame": {k: round(v, 3) for k, v in stages.items()} | {"tailKernel": round(prof.tail_ms, 3), "other": round(prof.other_ms, 3), "frame": round(prof.device_ms, 3)},
                       "tail_kernel": {"paths": int(prof.tail_paths), "bounces": int(prof.tail_waves)}},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "kernel": dom, "peak_source": peak_src, "kernel_share_of_step": share,
                         "algorithmic_bytes_per_launch_set": algo[dom],
                         "note": "achieved = algorithmic bytes of all launches of the kernel family in one frame / their summed device time; "
                                 "traffic = ncu dram__bytes_read+write summed over the same launches (profiles/traffic.json), bytes per frame. "
                                 "The algorithmic bytes of the ray kernels count every node / leaf-record fetch (SURVEY 8d); with a scene "
                                 "that fits L1/L2 most of them never reach HBM, so frac can approach or pass 1 while traffic stays small"},
            "cpu_baseline": cpu,
            "e2e": {"value": paths_per_step * world / e2e_s / 1e6, "unit": "Mpaths/s",
                    "h2d_bytes_per_step": scene_bytes + (accum_bytes if world > 1 else 0),
                    "d2h_bytes_per_step": accum_bytes * (2 if world > 1 else 1), "breakdown": e2e_detail},
            "gpu_launches": int(launches), "clocks": clocks.summary()}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
