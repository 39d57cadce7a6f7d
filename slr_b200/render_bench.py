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


def _scene(args, directory=None):
    """Writes the workload's scene file (+ synthetic assets) and returns (path, width, height, spp, description).
    cornell_spheres renders the reference's own TestScenes/Cornell_Box_Spheres.txt UNCHANGED (size / sample override
    appended) when bench.py found the shipped scene files (args.ref_scenes); else the re-emitted text of scenes.py."""
    from . import scenes
    key, w, h, spp, desc = WORKLOADS[args.workload]
    if getattr(args, "size", 0):
        w = h = args.size
    if getattr(args, "spp", 0):
        spp = args.spp
    d = directory or tempfile.mkdtemp(prefix="slr_bench_")
    src = os.path.join(getattr(args, "ref_scenes", "") or "", "Cornell_Box_Spheres.txt")
    if args.workload == "cornell_spheres" and os.path.exists(src):
        path = scenes.write_reference_scene(src, d, w, h, spp)
        desc += "; scene file = the reference's TestScenes/Cornell_Box_Spheres.txt unchanged + appended size/spp override"
    else:
        path = scenes.SCENES[key](d, width=w, height=h, spp=spp)
    return path, w, h, spp, desc


def _scene_in_subprocess(args):
    """The same scene file written by a child process: the reference arm must not map this repo's native libraries
    (the synthetic .assbin / .exr writers live in libslrhost.so), so the parent only receives the path."""
    import subprocess
    d = tempfile.mkdtemp(prefix="slr_bench_ref_")
    code = ("import json, sys, types; sys.path.insert(0, %r); from slr_b200 import render_bench as rb; "
            "a = types.SimpleNamespace(**json.loads(sys.argv[1])); print(json.dumps(rb._scene(a, sys.argv[2])))" % ROOT)
    cfg = {"workload": args.workload, "size": getattr(args, "size", 0), "spp": getattr(args, "spp", 0),
           "ref_scenes": getattr(args, "ref_scenes", "")}
    out = subprocess.run([sys.executable, "-c", code, json.dumps(cfg), d], capture_output=True, text=True, check=True).stdout
    return tuple(json.loads(out.strip().splitlines()[-1]))


def _ref_step(path, w, h, spp, seed=0, want_image=False):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import render_util as ru
    t0 = time.perf_counter()
    img, j = ru.run_ref_render(path, spp, w, h, seed=seed)
    wall = time.perf_counter() - t0
    return (j, wall, img) if want_image else (j, wall)


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
    path, w, h, spp, desc = _scene_in_subprocess(args)
    total = args.steps + args.warmup
    # C1 (16.8 M paths, ~4 s of host time) is rendered whole every step -- the same 64 spp the GPU arm renders; the larger
    # workloads render a bounded number of passes per step (Mpaths/s is a rate: the reference's per-pass time is constant)
    step_spp = bounded_spp(w, h, spp)
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


def _profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def roofline_line(workload, kernel, algo_bytes, achieved_gbs, peak, peak_src, share, paths_per_frame):
    """The roofline object of the JSON line for the frame's dominant kernel family.

    The algorithmic-bytes figure (every node / leaf-record fetch counted at full size) over the kernel's
    measured time is always reported against the measured HBM peak (`hbm`). Whether HBM is what BOUNDS the kernel is decided
    from the committed ncu evidence of this workload: profiles/traffic.json holds the DRAM bytes the kernel family really
    moved per path; when that is under half of the algorithmic bytes the fetches are served by L1 / L2 (the scene is cache
    resident), the HBM fraction says nothing about kernel quality, and the line is labelled with what ncu shows bounds it:
    issue slots at the measured lanes-active (profiles/kernel_metrics.json, copied from the `--set full` capture)."""
    t = _profile_json("traffic.json").get(workload, {}).get(kernel)
    traffic = None
    if t is not None:
        per_path = t.get("dram_bytes_per_path")
        traffic = per_path * paths_per_frame if per_path is not None else t.get("dram_bytes_per_frame")
    m = _profile_json("kernel_metrics.json").get(workload, {}).get(kernel)
    hbm = {"achieved": achieved_gbs, "peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak, "peak_source": peak_src,
           "algorithmic_bytes_per_frame": algo_bytes}
    cache_resident = traffic is not None and traffic < 0.5 * algo_bytes
    if cache_resident and m is not None:
        return {"bound": "issue", "achieved": m["issue_slot_utilisation_pct"], "peak": 100.0, "unit": "% of issue slots",
                "frac": m["issue_slot_utilisation_pct"] / 100.0, "traffic": traffic, "kernel": kernel, "kernel_share_of_step": share,
                "lanes_active_of_32": m["lanes_active"], "l1_hit_pct": m.get("l1_hit_pct"), "achieved_occupancy_pct": m.get("occupancy_pct"),
                "source": m["source"], "hbm": hbm,
                "note": "the kernel's node / leaf-record fetches hit L1 / L2 (measured DRAM traffic is %.0f %% of the algorithmic bytes), "
                        "so the bound is instruction issue on partly filled warps, not bandwidth: achieved = ncu smsp__issue_active of the "
                        "committed capture of this workload (not re-measured in this run); `hbm` = the algorithmic bytes of all launches of "
                        "the kernel family in one frame / their device time in THIS run, against the measured HBM peak"
                        % (100.0 * traffic / algo_bytes)}
    return {"bound": "hbm", "achieved": achieved_gbs, "peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak, "traffic": traffic,
            "kernel": kernel, "peak_source": peak_src, "kernel_share_of_step": share, "algorithmic_bytes_per_launch_set": algo_bytes,
            "lanes_active_of_32": m["lanes_active"] if m else None, "issue_slot_utilisation_pct": m["issue_slot_utilisation_pct"] if m else None,
            "note": "achieved = algorithmic bytes of all launches of the kernel family in one frame / their summed device time; "
                    "traffic = ncu dram__bytes_read+write of the same launches (profiles/traffic.json), scaled to this frame's paths"}


def image_parity(capi, path, w, h, spp, gpu_accum):
    """The frame the e2e leg just rendered (final pipeline, full size) against the reference's PathTracingRenderer at the same
    size and sample count: relRMSE(gpu, ref1) next to the reference's own two-seed noise floor relRMSE(ref2, ref1), on
    linear sRGB with the largest 0.5 % of squared errors trimmed on both sides (tests/test_gpu_render.py's statistic)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import render_util as ru
    gpu = capi.accum_to_rgb(np.array(gpu_accum, copy=True), 1.0 / spp)
    ref1 = capi.accum_to_rgb(_ref_step(path, w, h, spp, seed=1509761209, want_image=True)[2], 1.0 / spp)
    ref2 = capi.accum_to_rgb(_ref_step(path, w, h, spp, seed=20240229, want_image=True)[2], 1.0 / spp)
    (ref1, gpu, ref2), d1 = ru.sanitize_reference(ref1, gpu, ref2)
    (ref2, gpu, ref1), d2 = ru.sanitize_reference(ref2, gpu, ref1)
    floor = ru.rel_rmse(ref2, ref1, trim=0.005)
    got = ru.rel_rmse(gpu, ref1, trim=0.005)
    clip = float(np.percentile(ref1, 99.8))
    ratio = np.minimum(gpu, clip).reshape(-1, 3).mean(0) / np.minimum(ref1, clip).reshape(-1, 3).mean(0)
    return {"width": w, "height": h, "spp": spp, "rel_rmse_gpu_vs_ref": round(got, 5), "rel_rmse_floor_ref_vs_ref": round(floor, 5),
            "ratio_to_floor": round(got / floor, 4), "tolerance": "<= 1.25 x floor", "within_tolerance": bool(got <= 1.25 * floor),
            "image_mean_ratio_rgb": [round(float(x), 5) for x in ratio], "reference_nan_pixels_dropped": int(d1 + d2)}


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
    # weak (default, what the driver's scaling run measures): every rank renders `spp` samples of its own, the frame has
    # spp x N; strong (--scaling strong): `spp` is the FRAME's sample count, partitioned over the ranks (north_star: "the
    # frame's samples-per-pixel are partitioned across the GPUs")
    mode = getattr(args, "scaling", "weak") or "weak"
    spp_begin, spp_end = sample_range(rank, world, spp, mode)
    my_spp = spp_end - spp_begin
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
    # one event per step boundary: the K-step span e[0] -> e[K] is the contract's number; the per-step times next to it
    # show whether a single host stall (the reduce of a straggling rank, a scheduler hiccup) sits inside it
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches = 0
    rays = 0
    with bench.ClockSampler(dev) as clocks:
        ev[0].record(stream)
        for k in range(args.steps):
            st = frame()
            ev[k + 1].record(stream)
            launches += st.kernel_launches + (1 if dist is not None else 0) + 1     # + reduce, + clear
            rays += st.rays
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = ev[0].elapsed_time(ev[-1])
    step_ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    my_steps = {"median": step_ms[len(step_ms) // 2], "min": step_ms[0], "max": step_ms[-1], "sum": float(sum(step_ms))}
    paths_per_step = w * h * my_spp

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
        _, hst = capi.host_render(hs, w, h, my_spp, seed, dev, spp_begin=spp_begin, out=pinned.numpy())
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

    per_rank_steps = [my_steps]
    total_paths_per_step = paths_per_step
    if dist is not None:
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        tr = torch.tensor([float(rays), float(paths_per_step)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tr)
        rays, total_paths_per_step = float(tr[0]), int(tr[1])
        g = torch.zeros((world, 4), device="cuda", dtype=torch.float64)
        g[rank] = torch.tensor([my_steps["median"], my_steps["min"], my_steps["max"], my_steps["sum"]], dtype=torch.float64)
        dist.all_reduce(g)
        per_rank_steps = [{"median": float(r[0]), "min": float(r[1]), "max": float(r[2]), "sum": float(r[3])} for r in g.cpu()]
    else:
        rays = float(rays)
    if rank != 0:
        dist.destroy_process_group()
        return

    peak, peak_src = bench.measured_peaks()
    value = total_paths_per_step * args.steps / (ms * 1e-3) / 1e6
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
    roof = roofline_line(args.workload, dom, algo[dom], ach, peak, peak_src, share, paths_per_step)
    # SLR_BENCH_AB=1 (kernel A/B sweeps, tools/r02_call*.sh): skip the CPU legs, only the device numbers are read
    ab = os.environ.get("SLR_BENCH_AB") == "1"
    cpu = cpu_baseline(path, w, h, spp) if (world == 1 and not ab) else None      # the CPU leg runs at N = 1 only
    parity = image_parity(capi, path, w, h, my_spp, pinned.numpy()) if (world == 1 and not ab and w * h * my_spp <= getattr(args, "parity_paths", 40e6)) else None
    line = {"metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": mode, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": w, "height": h, "spp_per_gpu": my_spp, "frame_spp": total_paths_per_step // (w * h),
                       "paths_per_gpu_per_step": paths_per_step, "rays_per_path": rays / (total_paths_per_step * args.steps),
                       "mrays_per_s": rays / (ms * 1e-3) / 1e6, "waves_per_frame": int(prof.waves),
                       "extend_nodes_per_ray": round(prof.extend_nodes / max(n_ext, 1), 3), "extend_leaf_records_per_ray": round(prof.extend_leaf_records / max(n_ext, 1), 3),
                       "shadow_nodes_per_ray": round(prof.shadow_nodes / max(n_sh, 1), 3), "shadow_leaf_records_per_ray": round(prof.shadow_leaf_records / max(n_sh, 1), 3),
                       "class_hits_per_frame": [int(x) for x in prof.class_hits], "extend_rays_per_frame": int(n_ext), "shadow_rays_per_frame": int(n_sh), "pool": int(getattr(args, "pool", 0) or (1 << 24)),
                       "triangles": int(hs.desc.num_triangles), "qbvh_nodes": int(hs.desc.num_bvh_nodes),
                       "scene_bytes": scene_bytes, "host_scene_build_s": round(hs.build_seconds, 3),
                       "l2_policy": "per-step working set (wavefront queues + accumulation buffer, > 500 MB) is larger than L2; "
                                    "the scene itself (QBVH + leaf records) is L2-resident by design",
                       "stage_ms_profiled_frame": {k: round(v, 3) for k, v in stages.items()} | {"tailKernel": round(prof.tail_ms, 3), "other": round(prof.other_ms, 3), "frame": round(prof.device_ms, 3)},
                       "tail_kernel": {"paths": int(prof.tail_paths), "bounces": int(prof.tail_waves)},
                       # device time of every timed step (CUDA events at the step boundaries), per rank: a straggler or a
                       # host stall inside the K-step span shows up as max >> median
                       "step_ms_per_rank": [{k: round(v, 3) for k, v in r.items()} for r in per_rank_steps]},
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": total_paths_per_step / e2e_s / 1e6, "unit": "Mpaths/s",
                    "h2d_bytes_per_step": scene_bytes + (accum_bytes if world > 1 else 0),
                    "d2h_bytes_per_step": accum_bytes * (2 if world > 1 else 1), "breakdown": e2e_detail},
            "gpu_launches": int(launches), "clocks": clocks.summary()}
    if parity is not None:
        line["image_parity"] = parity
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
