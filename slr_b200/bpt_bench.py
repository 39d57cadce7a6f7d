"""Bidirectional path tracing workload of bench.py (--workload cornell_spheres_bpt): the reference's
TestScenes/Cornell_Box_Spheres.txt as it ships -- it selects "BPT" -- at 512x512, 16 samples per pixel per step, through
the GPU twin of BidirectionalPathTracingRenderer (csrc/bpt.cu) next to the reference's own on the host cores. One step =
one frame; metric = bidirectional samples per second (a sample = one light subpath + one eye subpath + all connections).

  value         device-resident scene, slrgpu_render_device with SLRGPU_RENDER_BPT, CUDA events on the launch stream
  e2e           slrhost_render_bpt with host buffers: scene upload + render + frame download, wall clock, median
  cpu_baseline  oracle/_ref/ref_render ... bpt (the reference's BidirectionalPathTracingRenderer::render, all host threads)
  image_parity  the e2e frame against two seeds of the reference's BPT at the same size and sample count
Single GPU only (the multi-GPU partition is the path tracer's: sample ranges + one sum, slrgpu_render_multi).
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

from . import render_bench as rb

METRIC = "Msamples/s (bidirectional path tracing, camera samples per second)"
SIZE, SPP = 512, 16


def _args_scene(args):
    a = type("A", (), {})()
    a.workload = "cornell_spheres"
    a.size = getattr(args, "size", 0) or SIZE
    a.spp = getattr(args, "spp", 0) or SPP
    a.ref_scenes = getattr(args, "ref_scenes", "")
    return a


def _describe(desc, w, h, spp):
    tail = desc.split(";", 1)[1] if ";" in desc else ""
    return (f"Cornell_Box_Spheres {w}x{h} {spp}spp bidirectional PT spectral (the renderer the shipped scene file selects; "
            f"BASELINE configs[0]'s scene)" + (";" + tail if tail else ""))


def _ref_bpt(path, w, h, spp, seed=0):
    sys.path.insert(0, os.path.join(rb.ROOT, "tests"))
    import render_util as ru
    return ru.run_ref_render(path, spp, w, h, seed=seed, bpt=True)


def run_reference(args, rank):
    if rank != 0:
        return
    path, w, h, spp, desc = rb._scene_in_subprocess(_args_scene(args))
    step_spp = max(1, min(spp, 4))
    vals = []
    for i in range(args.steps + args.warmup):
        _, j = _ref_bpt(path, w, h, step_spp)
        if i >= args.warmup:
            vals.append(j)
    mp = float(np.mean([j["mpaths_per_s"] for j in vals]))
    cb = {"value": mp, "unit": "Msamples/s", "cores": vals[0]["threads"], "kind": "reference",
          "sample": f"{w}x{h}, {step_spp} spp per step through the reference's BidirectionalPathTracingRenderer::render "
                    f"({vals[0]['threads']} threads); mean of {len(vals)} steps"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": mp, "unit": "Msamples/s", "n_gpus": 0, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([j["render_s"] for j in vals])), "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": _describe(desc, w, h, spp), "width": w, "height": h, "spp": spp, "spp_per_step": step_spp},
                      "cpu_baseline": cb, "e2e": {"value": mp, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main(args, rank, world):
    if args.impl == "reference":
        return run_reference(args, rank)
    if world > 1:
        raise SystemExit("cornell_spheres_bpt is a single-GPU workload")
    import torch
    import bench
    from . import capi
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    path, w, h, spp, desc = rb._scene(_args_scene(args))
    desc = _describe(desc, w, h, spp)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    gs = capi.GpuScene(hs, device=dev)
    chan = capi.gpu.slrgpu_scene_channels(gs.handle)
    accum = torch.zeros((h, w, chan), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    seed = 1509761209
    params = capi.RenderParams(C.sizeof(capi.RenderParams), w, h, 0, spp, 0.0, 0.0, seed, 0, 0, capi.RENDER_BPT)

    def frame():
        st = capi.RenderStats()
        accum.zero_()
        rc = capi.gpu.slrgpu_render_device(gs.handle, C.byref(params), C.c_void_p(accum.data_ptr()), C.c_void_p(stream.cuda_stream), C.byref(st))
        if rc != 0:
            raise RuntimeError(capi.gpu.slrgpu_last_error().decode())
        return st

    for _ in range(args.warmup):
        frame()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches = 0
    with bench.ClockSampler(dev) as clocks:
        ev[0].record(stream)
        for k in range(args.steps):
            st = frame()
            ev[k + 1].record(stream)
            launches += st.kernel_launches + 1
        torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[-1])
    step_ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    samples = w * h * spp
    value = samples * args.steps / (ms * 1e-3) / 1e6

    pinned = torch.empty((h, w, chan), dtype=torch.float32).pin_memory()
    times = []
    hst = None
    for i in range(max(3, min(args.steps, 10)) + 1):
        t0 = time.perf_counter()
        _, hst = capi.host_render(hs, w, h, spp, seed, dev, out=pinned.numpy(), method="BPT")
        if i:
            times.append(time.perf_counter() - t0)
    e2e_s = float(np.median(times))
    ab = os.environ.get("SLR_BENCH_AB") == "1"
    cpu = None
    parity = None
    if not ab:
        cspp = max(1, min(spp, 4))
        _, j = _ref_bpt(path, w, h, cspp)
        cpu = {"value": j["mpaths_per_s"], "unit": "Msamples/s", "cores": j["threads"], "kind": "reference",
               "sample": f"the same scene at {w}x{h}, {cspp} spp ({w * h * cspp} samples) through the reference's "
                         f"BidirectionalPathTracingRenderer::render, {j['threads']} threads; render {j['render_s']:.2f} s"}
        sys.path.insert(0, os.path.join(rb.ROOT, "tests"))
        import render_util as ru
        gpu = capi.accum_to_rgb(np.array(pinned.numpy(), copy=True), 1.0 / spp)
        ref1 = capi.accum_to_rgb(_ref_bpt(path, w, h, spp, seed=1509761209)[0], 1.0 / spp)
        ref2 = capi.accum_to_rgb(_ref_bpt(path, w, h, spp, seed=20240229)[0], 1.0 / spp)
        (ref1, gpu, ref2), d1 = ru.sanitize_reference(ref1, gpu, ref2)
        (ref2, gpu, ref1), d2 = ru.sanitize_reference(ref2, gpu, ref1)
        floor = ru.rel_rmse(ref2, ref1, trim=0.005)
        got = ru.rel_rmse(gpu, ref1, trim=0.005)
        clip = float(np.percentile(ref1, 99.8))
        ratio = np.minimum(gpu, clip).reshape(-1, 3).mean(0) / np.minimum(ref1, clip).reshape(-1, 3).mean(0)
        parity = {"width": w, "height": h, "spp": spp, "against": "the reference's BidirectionalPathTracingRenderer, two seeds",
                  "rel_rmse_gpu_vs_ref": round(got, 5), "rel_rmse_floor_ref_vs_ref": round(floor, 5), "ratio_to_floor": round(got / floor, 4),
                  "tolerance": "<= 1.25 x floor", "within_tolerance": bool(got <= 1.25 * floor),
                  "image_mean_ratio_rgb": [round(float(x), 5) for x in ratio], "reference_nan_pixels_dropped": int(d1 + d2)}
    m = rb._profile_json("kernel_metrics.json").get("cornell_spheres_bpt", {})
    dom = m.get("connectKernel")
    roof = None
    if dom is not None:
        roof = {"bound": "issue", "achieved": dom["issue_slot_utilisation_pct"], "peak": 100.0, "unit": "% of issue slots",
                "frac": dom["issue_slot_utilisation_pct"] / 100.0, "traffic": dom.get("dram_bytes_per_launch"), "kernel": "connectKernel",
                "kernel_share_of_step": dom.get("share_of_frame"), "lanes_active_of_32": dom["lanes_active"],
                "achieved_occupancy_pct": dom.get("occupancy_pct"), "source": dom["source"],
                "note": "committed ncu capture of this workload (not re-measured in this run); the kernel evaluates two BSDFs, one "
                        "visibility ray and a MIS sum per connection: latency / issue bound on divergent lanes, far from HBM"}
    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": w, "height": h, "spp_per_gpu": spp, "samples_per_step": samples,
                       "rays_per_sample": st.rays / samples, "subpath_rays_per_sample": st.extend_rays / samples,
                       "connections_per_sample": st.class_hits[8] / samples, "subpaths_cut_at_64_vertices": int(st.tail_paths),
                       "batches_per_frame": int(st.waves), "triangles": int(hs.desc.num_triangles),
                       "l2_policy": "per-step working set (6.7 GB of vertex storage per batch of 262144 samples) is larger than L2",
                       "step_ms": {"median": round(step_ms[len(step_ms) // 2], 3), "min": round(step_ms[0], 3), "max": round(step_ms[-1], 3)}},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": samples / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(gs.device_bytes),
                    "d2h_bytes_per_step": w * h * chan * 4,
                    "breakdown": {"wall_ms_median": round(1e3 * e2e_s, 2), "device_ms": round(1e3 * hst["device_s"], 2), "scene_upload_ms": round(1e3 * hst["upload_s"], 2)}},
            "gpu_launches": int(launches), "clocks": clocks.summary()}
    if parity is not None:
        line["image_parity"] = parity
    print(json.dumps(line))
