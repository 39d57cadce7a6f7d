#!/usr/bin/env python
"""bench.py -- throughput of the SLR hot path on B200 next to the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic input. Prints ONE JSON line
(rank 0). See DESIGN.md "Measurement" for the definition of every field.

Workloads
  cornell_spheres (default)  the path tracer on BASELINE.json configs[0], Cornell_Box_Spheres 512x512 64 spp
              spectral -- the configuration north_star's >= 100x target is quoted on; metric Mpaths/s
              (slr_b200/render_bench.py).
  materials | ibl | instanced   BASELINE configs[1..3] through the same code (slr_b200/render_bench.py WORKLOADS).
  cornell_spheres_bpt   the same scene through the GPU twin of the reference's BidirectionalPathTracingRenderer -- the renderer
              the shipped scene file selects -- next to the reference's own (slr_b200/bpt_bench.py); metric Msamples/s.
  intersect   incoherent-ray closest-hit microbench (BASELINE.json configs[4] at a single-GPU size):
              heightfield triangle mesh -> host SBVH -> QBVH, random rays; metric Mrays/s.

Timing: CUDA events on the launching stream, W >= 3 warm-up steps, barrier + synchronize on both
sides, max over ranks. The ray batch (>= 512 MB of SoA inputs+outputs per step at the default size)
is larger than L2 (126 MB), so no explicit L2 flush is needed between steps; config says so.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# --------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md "clocks line")
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (in-process: no
    nvidia-smi process start-up contending with the render loop's stream polling); falls back to
    `nvidia-smi --query-gpu` once per second when the NVML binding is missing."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.1):
        self.index = index
        self.period = period
        self.samples = []          # (sm_mhz, max_mhz, [bool x 4])
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
        bits = [0x8, 0x40, 0x20, 0x4]     # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap
        self.samples.append((float(sm), float(mx), [bool(r & b) for b in bits]))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 6:
            self.samples.append((float(parts[0]), float(parts[1]), [p.lower().startswith("active") for p in parts[2:6]]))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(self.period if self._nvml is not None else 1.0)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2][i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
# intersect workload
# --------------------------------------------------------------------------------------------------
def intersect_inputs(args, rank):
    from slr_b200 import synth
    pos, idx = synth.heightfield(args.grid)
    rays = synth.random_rays(args.rays, pos.min(0), pos.max(0), seed=12345 + 1000 * rank)
    return pos, idx, rays


def workload_name(args):
    return (f"intersect: {args.rays} incoherent rays vs {2 * args.grid * args.grid}-triangle heightfield "
            f"(host SBVH->QBVH), closest hit")


def run_intersect_gpu(args, rank, world, dist):
    import torch
    from slr_b200 import capi
    import oracle_util as ou

    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    pos, idx, rays = intersect_inputs(args, rank)
    hs = ou.build_host_scene([(pos, idx)], [(0, 0, None)])
    gs = capi.GpuScene(hs, device=dev)
    n = args.rays
    keys = ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")

    # ---- device-resident inputs (the `value` leg)
    host_pinned = {k: torch.from_numpy(np.ascontiguousarray(rays[k])).pin_memory() for k in keys}
    d_in = {k: host_pinned[k].cuda(non_blocking=True) for k in keys}
    d_prim = torch.empty(n, dtype=torch.int32, device="cuda")
    d_inst = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    d_nodes = torch.empty(n, dtype=torch.int32, device="cuda")
    d_tris = torch.empty(n, dtype=torch.int32, device="cuda")
    rb = capi.RayBatch(*[C.cast(d_in[k].data_ptr(), capi.PF) for k in keys])
    hb = capi.HitBatch(C.cast(d_prim.data_ptr(), capi.PU32), C.cast(d_inst.data_ptr(), capi.PU32),
                       C.cast(d_t.data_ptr(), capi.PF), None, None, None, None)
    hb_cnt = capi.HitBatch(C.cast(d_prim.data_ptr(), capi.PU32), C.cast(d_inst.data_ptr(), capi.PU32),
                           C.cast(d_t.data_ptr(), capi.PF), None, None,
                           C.cast(d_nodes.data_ptr(), capi.PU32), C.cast(d_tris.data_ptr(), capi.PU32))
    stream = torch.cuda.current_stream()

    def launch(h):
        rc = capi.gpu.slrgpu_intersect_batch_device(gs.handle, C.byref(rb), n, C.byref(h), C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(capi.gpu.slrgpu_last_error().decode())

    # algorithmic bytes per ray, counted by the oracle-order traversal (one counting launch, untimed)
    launch(hb_cnt)
    torch.cuda.synchronize()
    tot_nodes = int(d_nodes.to(torch.int64).sum().item())
    tot_tris = int(d_tris.to(torch.int64).sum().item())
    hit_rate = float((d_prim != -1).float().mean().item())
    algo_bytes = 32 * n + 128 * tot_nodes + 48 * tot_tris + 16 * n

    for _ in range(args.warmup):
        launch(hb)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            launch(hb)
        e1.record(stream)
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)

    # ---- end to end through the host-buffer C-ABI call (H2D + kernel + D2H inside the timed region)
    comps = [np.ascontiguousarray(rays[k], np.float32) for k in keys]
    rb_h = capi.RayBatch(*[c.ctypes.data_as(capi.PF) for c in comps])
    o_prim, o_inst, o_t = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.float32)
    hb_h = capi.HitBatch(o_prim.ctypes.data_as(capi.PU32), o_inst.ctypes.data_as(capi.PU32), o_t.ctypes.data_as(capi.PF),
                         None, None, None, None)
    e2e_steps = max(1, min(args.steps, 3))
    capi.gpu.slrgpu_intersect_batch(gs.handle, C.byref(rb_h), n, C.byref(hb_h), None)   # warm
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = capi.gpu.slrgpu_intersect_batch(gs.handle, C.byref(rb_h), n, C.byref(hb_h), None)
        assert rc == 0, capi.gpu.slrgpu_last_error()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    return {"ms_total": ms, "units": n * args.steps, "algo_bytes_per_launch": algo_bytes, "launches": args.steps,
            "nodes_per_ray": tot_nodes / n, "tris_per_ray": tot_tris / n, "hit_rate": hit_rate,
            "clocks": clocks.summary(), "e2e_units_per_s": n / e2e_s, "h2d": 32 * n, "d2h": 12 * n,
            "scene_bytes": gs.device_bytes, "host_scene": hs, "rays": rays, "build_s": hs.build_seconds, "gpu_scene": gs}


def cpu_baseline_intersect(args, sample_rays, kind_pref="reference", gpu_scene=None):
    """The reference's own QBVH::intersect on the box's host cores (oracle/_ref/ref_intersect) on a
    bounded sample of the same workload; falls back to the scalar restatement (1 core)."""
    import oracle_util as ou
    from slr_b200 import synth
    pos, idx = synth.heightfield(args.grid)
    rays = synth.random_rays(sample_rays, pos.min(0), pos.max(0), seed=12345)
    if kind_pref == "reference" and ou.have_ref():
        ref_hits, _, info = ou.run_ref_intersect([(pos, idx)], [(0, 0, None)], rays, want_trees=False)
        best = min(info["qbvh_1t_s"], info["qbvh_nt_s"])
        cores = 1 if info["qbvh_1t_s"] <= info["qbvh_nt_s"] else info["threads"]
        parity = None
        if gpu_scene is not None:
            # the same sample through the CUDA path: hit ids against the reference's own QBVH on this box
            got = gpu_scene.intersect(rays)
            hit = ref_hits["prim"] != 0xFFFFFFFF
            parity = {"rays": int(sample_rays), "hits": int(hit.sum()),
                      "prim_ids_equal": bool(np.array_equal(got["prim"], ref_hits["prim"])),
                      "t_bit_equal": bool(np.array_equal(got["t"].view(np.uint32)[hit], ref_hits["t"].view(np.uint32)[hit]))}
        return {"value": sample_rays / best / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference", "parity_vs_gpu": parity,
                "reference_build_s": round(info.get("sbvh_build_s", 0.0) + info.get("qbvh_build_s", 0.0), 2),
                "sample": f"{sample_rays} rays of the same batch through QBVH::intersect; best of 1 thread "
                          f"({sample_rays / info['qbvh_1t_s'] / 1e6:.3f}) and {info['threads']} threads "
                          f"({sample_rays / info['qbvh_nt_s'] / 1e6:.3f} Mrays/s)",
                "threads_available": info["threads"]}
    hs = ou.build_host_scene([(pos, idx)], [(0, 0, None)])
    t0 = time.perf_counter()
    ou.restate_intersect(hs, rays)
    dt = time.perf_counter() - t0
    return {"value": sample_rays / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
            "sample": f"{sample_rays} rays of the same batch through the scalar C restatement"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--grid", type=int, default=500, help="heightfield resolution (2*grid^2 triangles)")
    ap.add_argument("--rays", type=int, default=16 * 1024 * 1024)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--size", type=int, default=0, help="render workloads: override the image size")
    ap.add_argument("--spp", type=int, default=0, help="render workloads: override the samples per pixel (per GPU)")
    ap.add_argument("--pool", type=int, default=0, help="render workloads: paths in flight (0 = library default)")
    ap.add_argument("--parity-paths", type=float, default=40e6,
                    help="render workloads: largest frame (paths) whose image is also compared with two seeds of the reference (image_parity)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="render workloads, N > 1: weak = --spp samples per GPU (frame = spp x N, the default the driver measures); "
                         "strong = --spp is the frame's sample count, partitioned over the GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # the reference's shipped scene files, where the oracle build put them (git-ignored, travels to the GPU box): C1 renders
    # TestScenes/Cornell_Box_Spheres.txt unchanged when present
    args.ref_scenes = os.path.join(ROOT, "oracle", "_ref", "TestScenes")

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    try:
        from slr_b200 import render_bench      # present once the renderer is built
    except ImportError:
        render_bench = None
    if args.workload is None:
        args.workload = "cornell_spheres" if render_bench is not None else "intersect"
    if args.workload == "cornell_spheres_bpt":
        from slr_b200 import bpt_bench
        return bpt_bench.main(args, rank, world)
    if args.workload != "intersect":
        if render_bench is None:
            raise SystemExit(f"workload {args.workload} needs the renderer")
        return render_bench.main(args, rank, world)

    if args.impl == "reference":
        if rank != 0:
            return
        t0 = time.perf_counter()
        vals = []
        for _ in range(max(1, args.steps)):
            vals.append(cpu_baseline_intersect(args, args.cpu_sample))
        best = max(vals, key=lambda v: v["value"])
        wall = time.perf_counter() - t0
        line = {"impl": "reference", "metric": "Mrays/s (closest-hit, incoherent rays)", "value": best["value"],
                "unit": "Mrays/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args)}, "cpu_baseline": best,
                "e2e": {"value": best["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    r = run_intersect_gpu(args, rank, world, dist)
    ms = r["ms_total"]
    e2e = r["e2e_units_per_s"]
    if dist is not None:
        import torch
        t = torch.tensor([ms, 1.0 / e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e = float(t[0]), 1.0 / float(t[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    value = r["units"] * world / (ms * 1e-3) / 1e6
    ach = r["algo_bytes_per_launch"] / (ms * 1e-3 / r["launches"]) / 1e9
    cpu = cpu_baseline_intersect(args, args.cpu_sample, gpu_scene=r["gpu_scene"])
    # measured DRAM bytes per ray and the issue / lane figures of the committed ncu capture of this mesh size, if there is one
    traffic, km = None, None
    try:
        key = f"intersect_grid{args.grid}"
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(key, {}).get("intersectBatchKernel")
        if t is not None:
            traffic = t["dram_bytes_per_ray"] * args.rays
        with open(os.path.join(ROOT, "profiles", "kernel_metrics.json")) as f:
            km = json.load(f).get(key, {}).get("intersectBatchKernel")
    except (OSError, ValueError, KeyError):
        pass
    line = {"metric": "Mrays/s (closest-hit, incoherent rays)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "rays_per_gpu_per_step": args.rays,
                       "triangles": 2 * args.grid * args.grid, "hit_rate": round(r["hit_rate"], 4),
                       "nodes_per_ray": round(r["nodes_per_ray"], 3), "tris_per_ray": round(r["tris_per_ray"], 3),
                       "l2_policy": "inputs larger than L2 (ray SoA + results >= 44 B/ray x rays)",
                       "scene_bytes": int(r["scene_bytes"]), "host_bvh_build_s": round(r["build_s"], 2)},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "kernel": "intersectBatchKernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_ray": r["algo_bytes_per_launch"] / args.rays,
                         "lanes_active_of_32": km["lanes_active"] if km else None,
                         "issue_slot_utilisation_pct": km["issue_slot_utilisation_pct"] if km else None,
                         "note": "achieved = algorithmic bytes (32 B ray + 128 B per node popped + 48 B per leaf record tested + 16 B hit) of one "
                                 "launch / its device time; traffic = ncu dram__bytes of one launch of this mesh size (profiles/traffic.json), "
                                 "bytes per launch -- below the algorithmic figure because the upper tree levels are served by L2. The kernel is "
                                 "bound by the LATENCY of its dependent node fetches (ncu: long-scoreboard stalls dominate, DRAM throughput "
                                 "14 %), see profiles/r02_ncu_c4_c5_final.md"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e * world / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": args.steps, "clocks": r["clocks"]}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
