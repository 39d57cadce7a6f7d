#!/usr/bin/env python
"""One frame of a render workload (or one launch of the intersect workload) between cudaProfilerStart / Stop, for
`ncu --profile-from-start off`: the capture then holds exactly the launches of ONE steady-state frame.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/launches_X.csv python tools/ncu_frame.py --workload materials --spp 16
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:extendKernel -c 1 -o ... (same command)

Prints one JSON line: workload, paths of the frame, RenderStats. Not a benchmark: nothing measured under ncu is a bench value."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cornell_spheres")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--grid", type=int, default=2236)
    ap.add_argument("--rays", type=int, default=16 * 1024 * 1024)
    ap.add_argument("--warm", type=int, default=2)
    args = ap.parse_args()
    args.ref_scenes = os.path.join(ROOT, "oracle", "_ref", "TestScenes")
    # the device-driven wave loop is a conditional graph node, whose kernels ncu cannot profile one by one: take the
    # host-driven loop (same kernels, two waves per plain graph launch)
    os.environ.setdefault("SLRGPU_HOST_LOOP", "1")
    import torch
    from slr_b200 import capi, render_bench
    torch.cuda.set_device(0)
    if args.workload == "intersect":
        import numpy as np
        import oracle_util as ou
        from slr_b200 import synth
        pos, idx = synth.heightfield(args.grid)
        rays = synth.random_rays(args.rays, pos.min(0), pos.max(0), seed=12345)
        hs = ou.build_host_scene([(pos, idx)], [(0, 0, None)])
        gs = capi.GpuScene(hs)
        n = args.rays
        keys = ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")
        d_in = {k: torch.from_numpy(np.ascontiguousarray(rays[k])).cuda() for k in keys}
        d_prim = torch.empty(n, dtype=torch.int32, device="cuda")
        d_inst = torch.empty(n, dtype=torch.int32, device="cuda")
        d_t = torch.empty(n, dtype=torch.float32, device="cuda")
        rb = capi.RayBatch(*[C.cast(d_in[k].data_ptr(), capi.PF) for k in keys])
        hb = capi.HitBatch(C.cast(d_prim.data_ptr(), capi.PU32), C.cast(d_inst.data_ptr(), capi.PU32), C.cast(d_t.data_ptr(), capi.PF),
                           None, None, None, None)
        stream = torch.cuda.current_stream()

        def launch():
            rc = capi.gpu.slrgpu_intersect_batch_device(gs.handle, C.byref(rb), n, C.byref(hb), C.c_void_p(stream.cuda_stream))
            assert rc == 0, capi.gpu.slrgpu_last_error()
        for _ in range(args.warm):
            launch()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        launch()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"workload": "intersect", "rays": n, "triangles": 2 * args.grid * args.grid}))
        return
    path, w, h, spp, desc = render_bench._scene(args)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    gs = capi.GpuScene(hs)
    chan = capi.gpu.slrgpu_scene_channels(gs.handle)
    accum = torch.zeros((h, w, chan), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    params = capi.RenderParams(C.sizeof(capi.RenderParams), w, h, 0, spp, 0.0, 0.0, 1509761209, 0, args.pool, 0)

    def frame():
        st = capi.RenderStats()
        rc = capi.gpu.slrgpu_render_device(gs.handle, C.byref(params), C.c_void_p(accum.data_ptr()), C.c_void_p(stream.cuda_stream), C.byref(st))
        assert rc == 0, capi.gpu.slrgpu_last_error()
        return st
    for _ in range(args.warm):
        frame()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    st = frame()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps({"workload": args.workload, "width": w, "height": h, "spp": spp, "paths": w * h * spp, "rays": int(st.rays),
                      "waves": int(st.waves), "launches": int(st.kernel_launches), "device_ms_under_profiler": st.device_ms}))


if __name__ == "__main__":
    main()
