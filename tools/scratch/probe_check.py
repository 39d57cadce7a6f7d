import sys, json, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, render_util as ru
from slr_b200 import capi
work = tempfile.mkdtemp()
for name in ["diffuse", "spheres", "materials", "ibl", "instanced"]:
    path = ru.scene_file(name, work, 64, 64, 1)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    gs = capi.GpuScene(hs)
    c = [hs.desc.world_center[i] for i in range(3)]
    probes = ru.make_probes(c, hs.desc.world_radius, 20000, 99)
    got = capi.probe_shading(gs, probes)
    want = ru.run_ref_probe(path, probes)
    r = ru.compare_probes(got, want)
    print(name, json.dumps(r))
    np.save(f"gpurun_out/probe_{name}_gpu.npy", got); np.save(f"gpurun_out/probe_{name}_ref.npy", want); np.save(f"gpurun_out/probe_{name}_in.npy", probes)
