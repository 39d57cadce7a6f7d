"""One small bidirectional render for ncu: python tools/bpt_profile.py [scene] [size] [spp]"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402
from slr_b200 import capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "diffuse"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 64
work = tempfile.mkdtemp(prefix="bpt_prof_")
path = ru.scene_file(name, work, size, size, spp)
with capi.stdout_to_stderr():
    hs = capi.read_scene(path)
gs = capi.GpuScene(hs)
for _ in range(2):
    accum, st = capi.gpu_render(gs, size, size, 0, spp, flags=capi.RENDER_BPT)
print(name, size, spp, "device ms", st["device_ms"], "rays", st["rays"], "connections", st["class_hits"][8], "truncated", st["tail_paths"])
