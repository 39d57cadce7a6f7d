"""Prints the essentials of bench.py JSON lines read from files or stdin."""
import json
import sys

for src in (sys.argv[1:] or ["-"]):
    text = sys.stdin.read() if src == "-" else open(src).read()
    for line in text.splitlines():
        if not line.startswith("{"):
            continue
        j = json.loads(line)
        c = j.get("config", {})
        print(f"{src}: {j.get('impl', 'b200')} n_gpus={j.get('n_gpus')} value={j['value']:.2f} {j['unit']} ms/step={j['ms_per_step']:.2f} "
              f"e2e={j['e2e']['value']:.2f} cpu={(j.get('cpu_baseline') or {}).get('value')} launches={j.get('gpu_launches')}")
        if "stage_ms_profiled_frame" in c:
            print("   stages:", c["stage_ms_profiled_frame"], "rays/path", round(c.get("rays_per_path", 0), 3), "Mrays/s", round(c.get("mrays_per_s", 0), 1))
            print("   traversal: extend nodes/ray", c.get("extend_nodes_per_ray"), "leafrec/ray", c.get("extend_leaf_records_per_ray"), "| shadow", c.get("shadow_nodes_per_ray"), c.get("shadow_leaf_records_per_ray"), "| pool", c.get("pool"), "waves", c.get("waves_per_frame"))
        if "roofline" in j:
            r = j["roofline"]
            print(f"   roofline: {r['kernel']} {r['achieved']:.0f}/{r['peak']:.0f} {r['unit']} frac={r['frac']:.3f}")
        if j["e2e"].get("breakdown"):
            print("   e2e:", j["e2e"]["breakdown"])
        print("   clocks:", j.get("clocks"))
