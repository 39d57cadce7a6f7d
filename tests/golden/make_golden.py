"""Generates the golden hit records and tree digests from the COMPILED REFERENCE (oracle/_ref).

Run in the container that has /root/reference (python tests/golden/make_golden.py). The outputs are
committed so the tests can pin the restatement, the host BVH builder and the CUDA path on machines
that have neither /root/reference nor oracle/_ref. Inputs are regenerated deterministically from
tests/oracle_util.py's case functions, so only reference OUTPUTS are stored.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_util as ou  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    for name, fn in ou.CASES.items():
        meshes, placements, rays = fn()
        hits, trees, info = ou.run_ref_intersect(meshes, placements, rays)
        # QBVH and SBVH may legitimately disagree on rays that hit a shared edge/vertex at bit-equal
        # distances (different visiting order, later primitive wins): the goldens are QBVH's answers.
        out = {"prim": hits["prim"], "inst": hits["inst"], "t_bits": hits["t"].view(np.uint32),
               "u_bits": hits["u"].view(np.uint32), "v_bits": hits["v"].view(np.uint32),
               # ... and SBVH's own answers (the shipped default accelerator, SBVH::intersect), for the SBVH entry point
               "prim_sbvh": hits["prim_sbvh"], "inst_sbvh": hits["inst_sbvh"], "t_sbvh_bits": hits["t_sbvh"].view(np.uint32),
               "u_sbvh_bits": hits["u_sbvh"].view(np.uint32), "v_sbvh_bits": hits["v_sbvh"].view(np.uint32),
               "num_trees": np.array(len(trees)), "qbvh_vs_sbvh_mismatches": np.array(info["qbvh_vs_sbvh_mismatches"]),
               "ray_digest": np.array(digest(np.stack([rays[k] for k in ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")])))}
        for i, t in enumerate(trees):
            out[f"tree{i}_nodes_sha"] = np.array(digest(ou.normalise_ref_tree(t)))
            out[f"tree{i}_refs_sha"] = np.array(digest(t["refs"]))
            out[f"tree{i}_shape"] = np.array([t["nodes"].shape[0], t["refs"].shape[0]])
            out[f"tree{i}_costs"] = np.array([t["sbvh_cost"], t["qbvh_cost"]], np.float32)
        path = os.path.join(HERE, f"intersect_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "rays", len(hits["prim"]), "hits", int((hits["prim"] != 0xFFFFFFFF).sum()), "trees", len(trees),
              "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
