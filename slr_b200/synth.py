"""Synthetic geometry, ray batches and the binary files exchanged with the oracle drivers.

Everything here is deterministic (fixed formulas / fixed-seed xorshift128 with the reference's
bits->[0,1) mapping, libSLR/Core/RandomNumberGenerator.cpp:12-15) so the same inputs can be
regenerated on the GPU box without shipping them.
"""
import struct
import numpy as np


# ------------------------------------------------------------------------------------------------
# RNG: xorshift128 exactly as libSLR/RNGs/XORShiftRNG.cpp:21-36, vectorised over independent streams
# ------------------------------------------------------------------------------------------------
class XorShift128:
    def __init__(self, seed, streams=1):
        s = (np.uint32(seed) + np.arange(streams, dtype=np.uint32) * np.uint32(0x9E3779B9)).astype(np.uint32)
        st = []
        with np.errstate(over="ignore"):
            for i in range(4):
                s = (np.uint32(1812433253) * (s ^ (s >> np.uint32(30))) + np.uint32(i)).astype(np.uint32)
                st.append(s.copy())
        self.a = st
        for _ in range(50):
            self.next_uint()

    def next_uint(self):
        a = self.a
        t = a[0] ^ (a[0] << np.uint32(11))
        a[0], a[1], a[2] = a[1], a[2], a[3]
        a[3] = (a[3] ^ (a[3] >> np.uint32(19))) ^ (t ^ (t >> np.uint32(8)))
        return a[3].copy()

    def next_float(self):
        bits = (self.next_uint() >> np.uint32(9)) | np.uint32(0x3F800000)
        return bits.view(np.float32) - np.float32(1.0)


def uniform_floats(seed, n, k):
    """k arrays of n uniforms in [0,1) from 4096 interleaved xorshift streams."""
    streams = 4096
    rng = XorShift128(seed, streams)
    rounds = (n + streams - 1) // streams
    out = [np.empty(rounds * streams, np.float32) for _ in range(k)]
    for r in range(rounds):
        for j in range(k):
            out[j][r * streams:(r + 1) * streams] = rng.next_float()
    return [o[:n] for o in out]


# ------------------------------------------------------------------------------------------------
# meshes
# ------------------------------------------------------------------------------------------------
def heightfield(n, amplitude=1.0):
    """(n x n) quads = 2 n^2 triangles over [0,1]^2, y = 0.08 sin 23x cos 17z + 0.02 sin(131x + 57z)."""
    g = np.linspace(0.0, 1.0, n + 1, dtype=np.float64)
    x, z = np.meshgrid(g, g, indexing="xy")
    y = amplitude * (0.08 * np.sin(23 * x) * np.cos(17 * z) + 0.02 * np.sin(131 * x + 57 * z))
    pos = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (j * (n + 1) + i).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    idx = np.empty((2 * n * n, 3), np.uint32)
    idx[0::2] = np.stack([v00, v01, v11], axis=-1)
    idx[1::2] = np.stack([v00, v11, v10], axis=-1)
    return pos, idx


def uv_sphere(segments_u=64, segments_v=32, radius=1.0):
    """UV sphere with smooth normals, tangents d/du and texcoords; poles are triangle fans."""
    pos, nrm, tng, uv = [], [], [], []
    for j in range(segments_v + 1):
        theta = np.pi * j / segments_v
        for i in range(segments_u + 1):
            phi = 2 * np.pi * i / segments_u
            n = np.array([-np.sin(phi) * np.sin(theta), np.cos(theta), np.cos(phi) * np.sin(theta)])
            pos.append(radius * n)
            nrm.append(n)
            tng.append(np.array([-np.cos(phi), 0.0, -np.sin(phi)]))
            uv.append([i / segments_u, j / segments_v])
    idx = []
    w = segments_u + 1
    for j in range(segments_v):
        for i in range(segments_u):
            a, b, c, d = j * w + i, j * w + i + 1, (j + 1) * w + i, (j + 1) * w + i + 1
            if j != 0:
                idx.append([a, b, c])
            if j != segments_v - 1:
                idx.append([b, d, c])
    return (np.asarray(pos, np.float32), np.asarray(idx, np.uint32), np.asarray(nrm, np.float32),
            np.asarray(tng, np.float32), np.asarray(uv, np.float32))


def cube(lo=(-1, -1, -1), hi=(1, 1, 1)):
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4],
                    [3, 6, 2], [3, 7, 6], [0, 4, 7], [0, 7, 3], [1, 2, 6], [1, 6, 5]], np.uint32)
    return c, idx


# ------------------------------------------------------------------------------------------------
# transforms (row-major numpy 4x4, the usual maths convention; capi transposes to column-major floats)
# ------------------------------------------------------------------------------------------------
def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = [x, y, z]
    return m


def scale(x, y=None, z=None):
    y = x if y is None else y
    z = x if z is None else z
    return np.diag(np.array([x, y, z, 1], np.float32))


def rotate_y(a):
    c, s = np.float32(np.cos(np.float32(a))), np.float32(np.sin(np.float32(a)))
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


# ------------------------------------------------------------------------------------------------
# rays
# ------------------------------------------------------------------------------------------------
def random_rays(n, lo, hi, seed=12345, inflate=0.1, tmin=0.0, tmax=np.inf):
    """Origins uniform in the bounding box inflated by `inflate`, directions uniform on the sphere."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    ext = (hi - lo) * inflate
    lo, hi = lo - ext, hi + ext
    u = uniform_floats(seed, n, 5)
    ox = (lo[0] + (hi[0] - lo[0]) * u[0]).astype(np.float32)
    oy = (lo[1] + (hi[1] - lo[1]) * u[1]).astype(np.float32)
    oz = (lo[2] + (hi[2] - lo[2]) * u[2]).astype(np.float32)
    zc = 1.0 - 2.0 * u[3].astype(np.float64)
    r = np.sqrt(np.maximum(0.0, 1.0 - zc * zc))
    ph = 2 * np.pi * u[4].astype(np.float64)
    return {"ox": ox, "oy": oy, "oz": oz,
            "dx": (r * np.cos(ph)).astype(np.float32), "dy": zc.astype(np.float32), "dz": (r * np.sin(ph)).astype(np.float32),
            "tmin": np.full(n, tmin, np.float32), "tmax": np.full(n, tmax, np.float32)}


def aimed_rays(n, lo, hi, seed=777, height=1.5):
    """Rays from a plane above the box aimed at random points inside it (high hit rate, semi-coherent)."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    u = uniform_floats(seed, n, 6)
    org = np.stack([lo[0] + (hi[0] - lo[0]) * u[0], np.full(n, hi[1] + height), lo[2] + (hi[2] - lo[2]) * u[1]], -1)
    tgt = np.stack([lo[0] + (hi[0] - lo[0]) * u[2], lo[1] + (hi[1] - lo[1]) * u[3], lo[2] + (hi[2] - lo[2]) * u[4]], -1)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return {"ox": org[:, 0].astype(np.float32), "oy": org[:, 1].astype(np.float32), "oz": org[:, 2].astype(np.float32),
            "dx": d[:, 0].astype(np.float32), "dy": d[:, 1].astype(np.float32), "dz": d[:, 2].astype(np.float32),
            "tmin": np.zeros(n, np.float32), "tmax": np.full(n, np.inf, np.float32)}


def concat_rays(*batches):
    return {k: np.concatenate([b[k] for b in batches]) for k in batches[0]}


# ------------------------------------------------------------------------------------------------
# files shared with oracle/drivers (see oracle/drivers/geom_spec.h)
# ------------------------------------------------------------------------------------------------
def write_geom_spec(path, meshes, placements):
    """meshes: list of (positions, indices); placements: list of (mesh, mode, 4x4 row-major matrix)."""
    with open(path, "wb") as f:
        f.write(b"SLRG")
        f.write(struct.pack("<II", 1, len(meshes)))
        for pos, idx in meshes:
            pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
            idx = np.ascontiguousarray(idx, np.uint32).reshape(-1, 3)
            f.write(struct.pack("<II", pos.shape[0], idx.shape[0]))
            f.write(pos.tobytes())
            f.write(idx.tobytes())
        f.write(struct.pack("<I", len(placements)))
        for mesh, mode, m in placements:
            f.write(struct.pack("<II", mesh, mode))
            m = np.eye(4, dtype=np.float32) if m is None else np.asarray(m, np.float32)
            f.write(np.ascontiguousarray(m.T).tobytes())


def write_rays(path, rays):
    n = rays["ox"].shape[0]
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", n))
        for k in ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax"):
            f.write(np.ascontiguousarray(rays[k], np.float32).tobytes())


def read_hits(path):
    with open(path, "rb") as f:
        n = struct.unpack("<Q", f.read(8))[0]
        prim = np.frombuffer(f.read(4 * n), np.uint32)
        inst = np.frombuffer(f.read(4 * n), np.uint32)
        t = np.frombuffer(f.read(4 * n), np.float32)
        u = np.frombuffer(f.read(4 * n), np.float32)
        v = np.frombuffer(f.read(4 * n), np.float32)
    return {"prim": prim, "inst": inst, "t": t, "u": u, "v": v}


def read_trees(path):
    """Trees dumped by ref_intersect: list of dicts(nodes (n,32) uint32 view, refs, sbvh_cost, qbvh_cost)."""
    out = []
    with open(path, "rb") as f:
        na = struct.unpack("<I", f.read(4))[0]
        for _ in range(na):
            nn = struct.unpack("<I", f.read(4))[0]
            nodes = np.frombuffer(f.read(128 * nn), np.uint32).reshape(nn, 32)
            nr = struct.unpack("<I", f.read(4))[0]
            refs = np.frombuffer(f.read(4 * nr), np.uint32)
            costs = np.frombuffer(f.read(8), np.float32)
            out.append({"nodes": nodes, "refs": refs, "sbvh_cost": float(costs[0]), "qbvh_cost": float(costs[1])})
    return out
