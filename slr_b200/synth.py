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
        out = {"prim": prim, "inst": inst, "t": t, "u": u, "v": v}
        rest = f.read()
        if len(rest) == 20 * n:          # the SBVH pass of the same batch (ref_intersect)
            a = np.frombuffer(rest, np.uint32).reshape(5, n)
            out.update({"prim_sbvh": a[0].copy(), "inst_sbvh": a[1].copy(), "t_sbvh": a[2].view(np.float32).copy(),
                        "u_sbvh": a[3].view(np.float32).copy(), "v_sbvh": a[4].view(np.float32).copy()})
    return out


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


# --------------------------------------------------------------------------------------------------
# .assbin writer (Assimp binary dump, the subset slr_b200/host/assets/assbin.h documents): several
# meshes and materials in one file, one child node per mesh under the root.
# --------------------------------------------------------------------------------------------------
def _chunk(magic, payload):
    return struct.pack("<II", magic, len(payload)) + payload


def _aistring(s):
    b = s.encode()
    return struct.pack("<I", len(b)) + b


def write_assbin_scene(path, meshes, materials):
    """meshes: list of dicts {name, positions[n,3], indices[m,3], normals, tangents, uvs[n,2] (optional), material}
    materials: list of dicts {name, diffuse (r,g,b) optional}."""
    header = bytearray(512)
    sig = ("ASSIMP.binary-dump.%-25s" % "slr_b200 synthetic asset").encode()[:44]
    header[:len(sig)] = sig
    header[44:60] = struct.pack("<IIII", 3, 1, 0, 0)
    for i in range(44 + 20 + 256 + 128, 512):
        header[i] = 0xCD
    ident = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    children = b""
    for i, m in enumerate(meshes):
        children += _chunk(0x123C, _aistring(m.get("name", f"mesh{i}")) + struct.pack("<16f", *ident) + struct.pack("<II", 0, 1) + struct.pack("<I", i))
    root = _chunk(0x123C, _aistring("root") + struct.pack("<16f", *ident) + struct.pack("<II", len(meshes), 0) + children)
    body = struct.pack("<7I", 0, len(meshes), len(materials), 0, 0, 0, 0) + root
    for m in meshes:
        pos = np.ascontiguousarray(m["positions"], np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(m["indices"], np.uint32).reshape(-1, 3)
        nv = pos.shape[0]
        comps = 0x1
        payload = b""
        arrays = [pos.tobytes()]
        nrm = m.get("normals")
        tng = m.get("tangents")
        uv = m.get("uvs")
        if nrm is not None:
            comps |= 0x2
            nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1, 3)
            arrays.append(nrm.tobytes())
        if tng is not None and nrm is not None:
            comps |= 0x4
            tng = np.ascontiguousarray(tng, np.float32).reshape(-1, 3)
            arrays.append(tng.tobytes())
            arrays.append(np.cross(nrm, tng).astype(np.float32).tobytes())
        if uv is not None:
            comps |= 0x100
            uv3 = np.zeros((nv, 3), np.float32)
            uv3[:, :2] = np.asarray(uv, np.float32).reshape(-1, 2)
            arrays.append(struct.pack("<I", 2) + uv3.tobytes())
        payload = struct.pack("<6I", 0x4, nv, idx.shape[0], 0, int(m.get("material", 0)), comps) + b"".join(arrays)
        if nv < 65536:
            faces = np.empty((idx.shape[0], 4), np.uint16)
            faces[:, 0] = 3
            faces[:, 1:] = idx
        else:
            faces = np.empty(idx.shape[0], np.dtype([("n", "<u2"), ("i", "<u4", 3)]))
            faces["n"] = 3
            faces["i"] = idx
        payload += faces.tobytes()
        body += _chunk(0x1237, payload)
    for mat in materials:
        props = [_chunk(0x123E, _aistring("?mat.name") + struct.pack("<IIII", 0, 0, 4 + len(mat["name"]) + 1, 3)
                        + struct.pack("<I", len(mat["name"])) + mat["name"].encode() + b"\0")]
        if mat.get("diffuse") is not None:
            props.append(_chunk(0x123E, _aistring("$clr.diffuse") + struct.pack("<IIII", 0, 0, 12, 1) + struct.pack("<3f", *mat["diffuse"])))
        body += _chunk(0x123D, struct.pack("<I", len(props)) + b"".join(props))
    with open(path, "wb") as f:
        f.write(bytes(header) + _chunk(0x1239, body))


def quad_mesh(corners, normal, tangent):
    """Two triangles (0,1,2), (0,2,3) over four corners with a constant frame and unit uvs."""
    pos = np.asarray(corners, np.float32)
    return {"positions": pos, "indices": np.array([[0, 1, 2], [0, 2, 3]], np.uint32),
            "normals": np.tile(np.asarray(normal, np.float32), (4, 1)), "tangents": np.tile(np.asarray(tangent, np.float32), (4, 1)),
            "uvs": np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)}


def displaced_sphere(segments_u=128, segments_v=64, amplitude=0.08, freq=6.0):
    """UV sphere with a smooth radial displacement (bumpy ball): the instanced base mesh of config C4."""
    pos, idx, nrm, tng, uv = uv_sphere(segments_u, segments_v)
    r = 1.0 + amplitude * np.sin(freq * pos[:, 0]) * np.cos(freq * pos[:, 1] + 0.5) * np.sin(freq * pos[:, 2] + 1.0)
    return (pos * r[:, None]).astype(np.float32), idx, nrm, tng, uv


def sky_environment(width=2048, height=1024):
    """Synthetic HDR lat-long environment (RGBA float32): analytic sky gradient, a small bright sun and two
    coloured lobes -- a fixed formula, no random numbers."""
    v, u = np.meshgrid((np.arange(height) + 0.5) / height, (np.arange(width) + 0.5) / width, indexing="ij")
    theta, phi = v * np.pi, u * 2 * np.pi
    d = np.stack([-np.sin(phi) * np.sin(theta), np.cos(theta), np.cos(phi) * np.sin(theta)], -1)
    up = np.clip(d[..., 1], 0, 1)
    sky = np.stack([0.25 + 0.35 * (1 - up), 0.35 + 0.35 * (1 - up), 0.55 + 0.35 * up], -1)
    ground = np.array([0.18, 0.15, 0.12])
    img = np.where(d[..., 1:2] > 0, sky, ground * (0.4 + 0.6 * np.clip(-d[..., 1:2], 0, 1)))

    def lobe(direction, sharp, rgb):
        dd = np.asarray(direction, np.float64)
        dd = dd / np.linalg.norm(dd)
        return np.exp(sharp * (d @ dd - 1.0))[..., None] * np.asarray(rgb)
    img = img + lobe((0.5, 0.6, 0.4), 600.0, (900.0, 800.0, 600.0))     # sun, ~5 degrees
    img = img + lobe((-0.7, 0.3, -0.5), 20.0, (1.5, 0.4, 0.2))
    img = img + lobe((0.1, 0.2, -0.9), 30.0, (0.2, 0.6, 1.8))
    rgba = np.ones((height, width, 4), np.float32)
    rgba[..., :3] = img
    return rgba
