"""ctypes bindings of include/slrgpu.h and include/slrhost.h.

Mirrors the C structs field for field; ``check_abi()`` compares every ``ctypes.sizeof`` with the
library's ``slrgpu_struct_size`` so a drifted mirror fails loudly instead of corrupting memory.
"""
import contextlib
import ctypes as C
import os
import sys
import time
import numpy as np

_LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
INVALID_ID = 0xFFFFFFFF

c_f = C.c_float
c_u32 = C.c_uint32
c_u64 = C.c_uint64
PF = C.POINTER(c_f)
PU32 = C.POINTER(c_u32)
PU8 = C.POINTER(C.c_uint8)


class BvhNode(C.Structure):
    _fields_ = [("lo_x", c_f * 4), ("lo_y", c_f * 4), ("lo_z", c_f * 4),
                ("hi_x", c_f * 4), ("hi_y", c_f * 4), ("hi_z", c_f * 4),
                ("child", c_u32 * 4), ("top_axis", C.c_uint8), ("left_axis", C.c_uint8),
                ("right_axis", C.c_uint8), ("pad0", C.c_uint8), ("pad", c_u32 * 3)]


class LeafRecord(C.Structure):
    _fields_ = [("a", c_f * 4), ("b", c_f * 4), ("c", c_f * 4)]


class Instance(C.Structure):
    _fields_ = [("mat", c_f * 16), ("mat_inv", c_f * 16), ("root_node", c_u32),
                ("light_base", c_u32), ("num_lights", c_u32), ("light_index", c_u32),
                ("light_importance", c_f), ("sbvh_root_node", c_u32), ("motion", c_u32), ("pad", c_u32)]


class Triangle(C.Structure):
    _fields_ = [("v", c_u32 * 3), ("material", c_u32), ("normal_map", c_u32), ("alpha_map", c_u32),
                ("light_index", c_u32), ("pad", c_u32)]


class Vertex(C.Structure):
    _fields_ = [("position", c_f * 3), ("u", c_f), ("normal", c_f * 3), ("v", c_f),
                ("tangent", c_f * 3), ("pad", c_f)]


class Spectrum(C.Structure):
    _fields_ = [("kind", c_u32), ("data_offset", c_u32), ("num_samples", c_u32),
                ("p0", c_f), ("p1", c_f), ("p2", c_f), ("pad", c_u32 * 2)]


class Texture(C.Structure):
    _fields_ = [("kind", c_u32), ("mapping", c_u32), ("i0", c_u32), ("i1", c_u32),
                ("f0", c_f), ("f1", c_f), ("f2", c_f), ("f3", c_f),
                ("map_offset", c_f * 2), ("map_scale", c_f * 2)]


class SbvhNode(C.Structure):
    _fields_ = [("lo", c_f * 3), ("hi", c_f * 3), ("a", c_u32), ("b", c_u32)]


class Motion(C.Structure):
    _fields_ = [("mat_end", c_f * 16), ("mat_end_inv", c_f * 16), ("T0", c_f * 3), ("t_begin", c_f), ("T1", c_f * 3), ("t_end", c_f),
                ("R0", c_f * 4), ("R1", c_f * 4), ("S0", c_f * 16), ("S1", c_f * 16)]


class Image(C.Structure):
    _fields_ = [("format", c_u32), ("width", c_u32), ("height", c_u32), ("pad", c_u32),
                ("data_offset", c_u64), ("spectrum_type", c_u32), ("pad1", c_u32)]


class Material(C.Structure):
    _fields_ = [("kind", c_u32), ("tex", c_u32 * 4), ("sub", c_u32 * 2), ("f0", c_f)]


class Light(C.Structure):
    _fields_ = [("object", c_u32), ("importance", c_f), ("pmf", c_f), ("cdf_lo", c_f), ("cdf_hi", c_f), ("pad", c_u32 * 3)]


class Camera(C.Structure):
    _fields_ = [("mat", c_f * 16), ("mat_inv", c_f * 16), ("sensitivity", c_f), ("aspect", c_f),
                ("fov_y", c_f), ("lens_radius", c_f), ("img_plane_dist", c_f), ("obj_plane_dist", c_f),
                ("pad", c_f * 2)]


class Environment(C.Structure):
    _fields_ = [("present", c_u32), ("material", c_u32), ("map_width", c_u32), ("map_height", c_u32),
                ("row_pdf", PF), ("row_cdf", PF), ("row_integral", PF), ("marginal_pdf", PF),
                ("marginal_cdf", PF), ("marginal_integral", c_f), ("pad", c_f)]


class SpectralTables(C.Structure):
    _fields_ = [("upsample_grid", PF), ("upsample_grid_floats", c_u32),
                ("upsample_points", PF), ("upsample_points_floats", c_u32),
                ("xbar_16", PF), ("ybar_16", PF), ("zbar_16", PF), ("integral_cmf", c_f), ("pad", c_f)]


class SceneDesc(C.Structure):
    _fields_ = [("struct_size", c_u32), ("rgb_mode", c_u32),
                ("bvh_nodes", C.POINTER(BvhNode)), ("num_bvh_nodes", c_u32),
                ("leaf_records", C.POINTER(LeafRecord)), ("num_leaf_records", c_u32),
                ("instances", C.POINTER(Instance)), ("num_instances", c_u32),
                ("triangles", C.POINTER(Triangle)), ("num_triangles", c_u32),
                ("vertices", C.POINTER(Vertex)), ("num_vertices", c_u32),
                ("materials", C.POINTER(Material)), ("num_materials", c_u32),
                ("textures", C.POINTER(Texture)), ("num_textures", c_u32),
                ("spectra", C.POINTER(Spectrum)), ("num_spectra", c_u32),
                ("spectrum_data", PF), ("num_spectrum_floats", c_u32),
                ("images", C.POINTER(Image)), ("num_images", c_u32),
                ("image_data", PU8), ("image_data_bytes", c_u64),
                ("lights", C.POINTER(Light)), ("num_lights", c_u32),
                ("num_top_lights", c_u32), ("top_light_importance", c_f),
                ("world_center", c_f * 3), ("world_radius", c_f),
                ("camera", Camera), ("environment", Environment), ("spectral", SpectralTables),
                ("sbvh_nodes", C.POINTER(SbvhNode)), ("num_sbvh_nodes", c_u32),
                ("sbvh_leaf_records", C.POINTER(LeafRecord)), ("num_sbvh_leaf_records", c_u32),
                ("motions", C.POINTER(Motion)), ("num_motions", c_u32), ("camera_motion", c_u32), ("pad_motion", c_u32)]


class RayBatch(C.Structure):
    _fields_ = [(n, PF) for n in ("org_x", "org_y", "org_z", "dir_x", "dir_y", "dir_z", "tmin", "tmax")]


class HitBatch(C.Structure):
    _fields_ = [("prim", PU32), ("inst", PU32), ("t", PF), ("u", PF), ("v", PF),
                ("nodes_visited", PU32), ("tris_tested", PU32)]


class RenderParams(C.Structure):
    _fields_ = [("struct_size", c_u32), ("width", c_u32), ("height", c_u32),
                ("spp_begin", c_u32), ("spp_end", c_u32), ("time_start", c_f), ("time_end", c_f),
                ("rng_seed", C.c_int32), ("max_path_length", c_u32), ("pool_size", c_u32), ("flags", c_u32)]


class RenderStats(C.Structure):
    _fields_ = [("paths", c_u64), ("rays", c_u64), ("extend_rays", c_u64), ("shadow_rays", c_u64),
                ("kernel_launches", c_u64), ("device_ms", c_f), ("raygen_ms", c_f), ("extend_ms", c_f), ("surface_ms", c_f), ("material_ms", c_f),
                ("shadow_ms", c_f), ("other_ms", c_f), ("waves", c_u64), ("extend_nodes", c_u64),
                ("extend_leaf_records", c_u64), ("shadow_nodes", c_u64), ("shadow_leaf_records", c_u64),
                ("class_hits", c_u64 * 9), ("tail_paths", c_u64), ("tail_waves", c_u64), ("tail_ms", c_f), ("reserved0", c_f)]


_ABI_STRUCTS = [SceneDesc, BvhNode, LeafRecord, Instance, Triangle, Vertex, Spectrum, Texture, Image,
                Material, Light, Camera, Environment, SpectralTables, RayBatch, HitBatch, RenderParams,
                RenderStats, SbvhNode, Motion]

RENDER_PROFILE_STAGES = 0x1
RENDER_BPT = 0x2                 # bidirectional path tracing (include/slrgpu.h SLRGPU_RENDER_BPT)


def _load(name):
    path = os.path.join(_LIBDIR, name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the native libraries first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no Python/CPU fallback for the SLR hot path.")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


gpu = _load(os.environ.get("SLRGPU_LIB", "libslrgpu.so"))      # SLRGPU_LIB: tuning variants built by csrc/Makefile VARIANT=
host = _load("libslrhost.so")

# ---- slrgpu.h prototypes
gpu.slrgpu_device_count.restype = C.c_int
gpu.slrgpu_abi_version.restype = c_u32
gpu.slrgpu_last_error.restype = C.c_char_p
gpu.slrgpu_struct_size.restype = c_u32
gpu.slrgpu_struct_size.argtypes = [C.c_int]
gpu.slrgpu_scene_create.restype = C.c_int
gpu.slrgpu_scene_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.POINTER(C.c_void_p)]
gpu.slrgpu_scene_destroy.restype = None
gpu.slrgpu_scene_destroy.argtypes = [C.c_void_p]
gpu.slrgpu_scene_device_bytes.restype = c_u64
gpu.slrgpu_scene_device_bytes.argtypes = [C.c_void_p]
gpu.slrgpu_scene_channels.restype = c_u32
gpu.slrgpu_scene_channels.argtypes = [C.c_void_p]
gpu.slrgpu_intersect_batch.restype = C.c_int
gpu.slrgpu_intersect_batch.argtypes = [C.c_void_p, C.POINTER(RayBatch), c_u64, C.POINTER(HitBatch), PF]
gpu.slrgpu_intersect_batch_device.restype = C.c_int
gpu.slrgpu_intersect_batch_device.argtypes = [C.c_void_p, C.POINTER(RayBatch), c_u64, C.POINTER(HitBatch), C.c_void_p]
gpu.slrgpu_intersect_batch_sbvh.restype = C.c_int
gpu.slrgpu_intersect_batch_sbvh.argtypes = [C.c_void_p, C.POINTER(RayBatch), c_u64, C.POINTER(HitBatch), PF]
gpu.slrgpu_render_multi.restype = C.c_int
gpu.slrgpu_render_multi.argtypes = [C.POINTER(C.c_void_p), c_u32, C.POINTER(RenderParams), PF, C.POINTER(RenderStats)]
gpu.slrgpu_scene_poll_overflow.restype = C.c_int
gpu.slrgpu_scene_poll_overflow.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
gpu.slrgpu_intersect_launch_config.restype = C.c_int
gpu.slrgpu_intersect_launch_config.argtypes = [C.c_void_p, c_u64, PU32, PU32]
gpu.slrgpu_occluded_batch.restype = C.c_int
gpu.slrgpu_occluded_batch.argtypes = [C.c_void_p, C.POINTER(RayBatch), c_u64, PU8, PF]

gpu.slrgpu_probe_shading.restype = C.c_int
gpu.slrgpu_probe_shading.argtypes = [C.c_void_p, PF, c_u64, PF]
gpu.slrgpu_probe_shading_bpt.restype = C.c_int
gpu.slrgpu_probe_shading_bpt.argtypes = [C.c_void_p, PF, c_u64, PF]
gpu.slrgpu_release_workspaces.restype = None

# ---- slrhost.h prototypes
host.slrhost_last_error.restype = C.c_char_p
host.slrhost_builder_create.restype = C.c_void_p
host.slrhost_builder_destroy.argtypes = [C.c_void_p]
host.slrhost_builder_destroy.restype = None
host.slrhost_builder_add_mesh.restype = C.c_int
host.slrhost_builder_add_mesh.argtypes = [C.c_void_p, PF, PF, PF, PF, c_u32, PU32, c_u32]
host.slrhost_builder_place_mesh.restype = C.c_int
host.slrhost_builder_place_mesh.argtypes = [C.c_void_p, C.c_int, PF]
host.slrhost_builder_instance_mesh.restype = C.c_int
host.slrhost_builder_instance_mesh.argtypes = [C.c_void_p, C.c_int, PF]
host.slrhost_builder_finish.restype = C.c_int
host.slrhost_builder_finish.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
host.slrhost_scene_destroy.restype = None
host.slrhost_scene_destroy.argtypes = [C.c_void_p]
host.slrhost_scene_describe.restype = C.c_int
host.slrhost_scene_describe.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
host.slrhost_scene_stats.restype = C.c_int
host.slrhost_scene_stats.argtypes = [C.c_void_p, c_u32, C.POINTER(C.c_double)]
host.slrhost_scene_build_seconds.restype = C.c_double
host.slrhost_scene_build_seconds.argtypes = [C.c_void_p]


@contextlib.contextmanager
def stdout_to_stderr():
    """The host library reports model loading on stdout like the reference does ("Reading: ... done.");
    programs whose stdout is a protocol (bench.py prints ONE JSON line) route it to stderr."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class SlrError(RuntimeError):
    pass


def check_abi():
    """Every ctypes mirror must have the size the library was compiled with."""
    for i, s in enumerate(_ABI_STRUCTS):
        want = gpu.slrgpu_struct_size(i)
        if want != C.sizeof(s):
            raise SlrError(f"ABI mismatch: {s.__name__} is {C.sizeof(s)} bytes in ctypes, {want} in libslrgpu.so")
    return True


def _gpu_check(rc, what):
    if rc != 0:
        raise SlrError(f"{what} failed ({rc}): {gpu.slrgpu_last_error().decode()}")


def _host_check(rc, what):
    if rc < 0:
        raise SlrError(f"{what} failed: {host.slrhost_last_error().decode()}")
    return rc


def _pf(a):
    return a.ctypes.data_as(PF) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class HostScene:
    """A flattened scene owned by libslrhost (SoA buffers + build statistics)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        self.desc = SceneDesc()
        _host_check(host.slrhost_scene_describe(self._h, C.byref(self.desc)), "slrhost_scene_describe")

    def close(self):
        if self._h:
            host.slrhost_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def handle(self):
        return self._h

    def stats(self):
        out = []
        buf = (C.c_double * 10)()
        n = _host_check(host.slrhost_scene_stats(self._h, 0, buf), "slrhost_scene_stats")
        keys = ["numObjects", "sbvhNodes", "sbvhRefs", "sbvhDepth", "qbvhNodes", "qbvhDepth", "nodeBase", "leafBase",
                "sbvhCost", "qbvhCost"]
        for a in range(n):
            host.slrhost_scene_stats(self._h, a, buf)
            out.append({k: (buf[i] if i >= 8 else int(buf[i])) for i, k in enumerate(keys)})
        return out

    @property
    def build_seconds(self):
        return host.slrhost_scene_build_seconds(self._h)

    def nodes_array(self):
        """numpy view (copy) of the flattened QBVH nodes as a structured (n, 32) uint32 array."""
        n = self.desc.num_bvh_nodes
        return np.ctypeslib.as_array(C.cast(self.desc.bvh_nodes, PU32), shape=(n, 32)).copy()

    def leaves_array(self):
        n = self.desc.num_leaf_records
        return np.ctypeslib.as_array(C.cast(self.desc.leaf_records, PU32), shape=(n, 12)).copy()

    def lights_array(self):
        """The light lists (top-level entries first) as a structured array with the fields of SlrGpuLight."""
        n = self.desc.num_lights
        dt = np.dtype([("object", np.uint32), ("importance", np.float32), ("pmf", np.float32), ("cdf_lo", np.float32),
                       ("cdf_hi", np.float32), ("pad", np.uint32, (3,))])
        assert dt.itemsize == C.sizeof(Light)
        raw = np.ctypeslib.as_array(C.cast(self.desc.lights, C.POINTER(C.c_uint8)), shape=(n * dt.itemsize,)).copy()
        return raw.view(dt)


class SceneBuilder:
    """Programmatic scene graph (the subset of the scene language's builtins needed for geometry)."""

    def __init__(self):
        self._b = C.c_void_p(host.slrhost_builder_create())
        if not self._b:
            raise SlrError("slrhost_builder_create failed")

    def close(self):
        if self._b:
            host.slrhost_builder_destroy(self._b)
            self._b = None

    def __del__(self):
        self.close()

    def add_mesh(self, positions, indices, normals=None, tangents=None, uvs=None):
        pos = _f32(positions).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        nrm = _f32(normals).reshape(-1, 3) if normals is not None else None
        tng = _f32(tangents).reshape(-1, 3) if tangents is not None else None
        uv = _f32(uvs).reshape(-1, 2) if uvs is not None else None
        return _host_check(host.slrhost_builder_add_mesh(self._b, _pf(pos), _pf(nrm), _pf(tng), _pf(uv), pos.shape[0],
                                                          idx.ctypes.data_as(PU32), idx.shape[0]), "slrhost_builder_add_mesh")

    @staticmethod
    def _mat(m):
        if m is None:
            return None, None
        a = _f32(np.asarray(m, dtype=np.float32).T.reshape(16))  # numpy row-major 4x4 -> column-major floats
        return a, _pf(a)

    def place_mesh(self, mesh, matrix=None):
        keep, p = self._mat(matrix)
        _host_check(host.slrhost_builder_place_mesh(self._b, mesh, p), "slrhost_builder_place_mesh")

    def instance_mesh(self, mesh, matrix=None):
        keep, p = self._mat(matrix)
        _host_check(host.slrhost_builder_instance_mesh(self._b, mesh, p), "slrhost_builder_instance_mesh")

    def finish(self, rgb_mode=False):
        out = C.c_void_p()
        _host_check(host.slrhost_builder_finish(self._b, 1 if rgb_mode else 0, C.byref(out)), "slrhost_builder_finish")
        return HostScene(out.value)


class GpuScene:
    """A scene resident in HBM (slrgpu_scene_create)."""

    def __init__(self, host_scene, device=0):
        self._host = host_scene            # keeps the SoA buffers alive
        self._s = C.c_void_p()
        _gpu_check(gpu.slrgpu_scene_create(C.byref(host_scene.desc), device, C.byref(self._s)), "slrgpu_scene_create")

    def close(self):
        if self._s:
            gpu.slrgpu_scene_destroy(self._s)
            self._s = None

    def __del__(self):
        self.close()

    @property
    def handle(self):
        return self._s

    @property
    def device_bytes(self):
        return gpu.slrgpu_scene_device_bytes(self._s)

    def intersect(self, rays, counters=False):
        """rays: dict of 8 float32 arrays (ox oy oz dx dy dz tmin tmax). Returns dict of result arrays."""
        comps = [_f32(rays[k]) for k in ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")]
        n = comps[0].shape[0]
        rb = RayBatch(*[_pf(c) for c in comps])
        out = {"prim": np.empty(n, np.uint32), "inst": np.empty(n, np.uint32), "t": np.empty(n, np.float32),
               "u": np.empty(n, np.float32), "v": np.empty(n, np.float32)}
        hb = HitBatch(out["prim"].ctypes.data_as(PU32), out["inst"].ctypes.data_as(PU32), _pf(out["t"]), _pf(out["u"]),
                      _pf(out["v"]), None, None)
        if counters:
            out["nodes"] = np.empty(n, np.uint32)
            out["tris"] = np.empty(n, np.uint32)
            hb.nodes_visited = out["nodes"].ctypes.data_as(PU32)
            hb.tris_tested = out["tris"].ctypes.data_as(PU32)
        ms = c_f(0)
        _gpu_check(gpu.slrgpu_intersect_batch(self._s, C.byref(rb), n, C.byref(hb), C.byref(ms)), "slrgpu_intersect_batch")
        out["kernel_ms"] = ms.value
        return out

    def intersect_sbvh(self, rays):
        """slrgpu_intersect_batch_sbvh: closest hits through the binary SBVH (the scene must carry it: set_option export_sbvh)."""
        comps = [_f32(rays[k]) for k in ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")]
        n = comps[0].shape[0]
        rb = RayBatch(*[_pf(c) for c in comps])
        out = {"prim": np.empty(n, np.uint32), "inst": np.empty(n, np.uint32), "t": np.empty(n, np.float32),
               "u": np.empty(n, np.float32), "v": np.empty(n, np.float32)}
        hb = HitBatch(out["prim"].ctypes.data_as(PU32), out["inst"].ctypes.data_as(PU32), _pf(out["t"]), _pf(out["u"]),
                      _pf(out["v"]), None, None)
        ms = c_f(0)
        _gpu_check(gpu.slrgpu_intersect_batch_sbvh(self._s, C.byref(rb), n, C.byref(hb), C.byref(ms)), "slrgpu_intersect_batch_sbvh")
        out["kernel_ms"] = ms.value
        return out

    def occluded(self, rays):
        comps = [_f32(rays[k]) for k in ("ox", "oy", "oz", "dx", "dy", "dz", "tmin", "tmax")]
        n = comps[0].shape[0]
        rb = RayBatch(*[_pf(c) for c in comps])
        occ = np.empty(n, np.uint8)
        ms = c_f(0)
        _gpu_check(gpu.slrgpu_occluded_batch(self._s, C.byref(rb), n, occ.ctypes.data_as(PU8), C.byref(ms)), "slrgpu_occluded_batch")
        return occ, ms.value


# ---- scene files, rendering front end, asset writers (slrhost.h)
host.slrhost_read_scene.restype = C.c_int
host.slrhost_read_scene.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
host.slrhost_scene_context.restype = C.c_int
host.slrhost_scene_context.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
host.slrhost_render.restype = C.c_int
host.slrhost_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, PF, C.POINTER(C.c_double)]
host.slrhost_render_debug.restype = C.c_int
host.slrhost_render_debug.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, PF, C.POINTER(C.c_double)]
host.slrhost_render_range.restype = C.c_int
host.slrhost_render_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, PF, C.POINTER(C.c_double)]
host.slrhost_render_bpt.restype = C.c_int
host.slrhost_render_bpt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, PF, C.POINTER(C.c_double)]
host.slrhost_scene_renderer_method.restype = C.c_int
host.slrhost_scene_renderer_method.argtypes = [C.c_void_p, C.c_char_p, c_u32]
host.slrhost_save_bmp.restype = C.c_int
host.slrhost_save_bmp.argtypes = [C.c_char_p, PF, C.c_int, C.c_int, C.c_int, c_f, c_f]
host.slrhost_accum_to_rgb.restype = C.c_int
host.slrhost_accum_to_rgb.argtypes = [PF, C.c_int, C.c_int, C.c_int, c_f, PF]
host.slrhost_write_assbin.restype = C.c_int
host.slrhost_write_assbin.argtypes = [C.c_char_p, PF, PF, PF, PF, c_u32, PU32, c_u32, C.c_char_p, PF]
host.slrhost_write_exr.restype = C.c_int
host.slrhost_write_exr.argtypes = [C.c_char_p, c_u32, c_u32, PF]
gpu.slrgpu_render.restype = C.c_int
gpu.slrgpu_render.argtypes = [C.c_void_p, C.POINTER(RenderParams), PF, C.POINTER(RenderStats)]
gpu.slrgpu_render_device.restype = C.c_int
gpu.slrgpu_render_device.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.c_void_p, C.c_void_p, C.POINTER(RenderStats)]


def read_scene(path, rgb_mode=False):
    """readScene + Scene::build: parse a scene-description file and flatten it."""
    out = C.c_void_p()
    _host_check(host.slrhost_read_scene(os.fsencode(path), 1 if rgb_mode else 0, C.byref(out)), "slrhost_read_scene")
    hs = HostScene(out.value)
    buf = (C.c_double * 8)()
    host.slrhost_scene_context(hs.handle, buf)
    hs.context = {"width": int(buf[0]), "height": int(buf[1]), "samples": int(buf[2]), "rngSeed": int(buf[3]),
                  "timeStart": buf[4], "timeEnd": buf[5], "brightness": buf[6], "hasRenderer": bool(buf[7])}
    name = C.create_string_buffer(32)
    host.slrhost_scene_renderer_method(hs.handle, name, 32)
    hs.context["method"] = name.value.decode()      # "PT", "BPT", "debug" as the scene file's setRenderer wrote it
    return hs


def host_render(host_scene, width=0, height=0, spp=0, seed=0, device=0, bmp_dir=None, spp_begin=0, out=None, method="PT"):
    """Renderer::render through the host's GPUPathTracingRenderer (method "PT") or GPUBidirectionalPathTracingRenderer
    ("BPT"), global samples [spp_begin, spp_begin + spp).
    Returns (accum[h, w, c], stats); `out` may supply the destination array (e.g. a pinned torch tensor's numpy view)."""
    ctx = getattr(host_scene, "context", {"width": 0, "height": 0})
    w = width or ctx["width"]
    h = height or ctx["height"]
    chan = 3 if host_scene.desc.rgb_mode else 16
    accum = out if out is not None else np.empty((h, w, chan), np.float32)
    assert accum.shape == (h, w, chan) and accum.dtype == np.float32 and accum.flags["C_CONTIGUOUS"]
    st = (C.c_double * 6)()
    t0 = time.perf_counter()
    entry = {"PT": host.slrhost_render_range, "BPT": host.slrhost_render_bpt}[method]
    _host_check(entry(host_scene.handle, device, width, height, spp_begin, spp, seed,
                      os.fsencode(bmp_dir) if bmp_dir else None, _pf(accum), st), "slrhost_render_range")
    call_s = time.perf_counter() - t0
    return accum, {"paths": int(st[0]), "rays": int(st[1]), "device_s": st[2], "wall_s": st[3], "upload_s": st[4], "channels": int(st[5]) % 1000,
                   "devices": int(st[5]) // 1000, "call_s": call_s}


def gpu_render(gpu_scene, width, height, spp_begin, spp_end, seed=1509761209, time_start=0.0, time_end=0.0, flags=0,
               pool_size=0, max_path_length=0):
    """slrgpu_render with host buffers. Returns (accum[h, w, c], RenderStats as dict)."""
    chan = gpu.slrgpu_scene_channels(gpu_scene.handle)
    accum = np.zeros((height, width, chan), np.float32)
    p = RenderParams(C.sizeof(RenderParams), width, height, spp_begin, spp_end, time_start, time_end, seed, max_path_length,
                     pool_size, flags)
    st = RenderStats()
    _gpu_check(gpu.slrgpu_render(gpu_scene.handle, C.byref(p), _pf(accum), C.byref(st)), "slrgpu_render")
    return accum, {k: (list(getattr(st, k)) if k == "class_hits" else getattr(st, k)) for k, _ in RenderStats._fields_}


def gpu_render_multi(gpu_scenes, width, height, spp_begin, spp_end, seed=1509761209, pool_size=0, flags=0):
    """slrgpu_render_multi over scene replicas (normally one per device). Returns (accum[h, w, c], RenderStats as dict)."""
    chan = gpu.slrgpu_scene_channels(gpu_scenes[0].handle)
    accum = np.zeros((height, width, chan), np.float32)
    p = RenderParams(C.sizeof(RenderParams), width, height, spp_begin, spp_end, 0.0, 0.0, seed, 0, pool_size, flags)
    st = RenderStats()
    handles = (C.c_void_p * len(gpu_scenes))(*[g.handle for g in gpu_scenes])
    _gpu_check(gpu.slrgpu_render_multi(handles, len(gpu_scenes), C.byref(p), _pf(accum), C.byref(st)), "slrgpu_render_multi")
    return accum, {k: (list(getattr(st, k)) if k == "class_hits" else getattr(st, k)) for k, _ in RenderStats._fields_}


DEBUG_FLOATS = 10


def host_render_debug(host_scene, width=0, height=0, seed=0, bmp_dir=None, device=0):
    """slrhost_render_debug: the GPU debug (AOV) renderer. Returns out[h, w, 10] (hit, geometric normal, shading
    normal, shading tangent) and writes the three BMPs into bmp_dir if given."""
    ctx = getattr(host_scene, "context", {"width": 0, "height": 0})
    w = width or int(ctx["width"])
    h = height or int(ctx["height"])
    out = np.zeros((h, w, DEBUG_FLOATS), np.float32)
    st = (C.c_double * 6)()
    _host_check(host.slrhost_render_debug(host_scene.handle, device, w, h, seed, os.fsencode(bmp_dir) if bmp_dir else None, _pf(out), st),
                "slrhost_render_debug")
    return out, {"paths": st[0], "rays": st[1], "device_s": st[2], "wall_s": st[3]}


def probe_shading(gpu_scene, probes):
    """slrgpu_probe_shading: probes[n, 14] -> out[n, 64] (layout in include/slrgpu.h)."""
    p = _f32(probes).reshape(-1, 14)
    out = np.zeros((p.shape[0], 64), np.float32)
    _gpu_check(gpu.slrgpu_probe_shading(gpu_scene.handle, _pf(p), p.shape[0], _pf(out)), "slrgpu_probe_shading")
    return out


def probe_shading_bpt(gpu_scene, probes):
    """slrgpu_probe_shading_bpt: the bidirectional path tracer's BSDF queries (reverse values, adjoint on odd probes):
    probes[n, 14] -> out[n, 64] (layout in include/slrgpu.h)."""
    p = _f32(probes).reshape(-1, 14)
    out = np.zeros((p.shape[0], 64), np.float32)
    _gpu_check(gpu.slrgpu_probe_shading_bpt(gpu_scene.handle, _pf(p), p.shape[0], _pf(out)), "slrgpu_probe_shading_bpt")
    return out


host.slrhost_set_option.restype = C.c_int
host.slrhost_set_option.argtypes = [C.c_char_p, C.c_int]


def set_option(name, value):
    """slrhost_set_option: process-wide options of the host library ("export_sbvh")."""
    _host_check(host.slrhost_set_option(name.encode(), int(value)), "slrhost_set_option")


host.slrhost_sample_animated.restype = C.c_int
host.slrhost_sample_animated.argtypes = [PF, PF, c_f, c_f, PF, PF, c_u32, PF, PF, PF]


def sample_animated(mat_begin, mat_end, t_begin, t_end, box, times):
    """slrhost_sample_animated -> (decomposition[46], motion bounds[6], sampled[n, 32])."""
    mb, me = _f32(np.asarray(mat_begin).T.reshape(-1)), _f32(np.asarray(mat_end).T.reshape(-1))        # row-major in, column-major ABI
    bx, tm = _f32(box), _f32(times)
    dec, bounds, out = np.zeros(46, np.float32), np.zeros(6, np.float32), np.zeros((tm.shape[0], 32), np.float32)
    _host_check(host.slrhost_sample_animated(_pf(mb), _pf(me), t_begin, t_end, _pf(bx), _pf(tm), tm.shape[0], _pf(dec), _pf(bounds), _pf(out)),
                "slrhost_sample_animated")
    return dec, bounds, out


host.slrhost_decode_png.restype = C.c_int
host.slrhost_decode_png.argtypes = [C.c_char_p, C.c_int, C.POINTER(c_u32), C.POINTER(c_u32), C.POINTER(c_u32), PU8, c_u64]


def decode_png(path, gamma_correction=False):
    """slrhost_decode_png: (pixels[h, w, c] uint8, has_alpha) the way the reference's loadPNG + libpng deliver them."""
    w, h, ch = c_u32(), c_u32(), c_u32()
    _host_check(host.slrhost_decode_png(os.fsencode(path), int(gamma_correction), C.byref(w), C.byref(h), C.byref(ch), None, 0), "slrhost_decode_png")
    c = ch.value & 0xFF
    out = np.empty((h.value, w.value, c), np.uint8)
    _host_check(host.slrhost_decode_png(os.fsencode(path), int(gamma_correction), C.byref(w), C.byref(h), C.byref(ch),
                                        out.ctypes.data_as(PU8), out.size), "slrhost_decode_png")
    return out, bool(ch.value & 0x100)


host.slrhost_read_exr.restype = C.c_int
host.slrhost_read_exr.argtypes = [C.c_char_p, C.POINTER(c_u32), C.POINTER(c_u32), PF, c_u64]


def read_exr(path):
    """slrhost_read_exr: rgba[h, w, 4] float32 through the host library's EXR reader."""
    w, h = c_u32(), c_u32()
    _host_check(host.slrhost_read_exr(os.fsencode(path), C.byref(w), C.byref(h), None, 0), "slrhost_read_exr")
    out = np.empty((h.value, w.value, 4), np.float32)
    _host_check(host.slrhost_read_exr(os.fsencode(path), C.byref(w), C.byref(h), _pf(out), out.size), "slrhost_read_exr")
    return out


def accum_to_rgb(accum, scale):
    h, w, c = accum.shape
    rgb = np.empty((h, w, 3), np.float32)
    a = np.ascontiguousarray(accum, np.float32)
    _host_check(host.slrhost_accum_to_rgb(_pf(a), w, h, c, scale, _pf(rgb)), "slrhost_accum_to_rgb")
    return rgb


def save_bmp(path, accum, scale, sensitivity=1.0):
    h, w, c = accum.shape
    a = np.ascontiguousarray(accum, np.float32)
    _host_check(host.slrhost_save_bmp(os.fsencode(path), _pf(a), w, h, c, scale, sensitivity), "slrhost_save_bmp")


def write_assbin(path, positions, indices, normals=None, tangents=None, uvs=None, material_name="material", diffuse=None):
    pos = _f32(positions).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
    nrm = _f32(normals).reshape(-1, 3) if normals is not None else None
    tng = _f32(tangents).reshape(-1, 3) if tangents is not None else None
    uv = _f32(uvs).reshape(-1, 2) if uvs is not None else None
    dif = _f32(diffuse) if diffuse is not None else None
    _host_check(host.slrhost_write_assbin(os.fsencode(path), _pf(pos), _pf(nrm), _pf(tng), _pf(uv), pos.shape[0],
                                          idx.ctypes.data_as(PU32), idx.shape[0], material_name.encode(), _pf(dif)),
                "slrhost_write_assbin")


def write_exr(path, rgba):
    a = _f32(rgba)
    h, w, _ = a.shape
    _host_check(host.slrhost_write_exr(os.fsencode(path), w, h, _pf(a)), "slrhost_write_exr")
