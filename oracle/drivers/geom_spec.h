// Binary geometry spec + ray/hit batch files exchanged between the Python tests, the reference
// driver (ref_intersect) and the restatement. All little-endian.
//   spec : "SLRG" u32 version(1) u32 numMeshes { u32 nv u32 nt  f32 pos[3nv]  u32 idx[3nt] }*
//          u32 numPlacements { u32 mesh u32 mode(0 bake, 1 instance) f32 mat[16] column-major }*
//   rays : u64 n, then 8 arrays of n f32: ox oy oz dx dy dz tmin tmax
//   hits : u64 n, u32 prim[n], u32 inst[n], f32 t[n], f32 u[n], f32 v[n] (QBVH::intersect), then the same five arrays from SBVH::intersect
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct GeomMesh { std::vector<float> pos; std::vector<uint32_t> idx; };
struct GeomPlacement { uint32_t mesh, mode; float mat[16]; };
struct GeomSpec { std::vector<GeomMesh> meshes; std::vector<GeomPlacement> placements; };
struct RayFile { uint64_t n = 0; std::vector<float> c[8]; };

static inline void rd(FILE* f, void* p, size_t bytes) {
    if (bytes && fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(1); }
}
static inline GeomSpec readGeomSpec(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(1); }
    char magic[4]; uint32_t ver, nm;
    rd(f, magic, 4); rd(f, &ver, 4); rd(f, &nm, 4);
    GeomSpec s;
    s.meshes.resize(nm);
    for (auto& m : s.meshes) {
        uint32_t nv, nt; rd(f, &nv, 4); rd(f, &nt, 4);
        m.pos.resize(3ull * nv); m.idx.resize(3ull * nt);
        rd(f, m.pos.data(), m.pos.size() * 4); rd(f, m.idx.data(), m.idx.size() * 4);
    }
    uint32_t np; rd(f, &np, 4);
    s.placements.resize(np);
    for (auto& p : s.placements) { rd(f, &p.mesh, 4); rd(f, &p.mode, 4); rd(f, p.mat, 64); }
    fclose(f);
    return s;
}
static inline RayFile readRays(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(1); }
    RayFile r; rd(f, &r.n, 8);
    for (int k = 0; k < 8; ++k) { r.c[k].resize(r.n); rd(f, r.c[k].data(), r.n * 4); }
    fclose(f);
    return r;
}
