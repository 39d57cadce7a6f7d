// Reader/writer for the subset of Assimp's binary dump format (.assbin) that SLR's model import uses.
// Header-only and dependency-free: the product's load3DModel uses the reader; the asset generators
// use the writer; the oracle's assimp stand-in (oracle/shim/assimp) includes the same reader so the
// reference and the product decode model files identically. Asset decoding is plumbing, not part of
// the rendering hot path.
//
// Layout (Assimp's AssbinExporter/AssbinLoader, uncompressed, non-shortened):
//   512-byte header: 44-byte "ASSIMP.binary-dump." signature + date, u32 versionMajor, versionMinor,
//     versionRevision, compileFlags, u16 shortened, u16 compressed, 256-byte source name,
//     128-byte command line, padding to 512 bytes
//   chunk = u32 magic, u32 payload size, payload
//   AISCENE 0x1239: u32 flags, numMeshes, numMaterials, numAnimations, numTextures, numLights,
//     numCameras; root AINODE chunk; AIMESH chunks; AIMATERIAL chunks
//   AINODE  0x123c: aiString name (u32 len + bytes), 16 f32 row-major transform, u32 numChildren,
//     u32 numMeshes, u32 mesh indices, child AINODE chunks
//   AIMESH  0x1237: u32 primitiveTypes, numVertices, numFaces, numBones, materialIndex, components
//     bit 0x1 positions, 0x2 normals, 0x4 tangents+bitangents, 0x100<<n texcoord set n, 0x10000<<n colour set n;
//     arrays of 3 f32 per vertex; per texcoord set u32 numUVComponents + 3 f32 per vertex; faces:
//     u16 numIndices then u16 (numVertices < 65536) or u32 indices
//   AIMATERIAL 0x123d: u32 numProperties, AIMATERIALPROPERTY chunks 0x123e: aiString key,
//     u32 semantic, index, dataLength, type, data bytes
// Only triangle meshes, one UV set, names, transforms and the material keys "?mat.name",
// "$clr.diffuse/specular/emissive" and "$tex.file" are interpreted; everything else is skipped.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace slr {
namespace assbin {

enum : uint32_t { kChunkScene = 0x1239, kChunkNode = 0x123c, kChunkMesh = 0x1237, kChunkMaterial = 0x123d, kChunkMaterialProperty = 0x123e };
enum : uint32_t { kHasPositions = 0x1, kHasNormals = 0x2, kHasTangents = 0x4, kHasTexCoordBase = 0x100, kHasColorBase = 0x10000 };
enum : uint32_t { kPrimitiveTriangle = 0x4 };
// aiTextureType values used as the "semantic" of $tex.file properties
enum : uint32_t { kTexDiffuse = 1, kTexSpecular = 2, kTexEmissive = 4, kTexHeight = 5, kTexNormals = 6, kTexOpacity = 8, kTexDisplacement = 9 };

struct MaterialProperty {
    std::string key;
    uint32_t semantic = 0, index = 0, type = 0;   // type: 1 float, 3 string, 5 buffer (aiPropertyTypeInfo)
    std::vector<uint8_t> data;
};
struct Material {
    std::vector<MaterialProperty> properties;
    bool getString(const std::string& key, uint32_t semantic, uint32_t index, std::string* out) const {
        for (const MaterialProperty& p : properties)
            if (p.key == key && p.semantic == semantic && p.index == index && p.type == 3 && p.data.size() >= 4) {
                uint32_t len; std::memcpy(&len, p.data.data(), 4);
                if (p.data.size() < 4 + (size_t)len) return false;
                out->assign((const char*)p.data.data() + 4, len);
                return true;
            }
        return false;
    }
    bool getColor(const std::string& key, float rgb[3]) const {
        for (const MaterialProperty& p : properties)
            if (p.key == key && p.type == 1 && p.data.size() >= 12) { std::memcpy(rgb, p.data.data(), 12); return true; }
        return false;
    }
    uint32_t textureCount(uint32_t semantic) const {
        uint32_t n = 0;
        for (const MaterialProperty& p : properties) if (p.key == "$tex.file" && p.semantic == semantic) n = std::max(n, p.index + 1);
        return n;
    }
    void setString(const std::string& key, const std::string& value, uint32_t semantic = 0, uint32_t index = 0) {
        MaterialProperty p; p.key = key; p.semantic = semantic; p.index = index; p.type = 3;
        uint32_t len = (uint32_t)value.size();
        p.data.resize(4 + len + 1, 0);
        std::memcpy(p.data.data(), &len, 4); std::memcpy(p.data.data() + 4, value.data(), len);
        properties.push_back(p);
    }
    void setColor(const std::string& key, float r, float g, float b) {
        MaterialProperty p; p.key = key; p.type = 1; p.data.resize(12);
        float c[3] = {r, g, b}; std::memcpy(p.data.data(), c, 12);
        properties.push_back(p);
    }
};
struct Mesh {
    std::string name;
    uint32_t primitiveTypes = kPrimitiveTriangle, materialIndex = 0, numUVComponents = 0;
    std::vector<float> positions, normals, tangents, bitangents, texCoords;   // 3 floats per vertex each
    std::vector<uint32_t> indices;                                            // 3 per face
    uint32_t numVertices() const { return (uint32_t)positions.size() / 3; }
};
struct Node {
    std::string name;
    float transform[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // row-major a1..d4
    std::vector<uint32_t> meshes;
    std::vector<Node> children;
};
struct Scene {
    std::vector<Mesh> meshes;
    std::vector<Material> materials;
    Node root;
};

// ---- reader
struct Reader {
    const uint8_t* p; const uint8_t* end; bool ok;
    Reader(const uint8_t* b, const uint8_t* e) : p(b), end(e), ok(true) {}
    template <typename T> T get() { T v{}; if (p + sizeof(T) > end) { ok = false; return v; } std::memcpy(&v, p, sizeof(T)); p += sizeof(T); return v; }
    std::string str() { uint32_t n = get<uint32_t>(); if (!ok || p + n > end) { ok = false; return ""; } std::string s((const char*)p, n); p += n; return s; }
    void floats(std::vector<float>& v, size_t n) { if (p + 4 * n > end) { ok = false; return; } v.resize(n); std::memcpy(v.data(), p, 4 * n); p += 4 * n; }
};

inline bool readNode(Reader& r, Node* n) {
    if (r.get<uint32_t>() != kChunkNode) return false;
    uint32_t size = r.get<uint32_t>();
    const uint8_t* chunkEnd = r.p + size;
    n->name = r.str();
    for (int i = 0; i < 16; ++i) n->transform[i] = r.get<float>();
    uint32_t nc = r.get<uint32_t>(), nm = r.get<uint32_t>();
    if (!r.ok || nm > 1u << 24 || nc > 1u << 24) return false;
    for (uint32_t i = 0; i < nm; ++i) n->meshes.push_back(r.get<uint32_t>());
    n->children.resize(nc);
    for (uint32_t i = 0; i < nc; ++i) if (!readNode(r, &n->children[i])) return false;
    if (r.p > chunkEnd) return false;
    r.p = chunkEnd;
    return r.ok;
}

inline bool readMesh(Reader& r, Mesh* m) {
    if (r.get<uint32_t>() != kChunkMesh) return false;
    uint32_t size = r.get<uint32_t>();
    const uint8_t* chunkEnd = r.p + size;
    m->primitiveTypes = r.get<uint32_t>();
    uint32_t nv = r.get<uint32_t>(), nf = r.get<uint32_t>();
    r.get<uint32_t>();                       // bones
    m->materialIndex = r.get<uint32_t>();
    uint32_t comps = r.get<uint32_t>();
    if (!r.ok) return false;
    if (comps & kHasPositions) r.floats(m->positions, 3ull * nv);
    if (comps & kHasNormals) r.floats(m->normals, 3ull * nv);
    if (comps & kHasTangents) { r.floats(m->tangents, 3ull * nv); r.floats(m->bitangents, 3ull * nv); }
    for (uint32_t c = 0; c < 8; ++c) if (comps & (kHasColorBase << c)) { std::vector<float> skip; r.floats(skip, 4ull * nv); }
    for (uint32_t t = 0; t < 8; ++t) {
        if (!(comps & (kHasTexCoordBase << t))) continue;
        uint32_t nuv = r.get<uint32_t>();
        std::vector<float> tc; r.floats(tc, 3ull * nv);
        if (t == 0) { m->numUVComponents = nuv; m->texCoords.swap(tc); }
    }
    m->indices.reserve(3ull * nf);
    for (uint32_t f = 0; f < nf && r.ok; ++f) {
        uint16_t ni = r.get<uint16_t>();
        uint32_t idx[3] = {0, 0, 0};
        for (uint16_t k = 0; k < ni; ++k) {
            uint32_t v = nv < 65536 ? (uint32_t)r.get<uint16_t>() : r.get<uint32_t>();
            if (k < 3) idx[k] = v;
        }
        if (ni == 3) { m->indices.push_back(idx[0]); m->indices.push_back(idx[1]); m->indices.push_back(idx[2]); }
        else m->primitiveTypes |= 0x8;      // polygon: the importer ignores non-triangle meshes
    }
    if (r.p > chunkEnd) return false;
    r.p = chunkEnd;
    return r.ok;
}

inline bool readMaterial(Reader& r, Material* m) {
    if (r.get<uint32_t>() != kChunkMaterial) return false;
    uint32_t size = r.get<uint32_t>();
    const uint8_t* chunkEnd = r.p + size;
    uint32_t np = r.get<uint32_t>();
    for (uint32_t i = 0; i < np && r.ok; ++i) {
        if (r.get<uint32_t>() != kChunkMaterialProperty) return false;
        uint32_t psize = r.get<uint32_t>();
        const uint8_t* pend = r.p + psize;
        MaterialProperty p;
        p.key = r.str();
        p.semantic = r.get<uint32_t>(); p.index = r.get<uint32_t>();
        uint32_t len = r.get<uint32_t>();
        p.type = r.get<uint32_t>();
        if (!r.ok || r.p + len > r.end) return false;
        p.data.assign(r.p, r.p + len);
        r.p = pend;
        m->properties.push_back(p);
    }
    r.p = chunkEnd;
    return r.ok;
}

inline bool load(const std::string& path, Scene* out, std::string* error) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { if (error) *error = "cannot open " + path; return false; }
    std::vector<uint8_t> buf;
    std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    auto bad = [&](const char* why) { if (error) *error = path + ": " + why; return false; };
    if (got != buf.size() || buf.size() < 512 + 8) return bad("file too short for an assbin header");
    if (std::memcmp(buf.data(), "ASSIMP.binary-dump.", 19) != 0) return bad("not an assbin file");
    uint16_t shortened, compressed;
    std::memcpy(&shortened, buf.data() + 44 + 16, 2);
    std::memcpy(&compressed, buf.data() + 44 + 18, 2);
    if (shortened || compressed) return bad("shortened/compressed assbin files are not supported");
    Reader r(buf.data() + 512, buf.data() + buf.size());
    if (r.get<uint32_t>() != kChunkScene) return bad("missing scene chunk");
    r.get<uint32_t>();
    r.get<uint32_t>();                                   // flags
    uint32_t nMeshes = r.get<uint32_t>(), nMaterials = r.get<uint32_t>();
    for (int i = 0; i < 4; ++i) r.get<uint32_t>();       // animations, textures, lights, cameras
    if (!r.ok || nMeshes > 1u << 20 || nMaterials > 1u << 20) return bad("corrupt scene chunk");
    if (!readNode(r, &out->root)) return bad("corrupt node chunk");
    out->meshes.resize(nMeshes);
    for (uint32_t i = 0; i < nMeshes; ++i) if (!readMesh(r, &out->meshes[i])) return bad("corrupt mesh chunk");
    out->materials.resize(nMaterials);
    for (uint32_t i = 0; i < nMaterials; ++i) if (!readMaterial(r, &out->materials[i])) return bad("corrupt material chunk");
    return true;
}

// ---- writer
struct Writer {
    std::vector<uint8_t> b;
    template <typename T> void put(T v) { const uint8_t* s = (const uint8_t*)&v; b.insert(b.end(), s, s + sizeof(T)); }
    void str(const std::string& s) { put<uint32_t>((uint32_t)s.size()); b.insert(b.end(), s.begin(), s.end()); }
    void floats(const std::vector<float>& v) { const uint8_t* s = (const uint8_t*)v.data(); b.insert(b.end(), s, s + 4 * v.size()); }
    void chunk(uint32_t magic, const Writer& payload) { put<uint32_t>(magic); put<uint32_t>((uint32_t)payload.b.size()); b.insert(b.end(), payload.b.begin(), payload.b.end()); }
};

inline void writeNode(Writer& w, const Node& n) {
    Writer p;
    p.str(n.name);
    for (int i = 0; i < 16; ++i) p.put<float>(n.transform[i]);
    p.put<uint32_t>((uint32_t)n.children.size());
    p.put<uint32_t>((uint32_t)n.meshes.size());
    for (uint32_t m : n.meshes) p.put<uint32_t>(m);
    for (const Node& c : n.children) writeNode(p, c);
    w.chunk(kChunkNode, p);
}

inline bool save(const std::string& path, const Scene& s) {
    Writer w;
    char header[512];
    std::memset(header, 0, sizeof(header));
    std::snprintf(header, 45, "ASSIMP.binary-dump.%-25s", "slr_b200 synthetic asset");
    uint32_t ver[4] = {3, 1, 0, 0};
    std::memcpy(header + 44, ver, 16);
    std::memset(header + 44 + 20 + 256 + 128, 0xcd, 512 - (44 + 20 + 256 + 128));
    w.b.insert(w.b.end(), header, header + 512);
    Writer sc;
    sc.put<uint32_t>(0);
    sc.put<uint32_t>((uint32_t)s.meshes.size());
    sc.put<uint32_t>((uint32_t)s.materials.size());
    for (int i = 0; i < 4; ++i) sc.put<uint32_t>(0);
    writeNode(sc, s.root);
    for (const Mesh& m : s.meshes) {
        Writer p;
        const uint32_t nv = m.numVertices();
        p.put<uint32_t>(m.primitiveTypes); p.put<uint32_t>(nv); p.put<uint32_t>((uint32_t)m.indices.size() / 3);
        p.put<uint32_t>(0); p.put<uint32_t>(m.materialIndex);
        uint32_t comps = kHasPositions;
        if (m.normals.size() == m.positions.size()) comps |= kHasNormals;
        if (m.tangents.size() == m.positions.size() && m.bitangents.size() == m.positions.size()) comps |= kHasTangents;
        if (m.texCoords.size() == m.positions.size()) comps |= kHasTexCoordBase;
        p.put<uint32_t>(comps);
        p.floats(m.positions);
        if (comps & kHasNormals) p.floats(m.normals);
        if (comps & kHasTangents) { p.floats(m.tangents); p.floats(m.bitangents); }
        if (comps & kHasTexCoordBase) { p.put<uint32_t>(m.numUVComponents ? m.numUVComponents : 2); p.floats(m.texCoords); }
        for (size_t f = 0; f + 2 < m.indices.size(); f += 3) {
            p.put<uint16_t>(3);
            for (int k = 0; k < 3; ++k) { if (nv < 65536) p.put<uint16_t>((uint16_t)m.indices[f + k]); else p.put<uint32_t>(m.indices[f + k]); }
        }
        sc.chunk(kChunkMesh, p);
    }
    for (const Material& m : s.materials) {
        Writer p;
        p.put<uint32_t>((uint32_t)m.properties.size());
        for (const MaterialProperty& pr : m.properties) {
            Writer q;
            q.str(pr.key);
            q.put<uint32_t>(pr.semantic); q.put<uint32_t>(pr.index);
            q.put<uint32_t>((uint32_t)pr.data.size()); q.put<uint32_t>(pr.type);
            q.b.insert(q.b.end(), pr.data.begin(), pr.data.end());
            p.chunk(kChunkMaterialProperty, q);
        }
        sc.chunk(kChunkMaterial, p);
    }
    w.chunk(kChunkScene, sc);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite(w.b.data(), 1, w.b.size(), f) == w.b.size();
    std::fclose(f);
    return ok;
}

}  // namespace assbin
}  // namespace slr
