// Minimal OpenEXR reader/writer: single-part scanline files, NO_COMPRESSION, channels among
// A/B/G/R stored as half or float, increasing-Y line order. That is what the in-repo asset
// generators write and enough for SLR's setEnvironment / Image2D path (the reference reads EXR
// through Imf::RgbaInputFile into half RGBA, libSLRSceneGraph/Helper/image_loader.cpp:38-62).
// Header-only so the oracle's image-loader stand-in decodes files with the same code.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace slr {
namespace exr {

inline uint16_t floatToHalf(float f) { _Float16 h = (_Float16)f; uint16_t b; std::memcpy(&b, &h, 2); return b; }
inline float halfToFloat(uint16_t b) { _Float16 h; std::memcpy(&h, &b, 2); return (float)h; }

struct Image {
    uint32_t width = 0, height = 0;
    std::vector<uint16_t> rgba;      // half r,g,b,a per pixel, row 0 = top scanline
};

inline bool load(const std::string& path, Image* out, std::string* error) {
    auto bad = [&](const std::string& why) { if (error) *error = path + ": " + why; return false; };
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return bad("cannot open");
    std::vector<uint8_t> buf;
    std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    if (got != buf.size() || buf.size() < 16) return bad("short file");
    size_t p = 0;
    auto u32 = [&](size_t at) { uint32_t v; std::memcpy(&v, &buf[at], 4); return v; };
    if (u32(0) != 0x01312f76u) return bad("not an OpenEXR file");
    uint32_t version = u32(4);
    if ((version & 0xFF) != 2 || (version & 0x1E00)) return bad("only single-part scanline EXR is supported");
    p = 8;
    struct Channel { std::string name; int32_t type; };
    std::vector<Channel> channels;
    int32_t dw[4] = {0, 0, -1, -1};
    uint8_t compression = 255, lineOrder = 0;
    while (p < buf.size() && buf[p] != 0) {
        std::string name((const char*)&buf[p]); p += name.size() + 1;
        std::string type((const char*)&buf[p]); p += type.size() + 1;
        if (p + 4 > buf.size()) return bad("corrupt header");
        uint32_t size = u32(p); p += 4;
        if (p + size > buf.size()) return bad("corrupt header");
        if (name == "channels") {
            size_t q = p;
            while (q < p + size && buf[q] != 0) {
                Channel c; c.name = (const char*)&buf[q]; q += c.name.size() + 1;
                std::memcpy(&c.type, &buf[q], 4); q += 16;
                channels.push_back(c);
            }
        } else if (name == "compression") compression = buf[p];
        else if (name == "dataWindow") std::memcpy(dw, &buf[p], 16);
        else if (name == "lineOrder") lineOrder = buf[p];
        p += size;
    }
    ++p;
    if (compression != 0) return bad("compressed EXR is not supported (write it with NO_COMPRESSION)");
    if (lineOrder != 0) return bad("only increasing-Y line order is supported");
    const int32_t W = dw[2] - dw[0] + 1, H = dw[3] - dw[1] + 1;
    if (W <= 0 || H <= 0 || channels.empty()) return bad("bad data window / channel list");
    out->width = W; out->height = H;
    out->rgba.assign((size_t)W * H * 4, 0);
    for (size_t i = 0; i < (size_t)W * H; ++i) out->rgba[4 * i + 3] = floatToHalf(1.0f);
    size_t lineBytes = 0;
    for (const Channel& c : channels) lineBytes += (size_t)W * (c.type == 1 ? 2 : 4);
    if (p + 8ull * H > buf.size()) return bad("truncated offset table");
    for (int32_t y = 0; y < H; ++y) {
        uint64_t off; std::memcpy(&off, &buf[p + 8ull * y], 8);
        if (off + 8 + lineBytes > buf.size()) return bad("truncated scanline");
        int32_t yy = (int32_t)u32(off) - dw[1];
        if (yy < 0 || yy >= H) return bad("scanline outside the data window");
        size_t q = off + 8;
        for (const Channel& c : channels) {
            int slot = c.name == "R" ? 0 : c.name == "G" ? 1 : c.name == "B" ? 2 : c.name == "A" ? 3 : -1;
            for (int32_t x = 0; x < W; ++x) {
                uint16_t h;
                if (c.type == 1) { std::memcpy(&h, &buf[q], 2); q += 2; }
                else if (c.type == 2) { float v; std::memcpy(&v, &buf[q], 4); q += 4; h = floatToHalf(v); }
                else { uint32_t v; std::memcpy(&v, &buf[q], 4); q += 4; h = floatToHalf((float)v); }
                if (slot >= 0) out->rgba[((size_t)yy * W + x) * 4 + slot] = h;
            }
        }
    }
    return true;
}

// Writes half RGBA, uncompressed. `rgba` holds 4 floats per pixel, row 0 = top.
inline bool save(const std::string& path, uint32_t W, uint32_t H, const float* rgba) {
    std::vector<uint8_t> b;
    auto put = [&b](const void* s, size_t n) { b.insert(b.end(), (const uint8_t*)s, (const uint8_t*)s + n); };
    auto str = [&](const char* s) { put(s, std::strlen(s) + 1); };
    auto attr = [&](const char* name, const char* type, const void* data, uint32_t size) { str(name); str(type); put(&size, 4); put(data, size); };
    uint32_t magic = 0x01312f76u, version = 2;
    put(&magic, 4); put(&version, 4);
    std::vector<uint8_t> ch;
    for (const char* n : {"A", "B", "G", "R"}) {
        ch.insert(ch.end(), n, n + 2);
        int32_t type = 1, xs = 1, ys = 1; uint8_t lin[4] = {0, 0, 0, 0};
        ch.insert(ch.end(), (uint8_t*)&type, (uint8_t*)&type + 4); ch.insert(ch.end(), lin, lin + 4);
        ch.insert(ch.end(), (uint8_t*)&xs, (uint8_t*)&xs + 4); ch.insert(ch.end(), (uint8_t*)&ys, (uint8_t*)&ys + 4);
    }
    ch.push_back(0);
    attr("channels", "chlist", ch.data(), (uint32_t)ch.size());
    uint8_t zero = 0;
    attr("compression", "compression", &zero, 1);
    int32_t win[4] = {0, 0, (int32_t)W - 1, (int32_t)H - 1};
    attr("dataWindow", "box2i", win, 16);
    attr("displayWindow", "box2i", win, 16);
    attr("lineOrder", "lineOrder", &zero, 1);
    float one = 1.0f, center[2] = {0.0f, 0.0f};
    attr("pixelAspectRatio", "float", &one, 4);
    attr("screenWindowCenter", "v2f", center, 8);
    attr("screenWindowWidth", "float", &one, 4);
    b.push_back(0);
    const size_t lineBytes = (size_t)W * 2 * 4;
    uint64_t base = b.size() + 8ull * H;
    for (uint32_t y = 0; y < H; ++y) { uint64_t off = base + (uint64_t)y * (8 + lineBytes); put(&off, 8); }
    std::vector<uint16_t> line((size_t)W * 4);
    for (uint32_t y = 0; y < H; ++y) {
        int32_t yy = (int32_t)y, size = (int32_t)lineBytes;
        put(&yy, 4); put(&size, 4);
        const int order[4] = {3, 2, 1, 0};   // A B G R
        for (int c = 0; c < 4; ++c)
            for (uint32_t x = 0; x < W; ++x) line[(size_t)c * W + x] = floatToHalf(rgba[((size_t)y * W + x) * 4 + order[c]]);
        put(line.data(), lineBytes);
    }
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite(b.data(), 1, b.size(), f) == b.size();
    std::fclose(f);
    return ok;
}

}  // namespace exr
}  // namespace slr
