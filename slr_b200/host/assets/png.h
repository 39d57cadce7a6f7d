// PNG decoder for textures (written for this host library: zlib / libpng are not linked): RFC 1950 / 1951 inflate, the
// five PNG row filters, bit depths 1-16, grey / RGB / palette / RGBA, and the transformations the reference asks libpng for
// in loadPNG (libSLRSceneGraph/Helper/image_loader.cpp:186-280):
//   png_set_strip_16            16-bit samples keep their high byte
//   png_set_packing             1/2/4-bit samples are spread to one byte each WITHOUT scaling
//   png_set_palette_to_rgb      palette indices become RGB
//   png_set_filler(0xFF, AFTER) RGB and palette images get a fourth byte 0xFF (ColorFormat::RGB_8x4)
//   png_set_gamma(screen, file) screen = 1.0 (gammaCorrection false, every caller's default, API.hpp:33) or 2.2; file = the
//                               gAMA chunk or 0.45455: colour / grey samples go through libpng's 8-bit table
//                               floor(255 (v / 255)^g + .5) with g = 1e15 / screen / file rounded to 5 decimals, unless
//                               g is within 5 % of 1 -- so by default a texture is LINEARISED (g = 2.19998) when loaded
// Grey+alpha trips an assert in the reference, Adam7 interlacing is not handled there (no png_set_interlace_handling):
// both are errors here. Resulting layouts (getPNGInfo, image_loader.cpp:150-184): grey -> 1 byte (Gray8), RGB / palette ->
// 4 bytes (RGB_8x4), RGBA -> 4 bytes (RGBA8x4).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace slr {
namespace png {

struct Image {
    uint32_t width = 0, height = 0;
    uint32_t channels = 0;            // 1 (grey) or 4 (RGB + 0xFF filler, RGBA)
    bool hasAlpha = false;            // channels == 4 and the fourth byte is the file's alpha
    std::vector<uint8_t> pixels;      // row-major, top row first
};

// ---- inflate (RFC 1951) ------------------------------------------------------------------------------------------
class Inflater {
    const uint8_t* m_in;
    size_t m_size, m_pos = 0;
    uint32_t m_bits = 0;
    int m_count = 0;
    bool m_error = false;

    uint32_t bits(int n) {
        while (m_count < n) {
            if (m_pos >= m_size) { m_error = true; return 0; }
            m_bits |= (uint32_t)m_in[m_pos++] << m_count;
            m_count += 8;
        }
        const uint32_t v = m_bits & ((n == 32) ? 0xFFFFFFFFu : ((1u << n) - 1u));
        m_bits >>= n; m_count -= n;
        return v;
    }
    struct Huffman {
        uint16_t count[16] = {};
        uint16_t symbol[288] = {};
        bool build(const uint8_t* lengths, int n) {
            std::memset(count, 0, sizeof(count));
            for (int i = 0; i < n; ++i) ++count[lengths[i]];
            count[0] = 0;
            int left = 1;
            for (int len = 1; len < 16; ++len) { left <<= 1; left -= count[len]; if (left < 0) return false; }
            uint16_t offs[16];
            offs[1] = 0;
            for (int len = 1; len < 15; ++len) offs[len + 1] = offs[len] + count[len];
            for (int i = 0; i < n; ++i) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
            return true;
        }
    };
    int decode(const Huffman& h) {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; ++len) {
            code |= (int)bits(1);
            if (m_error) return -1;
            const int cnt = h.count[len];
            if (code - cnt < first) return h.symbol[index + (code - first)];
            index += cnt; first += cnt; first <<= 1; code <<= 1;
        }
        return -1;
    }
    bool codes(std::vector<uint8_t>& out, const Huffman& lit, const Huffman& dist) {
        static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint16_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint16_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        for (;;) {
            int sym = decode(lit);
            if (sym < 0) return false;
            if (sym < 256) out.push_back((uint8_t)sym);
            else if (sym == 256) return true;
            else {
                sym -= 257;
                if (sym >= 29) return false;
                const int len = lbase[sym] + (int)bits(lext[sym]);
                const int ds = decode(dist);
                if (ds < 0 || ds >= 30) return false;
                const size_t d = dbase[ds] + bits(dext[ds]);
                if (m_error || d > out.size()) return false;
                const size_t start = out.size() - d;
                for (int k = 0; k < len; ++k) out.push_back(out[start + k]);
            }
        }
    }
public:
    Inflater(const uint8_t* data, size_t size) : m_in(data), m_size(size) {}
    bool run(std::vector<uint8_t>& out) {
        for (;;) {
            const uint32_t last = bits(1), type = bits(2);
            if (m_error) return false;
            if (type == 0) {
                m_bits = 0; m_count = 0;
                if (m_pos + 4 > m_size) return false;
                const uint32_t len = m_in[m_pos] | (m_in[m_pos + 1] << 8), nlen = m_in[m_pos + 2] | (m_in[m_pos + 3] << 8);
                m_pos += 4;
                if ((len ^ 0xFFFFu) != nlen || m_pos + len > m_size) return false;
                out.insert(out.end(), m_in + m_pos, m_in + m_pos + len);
                m_pos += len;
            } else if (type == 1) {
                uint8_t l[288];
                for (int i = 0; i < 144; ++i) l[i] = 8;
                for (int i = 144; i < 256; ++i) l[i] = 9;
                for (int i = 256; i < 280; ++i) l[i] = 7;
                for (int i = 280; i < 288; ++i) l[i] = 8;
                uint8_t d[30];
                for (int i = 0; i < 30; ++i) d[i] = 5;
                Huffman lit, dist;
                lit.build(l, 288); dist.build(d, 30);
                if (!codes(out, lit, dist)) return false;
            } else if (type == 2) {
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                const int nlen = (int)bits(5) + 257, ndist = (int)bits(5) + 1, ncode = (int)bits(4) + 4;
                if (m_error || nlen > 286 || ndist > 30) return false;
                uint8_t lengths[320];
                std::memset(lengths, 0, sizeof(lengths));
                for (int i = 0; i < ncode; ++i) lengths[order[i]] = (uint8_t)bits(3);
                Huffman lencode;
                if (!lencode.build(lengths, 19)) return false;
                int idx = 0;
                uint8_t all[320];
                while (idx < nlen + ndist) {
                    int sym = decode(lencode);
                    if (sym < 0) return false;
                    if (sym < 16) all[idx++] = (uint8_t)sym;
                    else {
                        uint8_t prev = 0;
                        int rep;
                        if (sym == 16) { if (idx == 0) return false; prev = all[idx - 1]; rep = 3 + (int)bits(2); }
                        else if (sym == 17) rep = 3 + (int)bits(3);
                        else rep = 11 + (int)bits(7);
                        if (idx + rep > nlen + ndist) return false;
                        while (rep--) all[idx++] = prev;
                    }
                }
                if (all[256] == 0) return false;
                Huffman lit, dist;
                if (!lit.build(all, nlen)) return false;
                dist.build(all + nlen, ndist);        // an incomplete distance code is legal (one distance only)
                if (!codes(out, lit, dist)) return false;
            } else return false;
            if (last) return !m_error;
        }
    }
};

inline bool zlibDecompress(const std::vector<uint8_t>& in, std::vector<uint8_t>& out, std::string* err) {
    if (in.size() < 6 || (in[0] & 0x0F) != 8 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) { *err = "not a zlib stream"; return false; }
    Inflater inf(in.data() + 2, in.size() - 2);
    if (!inf.run(out)) { *err = "corrupt deflate stream"; return false; }
    uint32_t a = 1, b = 0;
    for (uint8_t v : out) { a = (a + v) % 65521u; b = (b + a) % 65521u; }
    const size_t n = in.size();
    const uint32_t want = ((uint32_t)in[n - 4] << 24) | ((uint32_t)in[n - 3] << 16) | ((uint32_t)in[n - 2] << 8) | in[n - 1];
    if (((b << 16) | a) != want) { *err = "zlib checksum mismatch"; return false; }
    return true;
}

inline uint32_t crc32(const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        ready = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

// libpng's png_reciprocal2 + png_gamma_significant + png_gamma_8bit_correct (png.c): the 8-bit gamma table of png_set_gamma
inline bool gammaTable(double screenGamma, double fileGamma, uint8_t table[256]) {
    const double a = std::floor(screenGamma * 100000.0 + .5), b = std::floor(fileGamma * 100000.0 + .5);
    const double r = std::floor(1e15 / a / b + .5);
    if (r >= 95000.0 && r <= 105000.0) return false;            // not significant: samples are passed through
    for (int i = 0; i < 256; ++i)
        table[i] = (i > 0 && i < 255) ? (uint8_t)std::floor(255.0 * std::pow(i / 255.0, r * .00001) + .5) : (uint8_t)i;
    return true;
}

inline bool load(const std::string& path, bool gammaCorrection, Image* out, std::string* err) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot open " + path; return false; }
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
    std::fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) { *err = path + ": not a PNG file"; return false; }
    auto be32 = [&file](size_t p) { return ((uint32_t)file[p] << 24) | ((uint32_t)file[p + 1] << 16) | ((uint32_t)file[p + 2] << 8) | file[p + 3]; };
    uint32_t width = 0, height = 0, depth = 0, colorType = 0, interlace = 0;
    std::vector<uint8_t> idat, palette;
    double fileGamma = 0.45455;
    uint8_t sbit[4] = {0, 0, 0, 0};
    bool haveSbit = false, haveHeader = false;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t len = be32(pos);
        if (pos + 12 + (size_t)len > file.size()) { *err = path + ": truncated PNG chunk"; return false; }
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        const uint8_t* data = &file[pos + 8];
        if (crc32(&file[pos + 4], 4 + len) != be32(pos + 8 + len)) { *err = path + ": PNG chunk CRC mismatch"; return false; }
        if (!std::memcmp(type, "IHDR", 4) && len == 13) {
            width = be32(pos + 8); height = be32(pos + 12); depth = data[8]; colorType = data[9]; interlace = data[12];
            haveHeader = true;
        } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!std::memcmp(type, "gAMA", 4) && len == 4) fileGamma = be32(pos + 8) / 100000.0;
        else if (!std::memcmp(type, "sBIT", 4) && len <= 4) { std::memcpy(sbit, data, len); haveSbit = true; }
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!haveHeader || width == 0 || height == 0) { *err = path + ": PNG without a valid IHDR"; return false; }
    if (interlace) { *err = path + ": interlaced PNG (the reference does not de-interlace either)"; return false; }
    if (colorType == 4) { *err = path + ": grey+alpha PNG is not supported (it trips an assert in the reference, image_loader.cpp:196)"; return false; }
    const uint32_t samples = colorType == 0 ? 1 : colorType == 2 ? 3 : colorType == 3 ? 1 : colorType == 6 ? 4 : 0;
    if (!samples || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16) || (colorType == 3 && depth == 16) ||
        ((colorType == 2 || colorType == 6) && depth < 8)) { *err = path + ": unsupported PNG colour type / bit depth"; return false; }
    if (colorType == 3 && palette.size() < 3) { *err = path + ": palette PNG without PLTE"; return false; }
    std::vector<uint8_t> raw;
    if (!zlibDecompress(idat, raw, err)) { *err = path + ": " + *err; return false; }
    const size_t bpp = std::max<size_t>(1, (size_t)samples * depth / 8);          // filter distance in bytes
    const size_t rowBytes = ((size_t)width * samples * depth + 7) / 8;
    if (raw.size() < (rowBytes + 1) * (size_t)height) { *err = path + ": PNG image data too short"; return false; }
    // ---- unfilter in place
    std::vector<uint8_t> prev(rowBytes, 0);
    for (uint32_t y = 0; y < height; ++y) {
        uint8_t* row = &raw[(rowBytes + 1) * (size_t)y];
        const uint8_t filter = row[0];
        uint8_t* cur = row + 1;
        for (size_t i = 0; i < rowBytes; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = cur[i];
            switch (filter) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                      v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
            default: *err = path + ": unknown PNG row filter"; return false;
            }
            cur[i] = (uint8_t)v;
        }
        std::memcpy(prev.data(), cur, rowBytes);
    }
    // ---- samples -> bytes (strip_16 / packing / shift), palette -> RGB, filler, gamma
    out->width = width; out->height = height;
    out->channels = (colorType == 0) ? 1 : 4;
    out->hasAlpha = colorType == 6;
    out->pixels.assign((size_t)width * height * out->channels, 0xFF);
    uint8_t gtab[256];
    const bool useGamma = gammaTable(gammaCorrection ? 2.2 : 1.0, fileGamma, gtab);
    auto sample = [&](const uint8_t* row, size_t index) -> uint8_t {      // index-th sample of a row, as one byte
        if (depth == 8) return row[index];
        if (depth == 16) return row[2 * index];                            // png_set_strip_16: the high byte
        const size_t bit = index * depth;
        return (uint8_t)((row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u));   // png_set_packing: no scaling
    };
    for (uint32_t y = 0; y < height; ++y) {
        const uint8_t* row = &raw[(rowBytes + 1) * (size_t)y + 1];
        uint8_t* dst = &out->pixels[(size_t)y * width * out->channels];
        for (uint32_t x = 0; x < width; ++x) {
            if (colorType == 0) {
                uint8_t v = sample(row, x);
                // png_set_shift: samples whose significant bits are fewer than the depth are shifted down to them
                if (haveSbit && sbit[0] && sbit[0] < std::min<uint32_t>(depth, 8)) v >>= (std::min<uint32_t>(depth, 8) - sbit[0]);
                dst[x] = useGamma ? gtab[v] : v;
            } else if (colorType == 3) {
                const uint32_t idx = sample(row, x);
                for (int c = 0; c < 3; ++c) {
                    const uint8_t v = (idx * 3 + c < palette.size()) ? palette[idx * 3 + c] : 0;
                    dst[4 * x + c] = useGamma ? gtab[v] : v;
                }
            } else {
                for (uint32_t c = 0; c < samples; ++c) {
                    uint8_t v = sample(row, (size_t)x * samples + c);
                    if (haveSbit && c < 4 && sbit[c] && sbit[c] < 8 && depth >= 8) v >>= (8 - sbit[c]);
                    dst[4 * x + c] = (c < 3 && useGamma) ? gtab[v] : v;      // alpha is never gamma corrected
                }
            }
        }
    }
    return true;
}

}  // namespace png
}  // namespace slr
