#include "images.h"
#include "exr.h"
#include "png.h"
#include <map>
#include <stdexcept>

namespace slr {

void sRGB_to_uvs(SpectrumType type, const float rgb[3], float uvs[3]) {
    float xyz[3];
    if (type == SpectrumType::Illuminant) {
        xyz[0] = (float)(0.4124564 * rgb[0] + 0.3575761 * rgb[1] + 0.1804375 * rgb[2]);
        xyz[1] = (float)(0.2126729 * rgb[0] + 0.7151522 * rgb[1] + 0.0721750 * rgb[2]);
        xyz[2] = (float)(0.0193339 * rgb[0] + 0.1191920 * rgb[1] + 0.9503041 * rgb[2]);
    } else {
        xyz[0] = (float)(0.4969 * rgb[0] + 0.3391 * rgb[1] + 0.1640 * rgb[2]);
        xyz[1] = (float)(0.2562 * rgb[0] + 0.6782 * rgb[1] + 0.0656 * rgb[2]);
        xyz[2] = (float)(0.0233 * rgb[0] + 0.1130 * rgb[1] + 0.8637 * rgb[2]);
    }
    float b = xyz[0] + xyz[1] + xyz[2];
    float xy[2] = {xyz[0] / b, xyz[1] / b};
    if (b == 0) xy[0] = xy[1] = (float)(1.0f / 3.0);
    uvs[0] = (float)(16.730260708356887 * xy[0] + 7.7801960340706 * xy[1] - 2.170152247475828);
    uvs[1] = (float)(-7.530081094743006 * xy[0] + 16.192422314095225 * xy[1] + 1.1125529268825947);
    uvs[2] = b;
}

void uvs_to_sRGB(SpectrumType type, const float uvs[3], float rgb[3]) {
    float xy[2];
    xy[0] = (float)(0.0491440520940413 * uvs[0] - 0.02361291916573777 * uvs[1] + 0.13292069743203658);
    xy[1] = (float)(0.022853819546830627 * uvs[0] + 0.05077639329371236 * uvs[1] - 0.006895157122499944);
    float b = uvs[2];
    float X = xy[0] * b, Y = xy[1] * b, Z = b - X - Y;
    if (type == SpectrumType::Illuminant) {
        rgb[0] = (float)(3.2404542 * X - 1.5371385 * Y - 0.4985314 * Z);
        rgb[1] = (float)(-0.9692660 * X + 1.8760108 * Y + 0.0415560 * Z);
        rgb[2] = (float)(0.0556434 * X - 0.2040259 * Y + 1.0572252 * Z);
    } else {
        rgb[0] = (float)(2.6897 * X - 1.2759 * Y - 0.4138 * Z);
        rgb[1] = (float)(-1.0221 * X + 1.9783 * Y + 0.0438 * Z);
        rgb[2] = (float)(0.0612 * X - 0.2245 * Y + 1.1633 * Z);
    }
}

// An 8-bit PNG converted to the texel format the renderer samples, case by case as TiledImage2D's constructor does
// (libSLR/Core/Image.h:121-335): RGB_8x4 / RGBA8x4 / Gray8 sources x AsIs / NormalTexture / AlphaTexture store modes.
static Image2DRef loadPNGImage(const std::string& path, ImageStoreMode mode, SpectrumType type, bool rgbMode) {
    png::Image src;
    std::string err;
    if (!png::load(path, false, &src, &err)) throw std::runtime_error(err);        // gammaCorrection = false: API.hpp:33
    auto img = std::make_shared<Image2D>();
    img->width = src.width; img->height = src.height; img->spectrumType = type;
    const size_t n = (size_t)src.width * src.height;
    const uint8_t* p = src.pixels.data();
    if (src.channels == 1) {
        if (mode == ImageStoreMode::NormalTexture) throw std::runtime_error("a grey-scale image cannot be used as a normal map: " + path);
        img->format = SLRGPU_IMG_GRAY8;
        img->data.assign(p, p + n);
        return img;
    }
    if (mode == ImageStoreMode::NormalTexture) {
        img->format = SLRGPU_IMG_RGB8x3;
        img->data.resize(n * 3);
        for (size_t i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) img->data[3 * i + c] = p[4 * i + c];
    } else if (mode == ImageStoreMode::AlphaTexture) {
        img->format = SLRGPU_IMG_GRAY8;
        img->data.resize(n);
        for (size_t i = 0; i < n; ++i) img->data[i] = src.hasAlpha ? p[4 * i + 3] : p[4 * i];      // RGBA: .a, RGB_: .r
    } else if (rgbMode) {
        if (src.hasAlpha) { img->format = SLRGPU_IMG_RGBA8x4; img->data.assign(p, p + n * 4); }
        else {
            img->format = SLRGPU_IMG_RGB8x3;
            img->data.resize(n * 3);
            for (size_t i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) img->data[3 * i + c] = p[4 * i + c];
        }
    } else {
        const int stride = src.hasAlpha ? 4 : 3;
        img->format = src.hasAlpha ? SLRGPU_IMG_UVSA16Fx4 : SLRGPU_IMG_UVS16Fx3;
        img->data.resize(n * stride * 2);
        uint16_t* dst = reinterpret_cast<uint16_t*>(img->data.data());
        for (size_t i = 0; i < n; ++i) {
            const float rgb[3] = {p[4 * i] / 255.0f, p[4 * i + 1] / 255.0f, p[4 * i + 2] / 255.0f};
            float uvs[3];
            sRGB_to_uvs(type, rgb, uvs);
            for (int c = 0; c < 3; ++c) dst[stride * i + c] = exr::floatToHalf(uvs[c]);
            if (src.hasAlpha) dst[4 * i + 3] = exr::floatToHalf(p[4 * i + 3] / 255.0f);
        }
    }
    return img;
}

Image2DRef loadImageCached(const std::string& path, ImageStoreMode mode, SpectrumType type, bool rgbMode) {
    static std::map<std::string, Image2DRef> cache;
    auto it = cache.find(path);
    if (it != cache.end()) return it->second;      // keyed by path only, as the reference does
    const size_t dot = path.find_last_of('.');
    const std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
    if (ext == "png") {
        Image2DRef img = loadPNGImage(path, mode, type, rgbMode);
        cache[path] = img;
        return img;
    }
    if (ext != "exr") throw std::runtime_error("image format ." + ext + " is not supported by this build (.png and uncompressed .exr are): " + path);
    exr::Image src;
    std::string err;
    if (!exr::load(path, &src, &err)) throw std::runtime_error(err);
    auto img = std::make_shared<Image2D>();
    img->width = src.width; img->height = src.height; img->spectrumType = type;
    const size_t n = (size_t)src.width * src.height;
    if (mode == ImageStoreMode::NormalTexture) throw std::runtime_error("a half-float image cannot be used as a normal map: " + path);
    if (mode == ImageStoreMode::AlphaTexture) {
        img->format = SLRGPU_IMG_GRAY8;
        img->data.resize(n);
        for (size_t i = 0; i < n; ++i) {
            float a = exr::halfToFloat(src.rgba[4 * i + 3]);
            img->data[i] = (uint8_t)std::min(uint32_t(255 * a), uint32_t(255));
        }
    } else if (rgbMode) {
        img->format = SLRGPU_IMG_RGBA16Fx4;
        img->data.resize(n * 8);
        std::memcpy(img->data.data(), src.rgba.data(), n * 8);
    } else {
        img->format = SLRGPU_IMG_UVSA16Fx4;
        img->data.resize(n * 8);
        uint16_t* dst = reinterpret_cast<uint16_t*>(img->data.data());
        for (size_t i = 0; i < n; ++i) {
            float rgb[3], uvs[3];
            for (int c = 0; c < 3; ++c) rgb[c] = std::max(exr::halfToFloat(src.rgba[4 * i + c]), 0.0f);
            sRGB_to_uvs(type, rgb, uvs);
            for (int c = 0; c < 3; ++c) dst[4 * i + c] = exr::floatToHalf(uvs[c]);
            dst[4 * i + 3] = src.rgba[4 * i + 3];
        }
    }
    cache[path] = img;
    return img;
}

}  // namespace slr
