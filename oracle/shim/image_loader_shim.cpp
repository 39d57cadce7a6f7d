// Replacement for libSLRSceneGraph/Helper/image_loader.cpp (which needs libpng + OpenEXR, neither is in this image): the
// same two entry points (image_loader.h:27-28) over the in-repo decoders -- uncompressed EXR (slr_b200/host/assets/exr.h)
// and PNG (assets/png.h, which applies the libpng transformations loadPNG asks for, image_loader.cpp:186-280, and is
// pinned on its own against Pillow- / OpenCV-written files in tests/test_png_reader.py). What the reference does with the
// decoded texels (TiledImage2D conversion, texture lookup) is the reference's own code. JPEG is not available.
// Test infrastructure for the oracle build only.
#include <libSLRSceneGraph/Helper/image_loader.h>
#include "../../slr_b200/host/assets/exr.h"
#include "../../slr_b200/host/assets/png.h"

static std::string extension(const std::string& p) { size_t d = p.find_last_of('.'); return d == std::string::npos ? "" : p.substr(d + 1); }

bool getImageInfo(const std::string& filePath, uint32_t* width, uint32_t* height, uint64_t* requiredSize, ColorFormat* color) {
    const std::string ext = extension(filePath);
    std::string err;
    if (ext == "exr") {
        slr::exr::Image img;
        if (!slr::exr::load(filePath, &img, &err)) { fprintf(stderr, "%s\n", err.c_str()); return false; }
        *width = img.width; *height = img.height; *requiredSize = (uint64_t)img.width * img.height * 8; *color = ColorFormat::RGBA16Fx4;
        return true;
    }
    if (ext == "png") {
        slr::png::Image img;
        if (!slr::png::load(filePath, false, &img, &err)) { fprintf(stderr, "%s\n", err.c_str()); return false; }
        *width = img.width; *height = img.height; *requiredSize = (uint64_t)img.width * img.height * img.channels;
        *color = img.channels == 1 ? ColorFormat::Gray8 : img.hasAlpha ? ColorFormat::RGBA8x4 : ColorFormat::RGB_8x4;      // getPNGInfo
        return true;
    }
    return false;
}

bool loadImage(const std::string& filePath, uint8_t* storage, bool gammaCorrection) {
    const std::string ext = extension(filePath);
    std::string err;
    if (ext == "exr") {
        slr::exr::Image img;
        if (!slr::exr::load(filePath, &img, &err)) return false;
        memcpy(storage, img.rgba.data(), (size_t)img.width * img.height * 8);
        return true;
    }
    if (ext == "png") {
        slr::png::Image img;
        if (!slr::png::load(filePath, gammaCorrection, &img, &err)) return false;
        memcpy(storage, img.pixels.data(), img.pixels.size());
        return true;
    }
    return false;
}
