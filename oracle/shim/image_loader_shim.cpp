// Replacement for libSLRSceneGraph/Helper/image_loader.cpp (which needs libpng + OpenEXR): the same
// two entry points (image_loader.h:27-28) over the in-repo uncompressed-EXR reader. PNG/JPEG are not
// available. Test infrastructure for the oracle build only.
#include <libSLRSceneGraph/Helper/image_loader.h>
#include "../../slr_b200/host/assets/exr.h"

static bool isExr(const std::string& p) { size_t d = p.find_last_of('.'); return d != std::string::npos && p.substr(d + 1) == "exr"; }

bool getImageInfo(const std::string& filePath, uint32_t* width, uint32_t* height, uint64_t* requiredSize, ColorFormat* color) {
    if (!isExr(filePath)) return false;
    slr::exr::Image img; std::string err;
    if (!slr::exr::load(filePath, &img, &err)) { fprintf(stderr, "%s\n", err.c_str()); return false; }
    *width = img.width; *height = img.height; *requiredSize = (uint64_t)img.width * img.height * 8; *color = ColorFormat::RGBA16Fx4;
    return true;
}

bool loadImage(const std::string& filePath, uint8_t* storage, bool) {
    if (!isExr(filePath)) return false;
    slr::exr::Image img; std::string err;
    if (!slr::exr::load(filePath, &img, &err)) return false;
    memcpy(storage, img.rgba.data(), (size_t)img.width * img.height * 8);
    return true;
}
