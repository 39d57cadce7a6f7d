// Image loading for textures and the environment map: decodes a file (EXR here; PNG needs zlib,
// which this build does not link) and converts it to the texel format the renderer samples, with
// the same per-texel conversion the reference applies when it builds a TiledImage2D
// (libSLR/Core/Image.h:121-351): colour images become (u, v, scale[, alpha]) halves for spectral
// up-sampling, normal maps stay 8-bit RGB, alpha maps become 8-bit grey.
#pragma once
#include "../shading.h"
#include <string>

namespace slr {

enum class ImageStoreMode { AsIs = 0, NormalTexture, AlphaTexture };

// Cached by path like the reference's s_imageDB (API.cpp:1375-1402). Throws std::runtime_error.
Image2DRef loadImageCached(const std::string& path, ImageStoreMode mode, SpectrumType type, bool rgbMode);

// Upsampling::sRGB_to_uvs (Spectrum.h:118-141)
void sRGB_to_uvs(SpectrumType type, const float rgb[3], float uvs[3]);
// Upsampling::uvs_to_sRGB (Spectrum.h:143-166)
void uvs_to_sRGB(SpectrumType type, const float uvs[3], float rgb[3]);

}  // namespace slr
