// Counter-based per-path random numbers: Philox4x32-10 keyed by the render seed, counter =
// (pixel, global sample index, block). Replaces the reference's per-thread xorshift128 streams
// (libSLR/Core/light_path_samplers.h:44-61, RNGs/XORShiftRNG.cpp:29-36): a path's numbers depend
// only on (pixel, sample, dimension), so any partition of the samples over GPUs draws the same set.
// The uint32 -> [0,1) mapping is the reference's (RandomNumberGenerator.cpp:12-15).
#pragma once
#include <stdint.h>

namespace slrgpu {

struct Rand4 { float x, y, z, w; };

__device__ __forceinline__ float bitsToUnitFloat(uint32_t bits) {
    return __uint_as_float((bits >> 9) | 0x3f800000u) - 1.0f;
}

__device__ __forceinline__ Rand4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Rand4 r;
    r.x = bitsToUnitFloat(c0); r.y = bitsToUnitFloat(c1); r.z = bitsToUnitFloat(c2); r.w = bitsToUnitFloat(c3);
    return r;
}

// Random block `block` of path (pixel, sample). Blocks 0,1: camera sample (time, pixel x, pixel y,
// wavelength offset | wavelength selection, lens u0, lens u1, -). Bounce k >= 1: block 2k = (light
// selection, light u0, light u1, BSDF component), block 2k+1 = (BSDF u0, BSDF u1, Russian roulette
// of the NEXT hit, -).
__device__ __forceinline__ Rand4 pathRandom(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t block) {
    return philox4x32(pixel, sample, block, 0x5352u, seed, 0x534C5242u);
}

}  // namespace slrgpu
