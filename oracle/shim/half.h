// Stand-in for OpenEXR's <half.h>, which the reference includes for texel storage only
// (libSLR/Core/Image.h:15). IEEE binary16, round-to-nearest-even via the compiler's _Float16.
// Test infrastructure: lets the untouched reference compile on this image.
#pragma once
struct half {
    _Float16 v;
    half() {}
    half(float f) : v((_Float16)f) {}
    operator float() const { return (float)v; }
};
