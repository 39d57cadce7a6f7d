"""The host library's own PNG decoder (slr_b200/host/assets/png.h: inflate, row filters, bit depths, palettes) against files
written by INDEPENDENT encoders -- Pillow (its own zlib settings / filter heuristics) and OpenCV (libpng) -- and against
Pillow's decode of the same files, plus the transformations the reference asks libpng for in loadPNG
(libSLRSceneGraph/Helper/image_loader.cpp:186-280): RGB gets a 0xFF filler byte, 16-bit samples keep the high byte,
sub-byte samples are unpacked unscaled, palettes expand to RGB, and colour / grey samples go through libpng's 8-bit gamma
table for screen gamma 1.0 against the file gamma 0.45455 (gammaCorrection = false is every caller's default, API.hpp:33),
i.e. floor(255 (v / 255)^2.19998 + .5). Byte work: exact.
"""
import os
import struct
import zlib

import numpy as np
import pytest

from slr_b200 import capi

PIL = pytest.importorskip("PIL.Image")


def libpng_gamma_table(screen=1.0, file_gamma=0.45455):
    r = np.floor(1e15 / np.floor(screen * 1e5 + .5) / np.floor(file_gamma * 1e5 + .5) + .5)
    if 95000 <= r <= 105000:
        return np.arange(256, dtype=np.uint8)
    t = np.floor(255.0 * np.power(np.arange(256) / 255.0, r * 1e-5) + .5).astype(np.uint8)
    t[0], t[255] = 0, 255
    return t


def pattern(h, w, c, seed):
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, c), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    for k in range(c):
        img[..., k] = (xx * (3 + k) + yy * (5 - k) + 17 * k) % 256          # smooth part: exercises every row filter
    noisy = rng.random((h, w)) < 0.3
    img[noisy] = rng.integers(0, 256, (int(noisy.sum()), c), dtype=np.uint8)
    return img


@pytest.mark.parametrize("mode,c", [("RGB", 3), ("RGBA", 4), ("L", 1)])
@pytest.mark.parametrize("size", [(37, 53), (64, 64), (1, 7)])
def test_pillow_written_png_decodes_like_libpng_would(mode, c, size, tmp_path):
    h, w = size
    src = pattern(h, w, c, seed=h * 100 + c)
    path = str(tmp_path / f"p_{mode}_{h}x{w}.png")
    PIL.fromarray(src[..., 0] if c == 1 else src, mode).save(path, optimize=bool(h % 2))
    got, has_alpha = capi.decode_png(path)
    table = libpng_gamma_table()
    back = np.asarray(PIL.open(path))                       # Pillow's own decode = the stored samples
    back = back[..., None] if back.ndim == 2 else back
    assert np.array_equal(back, src)
    if c == 1:
        assert got.shape == (h, w, 1) and not has_alpha
        assert np.array_equal(got[..., 0], table[src[..., 0]])
    else:
        assert got.shape == (h, w, 4) and has_alpha == (c == 4)
        assert np.array_equal(got[..., :3], table[src[..., :3]])
        assert np.array_equal(got[..., 3], src[..., 3] if c == 4 else np.full((h, w), 255, np.uint8))      # filler / alpha untouched
    # with gammaCorrection the product 2.2 x 0.45455 is within 5 % of 1: samples pass through unchanged
    same, _ = capi.decode_png(path, gamma_correction=True)
    assert np.array_equal(same[..., :min(c, 3)], src[..., :min(c, 3)])


def test_opencv_written_png(tmp_path):
    cv2 = pytest.importorskip("cv2")
    src = pattern(45, 31, 3, seed=9)
    path = str(tmp_path / "cv.png")
    assert cv2.imwrite(path, src[..., ::-1], [cv2.IMWRITE_PNG_COMPRESSION, 9])        # OpenCV takes BGR
    got, has_alpha = capi.decode_png(path, gamma_correction=True)
    assert not has_alpha and np.array_equal(got[..., :3], src)


def test_palette_16bit_and_subbyte_depths(tmp_path):
    table = libpng_gamma_table()
    # palette image (Pillow quantises to <= 256 colours): indices expand to RGB + filler
    src = pattern(24, 40, 3, seed=3)
    pal = PIL.fromarray(src, "RGB").quantize(colors=16)
    path = str(tmp_path / "pal.png")
    pal.save(path, bits=4)                                   # 4-bit palette indices
    want = np.asarray(pal.convert("RGB"))
    got, has_alpha = capi.decode_png(path)
    assert got.shape == (24, 40, 4) and not has_alpha
    assert np.array_equal(got[..., :3], table[want]) and (got[..., 3] == 255).all()
    # 16-bit grey: the high byte survives (png_set_strip_16)
    g16 = (np.arange(20 * 30, dtype=np.uint32).reshape(20, 30) * 109 % 65536).astype(np.uint16)
    path = str(tmp_path / "g16.png")
    PIL.fromarray(g16).save(path)
    got, _ = capi.decode_png(path)
    assert np.array_equal(got[..., 0], table[(g16 >> 8).astype(np.uint8)])
    # 1-bit grey: samples are unpacked WITHOUT scaling (png_set_packing), so white is 1, not 255
    bits = (pattern(9, 19, 1, seed=5)[..., 0] > 127)
    path = str(tmp_path / "g1.png")
    PIL.fromarray(bits).save(path, bits=1)
    got, _ = capi.decode_png(path)
    assert np.array_equal(got[..., 0], table[bits.astype(np.uint8)])


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def test_hand_assembled_png_with_each_filter_stored_blocks_and_gama(tmp_path):
    """A PNG assembled chunk by chunk: one row per filter type (computed here), a gAMA chunk of 1.0 (then screen 1.0 x file
    1.0 is not significant and samples pass through), compression level 0 (stored deflate blocks)."""
    h, w = 10, 13
    src = pattern(h, w, 3, seed=11).astype(np.int32)
    rows = bytearray()
    prev = np.zeros((w, 3), np.int32)
    for y in range(h):
        cur = src[y]
        f = y % 5
        left = np.vstack([np.zeros((1, 3), np.int32), cur[:-1]])
        upleft = np.vstack([np.zeros((1, 3), np.int32), prev[:-1]])
        if f == 0: enc = cur
        elif f == 1: enc = cur - left
        elif f == 2: enc = cur - prev
        elif f == 3: enc = cur - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            enc = cur - pred
        rows += bytes([f]) + (enc % 256).astype(np.uint8).tobytes()
        prev = cur
    for level in (0, 6):
        data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + \
            _chunk(b"gAMA", struct.pack(">I", 100000)) + _chunk(b"IDAT", zlib.compress(bytes(rows), level)) + _chunk(b"IEND", b"")
        path = str(tmp_path / f"hand{level}.png")
        with open(path, "wb") as f:
            f.write(data)
        got, _ = capi.decode_png(path)
        assert np.array_equal(got[..., :3], src.astype(np.uint8))
    # a flipped byte in the image data is caught (CRC), not decoded into a wrong texture
    bad = bytearray(data)
    bad[60] ^= 0x40
    with open(str(tmp_path / "bad.png"), "wb") as f:
        f.write(bytes(bad))
    with pytest.raises(capi.SlrError):
        capi.decode_png(str(tmp_path / "bad.png"))


def test_image_texture_from_png_reaches_the_scene(tmp_path):
    """A scene file whose material takes its colour from a PNG (Image2D + SpectrumTexture): the loader converts the texels
    to (u, v, scale) halves like TiledImage2D's constructor (Image.h:178-196)."""
    from slr_b200 import scenes
    path = scenes.SCENES["diffuse"](str(tmp_path), width=16, height=16, spp=1)
    png = str(tmp_path / "tex.png")          # Image2D takes the path as written (relative to the working directory, API.cpp:466)
    PIL.fromarray(pattern(32, 32, 3, seed=2), "RGB").save(png)
    text = open(path).read().replace('diffuseTex = SpectrumTexture(Spectrum(0.75, 0.75, 0.75));',
                                     f'diffuseTex = SpectrumTexture(Image2D("{png}"));', 1)
    with open(path, "w") as f:
        f.write(text)
    with capi.stdout_to_stderr():
        hs = capi.read_scene(path)
    assert hs.desc.num_images == 1
    img = hs.desc.images[0]
    assert (img.width, img.height) == (32, 32) and img.format == 5          # SLRGPU_IMG_UVS16Fx3
