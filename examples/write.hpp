// examples/write.hpp — drt::write_exr without OpenEXR.
//
// Stands in for the reference's src/write.hpp:10-26, which converts the image
// to Imf::Rgba (half precision, alpha = 1) and writes it through
// Imf::RgbaOutputFile.  OpenEXR is an empty submodule here, so this writes the
// same pixels as a plain OpenEXR 2 scanline file by hand: channels A, B, G, R
// of type HALF, increasing-y line order, row 0 = top, one scanline per chunk,
// NO_COMPRESSION (the library default would be PIZ; every EXR reader accepts
// both).  Same name, same signature, so `write_exr(path, img, w, h)` at
// src/render.cpp:90 compiles unchanged.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "drt/vector.hpp"

namespace drt {

namespace exr_detail {

// float -> IEEE binary16, round to nearest even; overflow -> inf, NaN stays NaN
// (what half(float) does in the reference's writer).
inline std::uint16_t to_half(float f)
{
    std::uint32_t x;
    std::memcpy(&x, &f, 4);
    const std::uint32_t sign = (x >> 16) & 0x8000u;
    const std::uint32_t mag = x & 0x7fffffffu;
    if (mag >= 0x7f800000u)                                   // inf / NaN
        return std::uint16_t(sign | 0x7c00u | (mag > 0x7f800000u ? 0x0200u | ((mag >> 13) & 0x3ffu) : 0u));
    if (mag >= 0x477ff000u) return std::uint16_t(sign | 0x7c00u);   // rounds to >= 65520: inf
    if (mag < 0x33000001u) return std::uint16_t(sign);              // <= 2^-25: rounds to zero
    std::uint32_t e = mag >> 23, m = mag & 0x7fffffu;
    std::uint32_t h;
    if (e < 113) {                                            // half subnormal: value = m' * 2^-24
        m |= 0x800000u;
        const std::uint32_t shift = 126 - e;                  // 14 .. 24
        h = m >> shift;
        const std::uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
    } else {
        h = ((e - 112) << 10) | (m >> 13);
        const std::uint32_t rem = m & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;   // carry into the exponent is correct
    }
    return std::uint16_t(sign | h);
}

struct Buffer {
    std::vector<unsigned char> b;
    void bytes(const void* p, std::size_t n) { const auto* c = static_cast<const unsigned char*>(p); b.insert(b.end(), c, c + n); }
    void str(const char* s) { bytes(s, std::strlen(s) + 1); }
    void u8(std::uint8_t v) { b.push_back(v); }
    void i32(std::int32_t v) { unsigned char c[4]; for (int i = 0; i < 4; ++i) c[i] = (std::uint32_t(v) >> (8 * i)) & 0xff; bytes(c, 4); }
    void u64(std::uint64_t v) { unsigned char c[8]; for (int i = 0; i < 8; ++i) c[i] = (v >> (8 * i)) & 0xff; bytes(c, 8); }
    void f32(float v) { std::int32_t i; std::memcpy(&i, &v, 4); i32(i); }
    void attr(const char* name, const char* type, std::int32_t size) { str(name); str(type); i32(size); }
};

} // namespace exr_detail

template <typename T>
inline void write_exr(const char* fname, const Vector<T, 3>* data, std::size_t width, std::size_t height)
{
    using namespace exr_detail;
    if (width == 0 || height == 0 || width > 0x3fffffffu || height > 0x3fffffffu)
        throw std::runtime_error("write_exr: bad image size");
    Buffer h;
    h.i32(20000630);                                          // magic 76 2f 31 01
    h.i32(2);                                                 // version 2, single-part scanline, no flags
    static const char* const names[4] = {"A", "B", "G", "R"}; // chlist is sorted by name
    h.attr("channels", "chlist", 4 * 18 + 1);
    for (const char* n : names) {
        h.str(n); h.i32(1 /* HALF */); h.u8(0 /* pLinear */); h.u8(0); h.u8(0); h.u8(0); h.i32(1); h.i32(1);
    }
    h.u8(0);
    h.attr("compression", "compression", 1); h.u8(0);         // NO_COMPRESSION
    for (const char* n : {"dataWindow", "displayWindow"}) {
        h.attr(n, "box2i", 16);
        h.i32(0); h.i32(0); h.i32(std::int32_t(width) - 1); h.i32(std::int32_t(height) - 1);
    }
    h.attr("lineOrder", "lineOrder", 1); h.u8(0);             // INCREASING_Y: row 0 is the top row
    h.attr("pixelAspectRatio", "float", 4); h.f32(1.0f);
    h.attr("screenWindowCenter", "v2f", 8); h.f32(0.0f); h.f32(0.0f);
    h.attr("screenWindowWidth", "float", 4); h.f32(1.0f);
    h.u8(0);                                                  // end of header

    const std::uint64_t row_bytes = std::uint64_t(width) * 4 * 2;
    const std::uint64_t chunk = 8 + row_bytes;                // y, size, pixel data
    const std::uint64_t first = h.b.size() + 8 * std::uint64_t(height);
    for (std::size_t y = 0; y < height; ++y) h.u64(first + chunk * y);

    std::FILE* f = std::fopen(fname, "wb");
    if (!f) throw std::runtime_error(std::string("write_exr: cannot open ") + fname);
    bool ok = std::fwrite(h.b.data(), 1, h.b.size(), f) == h.b.size();
    std::vector<unsigned char> line(chunk);
    for (std::size_t y = 0; y < height && ok; ++y) {
        Buffer c;
        c.b.reserve(chunk);
        c.i32(std::int32_t(y)); c.i32(std::int32_t(row_bytes));
        for (int ch = 0; ch < 4; ++ch) {                      // A, B, G, R planes of this scanline
            for (std::size_t x = 0; x < width; ++x) {
                const float v = ch == 0 ? 1.0f : float(double(data[y * width + x][3 - ch]));
                const std::uint16_t hv = to_half(v);
                c.u8(hv & 0xff); c.u8(hv >> 8);
            }
        }
        ok = std::fwrite(c.b.data(), 1, c.b.size(), f) == c.b.size();
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error(std::string("write_exr: write failed: ") + fname);
}

} // namespace drt
