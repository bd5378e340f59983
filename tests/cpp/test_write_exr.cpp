// Writes a small known image through examples/write.hpp (the stand-in for the
// reference's src/write.hpp:10-26); tests/test_cpp_headers.py decodes the file.
#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

#include "drt/vector.hpp"
#include "write.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    const std::size_t w = 7, h = 5;
    std::vector<drt::Vector<double, 3>> img(w * h);
    for (std::size_t y = 0; y < h; ++y)
        for (std::size_t x = 0; x < w; ++x)
            img[y * w + x] = drt::Vector<double, 3>{0.125 * double(x) + 1e-3 * double(y), std::ldexp(1.0, int(x) - 20 - int(y)),
                                                    double(y) * 1000.0 + 1.0 / 3.0};
    img[0] = drt::Vector<double, 3>{0.0, 1e6, -2.5};                       // zero, overflow -> inf, negative
    img[1] = drt::Vector<double, 3>{6.1e-5, 5.9604644775390625e-8, 2.9802322387695312e-8};   // subnormal half, min subnormal, tie -> 0
    img[2] = drt::Vector<double, 3>{65504.0, 65519.9, 65520.0};           // max half, rounds down, rounds to inf
    img[3] = drt::Vector<double, 3>{std::numeric_limits<double>::quiet_NaN(), 1.00048828125, 1.00146484375};   // NaN, two exact ties
    drt::write_exr(argv[1], img.data(), w, h);
    bool threw = false;
    try { drt::write_exr("/nonexistent-dir/x.exr", img.data(), w, h); } catch (const std::runtime_error&) { threw = true; }
    std::puts(threw ? "exr written" : "no throw");
    return threw ? 0 : 1;
}
