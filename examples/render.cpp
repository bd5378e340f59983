// examples/render.cpp — the reference application (src/render.cpp) on the GPU path.
//
// Same scene, same six command-line flags as the reference's TCLAP parser
// (src/args.hpp:20-67: -x/--width 640, -y/--height 480, -n/--samples 100,
// -b/--min-bounces 1, -p/--absorb-prob 0.5, -o/--output required), with the
// gradient call of src/render.cpp:79-80 enabled.  TCLAP and OpenEXR are not
// needed: flags are parsed by hand and the image is written by the
// dependency-free write_exr of examples/write.hpp (half RGBA scanlines, what
// src/write.hpp:10-26 produces through OpenEXR); an output name ending in
// ".pfm" selects a float32 PFM instead (no half-precision rounding, for diffs).
//
//   g++ -std=c++17 -O2 -Iinclude examples/render.cpp -o build/render
//       -Ldifferentiable-renderer_b200/lib -ldrtb   (plus an rpath to that lib directory)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "drt/bxdf.hpp"
#include "drt/camera.hpp"
#include "drt/emitter.hpp"
#include "drt/integrate.hpp"
#include "drt/pathtracer.hpp"
#include "drt/render.hpp"
#include "drt/shape.hpp"
#include "drt/vector.hpp"
#include "write.hpp"

using namespace drt;

struct Args {
    std::size_t width = 640, height = 480, samples = 100, min_bounces = 1;
    double absorb_prob = 0.5;
    std::string output;
    bool f32 = false;
    int gpus = 1;              // --gpus N (NEW): spread the image bands over N GPUs of one box
};

static bool parse_args(int argc, const char* argv[], Args* a)
{
    auto is = [](const char* s, const char* sh, const char* lg) { return !std::strcmp(s, sh) || !std::strcmp(s, lg); };
    for (int i = 1; i < argc; ++i) {
        const char* f = argv[i];
        if (!std::strcmp(f, "--f32")) { a->f32 = true; continue; }
        if (!std::strcmp(f, "--gpus") && i + 1 < argc) { a->gpus = int(std::strtol(argv[++i], nullptr, 10)); continue; }
        if (i + 1 >= argc) { std::fprintf(stderr, "PARSE ERROR: missing value for %s\n", f); return false; }
        const char* v = argv[++i];
        if (is(f, "-x", "--width")) a->width = std::strtoull(v, nullptr, 10);
        else if (is(f, "-y", "--height")) a->height = std::strtoull(v, nullptr, 10);
        else if (is(f, "-n", "--samples")) a->samples = std::strtoull(v, nullptr, 10);
        else if (is(f, "-b", "--min-bounces")) a->min_bounces = std::strtoull(v, nullptr, 10);
        else if (is(f, "-p", "--absorb-prob")) a->absorb_prob = std::strtod(v, nullptr);
        else if (is(f, "-o", "--output")) a->output = v;
        else { std::fprintf(stderr, "PARSE ERROR: unknown flag %s\n", f); return false; }
    }
    if (a->output.empty()) { std::fprintf(stderr, "PARSE ERROR: Required argument missing: output\n"); return false; }
    return true;
}

static bool write_pfm(const char* path, const Vector<double, 3>* img, std::size_t w, std::size_t h)
{
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    std::fprintf(f, "PF\n%zu %zu\n-1.0\n", w, h);             // little-endian, rows bottom-to-top
    std::vector<float> row(3 * w);
    for (std::size_t y = h; y-- > 0;) {
        for (std::size_t x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) row[3 * x + c] = float(img[y * w + x][c]);
        std::fwrite(row.data(), sizeof(float), row.size(), f);
    }
    return std::fclose(f) == 0;
}

int main(int argc, const char* argv[])
{
    Args args;
    if (!parse_args(argc, argv, &args)) return EXIT_FAILURE;

    using T = double;

    // Scene parameters, materials, shapes and order: src/render.cpp:26-59.
    Vector<T, 3, true> red(Vector<T, 3>{0.5, 0, 0}, true);
    Vector<T, 3, true> green(Vector<T, 3>{0, 0.5, 0}, true);
    Vector<T, 3, true> white(Vector<T, 3>{0.5, 0.5, 0.5}, true);
    Vector<T, 3, true> emission(Vector<T, 3>(1), true);

    auto diffuse_red = std::make_shared<DiffuseBxDF<T>>(red);
    auto diffuse_green = std::make_shared<DiffuseBxDF<T>>(green);
    auto diffuse_white = std::make_shared<DiffuseBxDF<T>>(white);
    auto specular_white = std::make_shared<SpecularBxDF<T>>(white, 30);   // unused, as in the reference
    auto emitter = std::make_shared<AreaEmitter<T>>(emission);

    Sphere<T> sphere_front(Vector<T, 3>{0., 0., 3.}, 1., diffuse_white);
    Sphere<T> sphere_back(Vector<T, 3>{-1., 1., 4.5}, 1., diffuse_white);
    Plane<T> left_plane(Vector<T, 3>{-1., 0., 0.}, -3., diffuse_red);
    Plane<T> right_plane(Vector<T, 3>{1., 0., 0.1}, -3., diffuse_green);
    Plane<T> back_plane(Vector<T, 3>{0., 0., -1.}, -6., diffuse_white);
    Plane<T> front_plane(Vector<T, 3>{0, 0, 1}, 0, diffuse_white);
    Plane<T> ground_plane(Vector<T, 3>{0., 1., 0.}, -3., diffuse_white);
    Plane<T> ceiling_plane(Vector<T, 3>{0., -1., 0.}, -3., diffuse_white);
    Sphere<T> light(Vector<T, 3>{0., 3., 3.}, 1., nullptr, emitter);

    Scene<T> scene;
    for (Shape<T>* s : std::initializer_list<Shape<T>*>{&sphere_front, &sphere_back, &left_plane, &right_plane,
                                                         &back_plane, &front_plane, &ground_plane, &ceiling_plane,
                                                         &light})
        scene.push_back(s);

    Camera<T> cam(args.width, args.height);
    cam.look_at(Vector<T, 3>{0, 0, 0}, Vector<T, 3>{0, 0, 1});
    gpu::PinnedImage img(args.width * args.height);       // pinned: the kernel writes the pixels into it, no copy after the render

    Pathtracer<T> tracer(args.absorb_prob, args.min_bounces);

    // The whole pixel loop of src/render.cpp:72-86, gradients included.
    drtb_stats st{};
    RenderOptions opt;
    opt.precision = args.f32 ? DRTB_F32 : DRTB_F64;
    for (int g = 0; g < args.gpus; ++g) opt.devices.push_back(g);     // --gpus N: image bands over N GPUs of the box
    opt.stats = &st;
    try {
        render(scene, cam, tracer, args.samples, img.data(), opt);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "render failed: %s\n", e.what());
        return EXIT_FAILURE;
    }
    std::printf("100.00%%\n");
    std::printf("paths %llu  segments/path %.3f  lit %.4f  kernel %.3f ms  %.1f Mpaths/s\n",
                (unsigned long long)st.paths, double(st.segments) / double(st.paths),
                double(st.lit_paths) / double(st.paths), st.kernel_ms, double(st.paths) / st.kernel_ms / 1e3);
    const char* names[4] = {"red", "green", "white", "emission"};
    const Vector<T, 3, true>* ps[4] = {&red, &green, &white, &emission};
    for (int k = 0; k < 4; ++k)
        std::printf("%s.grad = %.9f %.9f %.9f\n", names[k], ps[k]->grad()[0], ps[k]->grad()[1], ps[k]->grad()[2]);

    // Write radiance to file: src/render.cpp:90
    const std::string& out = args.output;
    const bool pfm = out.size() > 4 && out.compare(out.size() - 4, 4, ".pfm") == 0;
    try {
        if (pfm) {
            if (!write_pfm(out.c_str(), img.data(), args.width, args.height)) throw std::runtime_error("cannot write " + out);
        } else {
            write_exr(out.c_str(), img.data(), args.width, args.height);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return EXIT_FAILURE;
    }
    return 0;
}
