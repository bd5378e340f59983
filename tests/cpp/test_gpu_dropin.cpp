// GPU check of the drop-in C++ headers beyond the reference application (run by tests/test_gpu_cpp.py):
//   1. drt::render with RenderOptions::devices = {0 .. n-1} (drtb_multi_render: image bands over the GPUs of the box,
//      image assembled on the first GPU by the kernels' peer stores, gradients summed) against the same render on
//      one device: the image must be bit-equal, the gradients equal up to the order of addition;
//   2. a Scene with Triangle<T> shapes (gpu::flatten -> drtb_mesh -> GPU BVH) rendered on one device and on all of
//      them; the image is dumped for the Python side, which renders the same scene through the C ABI.
// usage: test_gpu_dropin <n_gpus> <out.bin>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include "drt/integrate.hpp"
#include "drt/pathtracer.hpp"
#include "drt/render.hpp"

using namespace drt;
using T = double;
using V = Vector<T, 3>;

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static double max_rel(const std::vector<V>& a, const std::vector<V>& b)
{
    double m = 0;
    for (std::size_t i = 0; i < a.size(); ++i)
        for (int c = 0; c < 3; ++c) m = std::fmax(m, std::fabs(a[i][c] - b[i][c]) / std::fmax(std::fabs(b[i][c]), 1e-12));
    return m;
}

int main(int argc, char** argv)
{
    const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 1;
    const char* out_path = argc > 2 ? argv[2] : nullptr;
    std::vector<int> all;
    for (int g = 0; g < n_gpus; ++g) all.push_back(g);

    // ---- 1. the Cornell box of src/render.cpp:26-59 on one device and on all of them
    Vector<T, 3, true> red(V{0.5, 0, 0}, true), green(V{0, 0.5, 0}, true), white(V{0.5, 0.5, 0.5}, true), emission(V(1), true);
    auto d_red = std::make_shared<DiffuseBxDF<T>>(red);
    auto d_green = std::make_shared<DiffuseBxDF<T>>(green);
    auto d_white = std::make_shared<DiffuseBxDF<T>>(white);
    auto emitter = std::make_shared<AreaEmitter<T>>(emission);
    Sphere<T> sphere_front(V{0., 0., 3.}, 1., d_white), sphere_back(V{-1., 1., 4.5}, 1., d_white);
    Plane<T> left_plane(V{-1., 0., 0.}, -3., d_red), right_plane(V{1., 0., 0.1}, -3., d_green), back_plane(V{0., 0., -1.}, -6., d_white);
    Plane<T> front_plane(V{0, 0, 1}, 0, d_white), ground_plane(V{0., 1., 0.}, -3., d_white), ceiling_plane(V{0., -1., 0.}, -3., d_white);
    Sphere<T> light(V{0., 3., 3.}, 1., nullptr, emitter);
    Scene<T> box{&sphere_front, &sphere_back, &left_plane, &right_plane, &back_plane, &front_plane, &ground_plane, &ceiling_plane, &light};
    const std::size_t W = 96, H = 72;                   // 9 bands of 8 rows: ragged over 2, 4 and 8 devices
    Camera<T> cam(W, H);
    cam.look_at(V{0, 0, 0}, V{0, 0, 1});
    for (int setting = 0; setting < 2; ++setting) {
        Pathtracer<T> tracer(setting ? 0.5 : 1.0, setting ? 1 : 8);
        const std::size_t spp = setting ? 40 : 8;
        std::vector<V> one(W * H), many(W * H);
        drtb_stats s1{}, sn{};
        const V r0 = red.grad(), w0 = white.grad(), e0 = emission.grad();    // gradients are ADDED to the handles (vector.hpp:185-188)
        RenderOptions o1; o1.stats = &s1;
        render(box, cam, tracer, spp, one.data(), o1);
        const V g_red = red.grad() - r0, g_white = white.grad() - w0, g_emit = emission.grad() - e0;
        const V r1 = red.grad(), w1 = white.grad(), e1 = emission.grad();
        RenderOptions on; on.devices = all; on.stats = &sn;
        render(box, cam, tracer, spp, many.data(), on);
        CHECK(max_rel(many, one) == 0.0);
        CHECK(sn.paths == s1.paths && sn.segments == s1.segments && sn.lit_paths == s1.lit_paths);
        for (int c = 0; c < 3; ++c) {                   // the same sums up to the order of addition over the devices
            CHECK(std::fabs((red.grad()[c] - r1[c]) - g_red[c]) <= 1e-10 * std::fabs(g_red[c]) + 1e-300);
            CHECK(std::fabs((white.grad()[c] - w1[c]) - g_white[c]) <= 1e-10 * std::fabs(g_white[c]));
            CHECK(std::fabs((emission.grad()[c] - e1[c]) - g_emit[c]) <= 1e-10 * std::fabs(g_emit[c]));
        }
        std::printf("box setting %d: %d device(s) match one device; red.grad = %.9f %.9f %.9f\n", setting, n_gpus,
                    g_red[0], g_red[1], g_red[2]);
    }

    // ---- 2. triangles: an open room of two-triangle walls around the analytic light
    Vector<T, 3, true> floor_col(V{0.7, 0.6, 0.5}, true), wall_col(V{0.3, 0.5, 0.8}, true);
    auto d_floor = std::make_shared<DiffuseBxDF<T>>(floor_col);
    auto d_wall = std::make_shared<DiffuseBxDF<T>>(wall_col);
    std::vector<std::unique_ptr<Triangle<T>>> tris;
    auto quad = [&](V a, V b, V c, V d, std::shared_ptr<BxDF<T>> m) {       // wound so that the normal cross(v1 - v0, v2 - v0) faces the room
        tris.emplace_back(new Triangle<T>(a, c, b, m));
        tris.emplace_back(new Triangle<T>(a, d, c, m));
    };
    quad(V{-3, -3, 0}, V{3, -3, 0}, V{3, -3, 6}, V{-3, -3, 6}, d_floor);           // floor y = -3
    quad(V{-3, -3, 6}, V{3, -3, 6}, V{3, 3, 6}, V{-3, 3, 6}, d_wall);              // back z = 6
    quad(V{-3, -3, 0}, V{-3, -3, 6}, V{-3, 3, 6}, V{-3, 3, 0}, d_wall);            // x = -3
    quad(V{3, -3, 0}, V{3, 3, 0}, V{3, 3, 6}, V{3, -3, 6}, d_floor);               // x = +3
    Scene<T> room{&sphere_front, &light};
    for (auto& t : tris) room.push_back(t.get());
    Camera<T> cam2(64, 48);
    cam2.look_at(V{0, 0, 0}, V{0, 0, 1});
    Pathtracer<T> tracer2(1.0, 4);
    std::vector<V> t_one(64 * 48), t_many(64 * 48);
    RenderOptions p1;
    render(room, cam2, tracer2, 16, t_one.data(), p1);
    const V g_floor = floor_col.grad(), g_wall = wall_col.grad();
    RenderOptions pn; pn.devices = all;
    render(room, cam2, tracer2, 16, t_many.data(), pn);
    CHECK(max_rel(t_many, t_one) == 0.0);
    double mean = 0;
    for (const V& v : t_one) mean += v[0] + v[1] + v[2];
    CHECK(mean > 0 && g_floor[0] > 0 && g_wall[2] > 0);
    for (int c = 0; c < 3; ++c) CHECK(std::fabs(floor_col.grad()[c] - 2 * g_floor[c]) <= 1e-10 * std::fabs(g_floor[c]));
    std::printf("triangle room: mean %.6f floor.grad = %.9f %.9f %.9f wall.grad = %.9f %.9f %.9f\n", mean / (64 * 48 * 3),
                g_floor[0], g_floor[1], g_floor[2], g_wall[0], g_wall[1], g_wall[2]);
    // a single ray through the tape-compatible entry point (Pathtracer::trace, batch of one)
    auto L = tracer2.trace(room, V{0, 0, 0}, normalize(V{0.1, -0.6, 1}));
    CHECK(std::isfinite(double(L.detach()[0])));
    if (out_path) {
        std::FILE* f = std::fopen(out_path, "wb");
        if (!f) return 2;
        std::fwrite(t_one.data(), sizeof(V), t_one.size(), f);
        const double g[6] = {g_floor[0], g_floor[1], g_floor[2], g_wall[0], g_wall[1], g_wall[2]};
        std::fwrite(g, sizeof(double), 6, f);
        std::fclose(f);
    }
    std::printf(failures ? "%d FAILURES\n" : "all GPU drop-in checks passed\n", failures);
    return failures ? 1 : 0;
}
