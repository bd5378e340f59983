// oracle/ref_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference headers (/root/reference/include/drt/*.hpp,
// located through -I at build time, never copied into this repo) through a
// restatement of the pixel loop of src/render.cpp:72-86 with the commented
// `radiance.backward(seed)` of src/render.cpp:79-80 enabled.  src/render.cpp
// itself cannot be built here: src/args.hpp needs TCLAP and src/write.hpp needs
// OpenEXR, both empty submodules; neither is on the hot path.
//
// Built into oracle/_ref/libdrt_ref.so by oracle/Makefile.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it.
//
// Two preprocessor tricks, both leave the reference sources untouched:
//   -Ddrt=drt_ref         (Makefile) renames the namespace so it can never
//                         collide with this repo's own include/drt headers;
//   #define rand ...      (below, AFTER <cstdlib> is in) makes the only RNG call
//                         site, include/drt/random.hpp:9, draw from the
//                         counter-based per-(pixel,sample) stream of drtb.h.
// Every standard header the reference includes is pulled in HERE, before the
// `rand` macro exists, so the macro can only ever touch random.hpp:9.
#include <cstdlib>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <array>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <tuple>
#include <type_traits>
#include <typeinfo>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/drtb.h"       // by path: this repo's include/ is NOT on the search path (see Makefile)

namespace {

inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

thread_local uint64_t g_key = 0;
thread_local uint32_t g_ctr = 0;
thread_local int      g_libc = 0;      // 1: as-shipped sequential glibc rand()
thread_local uint64_t g_draws = 0;

int libc_rand() { return ::rand(); }   // bound before the macro below

int oracle_rand()
{
    ++g_draws;
    if (g_libc) return libc_rand();
    uint32_t slot = g_ctr++;
    return int(splitmix64(g_key * 0x100000001B3ull + slot) % 2147483647ull);
}

} // namespace

#define rand oracle_rand
#include "drt/bxdf.hpp"
#include "drt/camera.hpp"
#include "drt/emitter.hpp"
#include "drt/integrate.hpp"     // must precede pathtracer.hpp (it forgets it)
#include "drt/pathtracer.hpp"
#include "drt/shape.hpp"
#include "drt/vector.hpp"
#undef rand

// Guard: these must be the REFERENCE's headers.  Its tape is a class tree with a
// VariableNode (reference vector.hpp:165-192); this repo's drop-in headers have
// no such type, so picking them up by mistake fails to compile right here.
static_assert(sizeof(drt::internal::VariableNode<double, 3>) > 0, "not the reference's include/drt");

using namespace drt;
using T = double;                                    // src/render.cpp:22

namespace {

// Test-only extension: the reference has no triangle shape.
// This subclass plugs triangles into the reference's OWN raycast / scatter /
// tape machinery, with the semantics fixed in include/drtb.h, so that the
// mesh extension of oracle/restate.c can be checked against reference code
// rather than against itself.
template <typename S>
class Triangle : public Shape<S> {
public:
    Triangle(Vector<S, 3> v0, Vector<S, 3> v1, Vector<S, 3> v2, std::shared_ptr<BxDF<S>> bxdf = nullptr,
             std::shared_ptr<Emitter<S>> emitter = nullptr)
      : Shape<S>(bxdf, emitter), m_v0(v0), m_e1(v1 - v0), m_e2(v2 - v0), m_n(normalize(cross(v1 - v0, v2 - v0))) {}

    bool intersect(Vector<S, 3> orig, Vector<S, 3> dir, double& t) const override
    {
        Vector<S, 3> p = cross(dir, m_e2);
        double det = dot(m_e1, p);
        if (det == 0.0) return false;
        double inv = 1.0 / det;
        Vector<S, 3> tv = orig - m_v0;
        double u = dot(tv, p) * inv;
        if (!(u >= 0.0 && u <= 1.0)) return false;
        Vector<S, 3> q = cross(tv, m_e1);
        double v = dot(dir, q) * inv;
        if (!(v >= 0.0 && u + v <= 1.0)) return false;
        t = dot(m_e2, q) * inv;
        return t > 0;
    }
    Vector<S, 3> normal(Vector<S, 3>) const override { return m_n; }

private:
    Vector<S, 3> m_v0, m_e1, m_e2, m_n;
};

// One private copy of the src/render.cpp:26-65 object graph (the reference is
// not thread-safe: VariableNode::m_grad is a plain +=, vector.hpp:187).
struct World {
    std::vector<Vector<T, 3, true>> params;
    std::vector<std::shared_ptr<BxDF<T>>> materials;
    std::vector<std::unique_ptr<Shape<T>>> shapes;
    Scene<T> scene;

    explicit World(const drtb_scene& s, const drtb_mesh* mesh = nullptr)
    {
        for (int k = 0; k < s.n_params; ++k) {
            Vector<T, 3> v{s.params[3*k], s.params[3*k+1], s.params[3*k+2]};
            params.emplace_back(v, true);
            // VariableNode::m_grad is never initialised (vector.hpp:168,191)
            params.back().grad() = Vector<T, 3>(0);
        }
        for (int m = 0; m < s.n_materials; ++m) {
            const drtb_material& mm = s.materials[m];
            if (mm.color < 0 || mm.color >= s.n_params)
                throw std::runtime_error("bad material");
            if (mm.type == DRTB_DIFFUSE)
                materials.push_back(std::make_shared<DiffuseBxDF<T>>(params[mm.color]));
            else if (mm.type == DRTB_SPECULAR)           // the reference's own SpecularBxDF, bxdf.hpp:85-124
                materials.push_back(std::make_shared<SpecularBxDF<T>>(params[mm.color], mm.exponent));
            else
                throw std::runtime_error("bad material");
        }
        for (int i = 0; i < s.n_prims; ++i) {
            const drtb_prim& p = s.prims[i];
            std::shared_ptr<BxDF<T>> bx =
                p.material >= 0 ? materials.at(p.material) : nullptr;
            std::shared_ptr<Emitter<T>> em =
                p.emission >= 0
                    ? std::make_shared<AreaEmitter<T>>(params.at(p.emission))
                    : nullptr;
            Vector<T, 3> a{p.v[0], p.v[1], p.v[2]};
            if (p.type == DRTB_SPHERE)
                shapes.emplace_back(new Sphere<T>(a, p.v[3], bx, em));
            else if (p.type == DRTB_PLANE)
                shapes.emplace_back(new Plane<T>(a, p.v[3], bx, em));
            else
                throw std::runtime_error("bad prim type");
            scene.push_back(shapes.back().get());
        }
        if (mesh) {                                   // triangles follow the analytic primitives
            std::vector<std::shared_ptr<BxDF<T>>> by_param(size_t(s.n_params));
            for (int64_t i = 0; i < mesh->n_triangles; ++i) {
                auto vtx = [&](int c) {
                    const double* q = mesh->vertices + 3 * int64_t(mesh->indices[3 * i + c]);
                    return Vector<T, 3>{q[0], q[1], q[2]};
                };
                const int col = mesh->color ? mesh->color[i] : -1, em = mesh->emission ? mesh->emission[i] : -1;
                std::shared_ptr<BxDF<T>> bx;
                if (col >= 0) {
                    if (!by_param.at(col)) by_param[col] = std::make_shared<DiffuseBxDF<T>>(params.at(col));
                    bx = by_param[col];
                }
                std::shared_ptr<Emitter<T>> e = em >= 0 ? std::make_shared<AreaEmitter<T>>(params.at(em)) : nullptr;
                shapes.emplace_back(new Triangle<T>(vtx(0), vtx(1), vtx(2), bx, e));
                scene.push_back(shapes.back().get());
            }
        }
    }
};

Camera<T> make_camera(const drtb_camera& c)
{
    auto v3 = [](const double* p) { return Vector<T, 3>{p[0], p[1], p[2]}; };
    return Camera<T>(c.width, c.height, c.vfov, v3(c.eye), v3(c.forward),
                     v3(c.right), v3(c.up));
}

inline bool row_in_shard(int y, const drtb_render_opts& o)
{
    if (o.shard_count <= 1) return true;
    int band = o.band_rows > 0 ? o.band_rows : 1;
    return (y / band) % o.shard_count == o.shard_index;
}

} // namespace

extern "C" {

// Rows are written compactly in increasing y, exactly as drtb_render does.
// rand_mode 0: counter stream (drtb.h); 1: as-shipped sequential glibc rand()
// (single thread, loop order of src/render.cpp:72-76).  Returns 0 or -1.
// gimg != nullptr: additionally the per-pixel gradient image of parameter gparam
// (the README.md:138-145 figure): every pixel's samples back-propagate into
// zeroed grad() accumulators, which are read and added to the running totals.
int drt_ref_render_gimg(const drtb_scene* s, const drtb_mesh* mesh, const drtb_render_opts* o,
                        const double* seed_img, double* img, double* grad, int gparam, double* gimg,
                        int n_threads, int rand_mode, uint64_t* draws_out)
{
    if (mesh && mesh->n_triangles == 0) mesh = nullptr;
    if (gimg && (gparam < 0 || gparam >= s->n_params)) return -1;
    try {
        const int W = s->camera.width, H = s->camera.height, spp = o->spp;
        std::vector<int> rows;
        for (int y = 0; y < H; ++y)
            if (row_in_shard(y, *o)) rows.push_back(y);
        if (rand_mode == 1) n_threads = 1;
        if (n_threads < 1) n_threads = 1;
        const int P = s->n_params;
        std::vector<double> gsum(size_t(n_threads) * P * 3, 0.0);
        uint64_t draws = 0;
        const double scale = o->seed_scale;
        const bool want_grad = (o->flags & DRTB_FLAG_GRAD) != 0;
        const uint64_t key0 = o->seed * 0x9E3779B97F4A7C15ull;

#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads) reduction(+ : draws)
#endif
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            World w(*s, mesh);
            Camera<T> cam = make_camera(s->camera);
            Pathtracer<T> tracer(o->absorb, size_t(o->min_bounces));
            g_libc = rand_mode;
            g_draws = 0;
            std::vector<double> gtot(size_t(P) * 3, 0.0);   // gimg mode: totals over finished pixels
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
            for (size_t r = 0; r < rows.size(); ++r) {
                const size_t y = size_t(rows[r]);
                for (size_t x = 0; x < size_t(W); ++x) {
                    Vector<T, 3> seed(scale);
                    if (seed_img)
                        for (int c = 0; c < 3; ++c)
                            seed[c] = scale * seed_img[(r * W + x) * 3 + c];
                    Vector<T, 3> pixel_radiance(0);
                    for (size_t i = 0; i < size_t(spp); ++i) {
                        g_key = key0 + (uint64_t(y) * W + x) * spp + i;
                        g_ctr = 0;
                        auto [dir, pdf] = cam.sample(x, y);       // camera.hpp:51
                        Vector<T, 3, true> radiance =
                            tracer.trace(w.scene, cam.eye(), dir); // pathtracer.hpp:121
                        pixel_radiance += radiance.detach() / pdf; // render.cpp:78
                        if (want_grad) radiance.backward(seed);    // render.cpp:80
                    }
                    Vector<T, 3> px = pixel_radiance / double(spp); // render.cpp:82
                    if (img)
                        for (int c = 0; c < 3; ++c)
                            img[(r * W + x) * 3 + c] = px[c];
                    if (gimg) {
                        for (int c = 0; c < 3; ++c)
                            gimg[(r * W + x) * 3 + c] = w.params[gparam].grad()[c];
                        for (int k = 0; k < P; ++k) {
                            for (int c = 0; c < 3; ++c) gtot[size_t(k) * 3 + c] += w.params[k].grad()[c];
                            w.params[k].grad() = Vector<T, 3>(0);
                        }
                    }
                }
            }
            for (int k = 0; k < P; ++k)
                for (int c = 0; c < 3; ++c)
                    gsum[(size_t(tid) * P + k) * 3 + c] = w.params[k].grad()[c] + gtot[size_t(k) * 3 + c];
            draws += g_draws;
        }
        if (grad) {
            for (int j = 0; j < P * 3; ++j) {
                double acc = 0;
                for (int t = 0; t < n_threads; ++t) acc += gsum[size_t(t) * P * 3 + j];
                grad[j] = acc;
            }
        }
        if (draws_out) *draws_out = draws;
        return 0;
    } catch (...) {
        return -1;
    }
}

int drt_ref_render_mesh(const drtb_scene* s, const drtb_mesh* mesh, const drtb_render_opts* o,
                        const double* seed_img, double* img, double* grad,
                        int n_threads, int rand_mode, uint64_t* draws_out)
{
    return drt_ref_render_gimg(s, mesh, o, seed_img, img, grad, -1, nullptr, n_threads, rand_mode, draws_out);
}

int drt_ref_render(const drtb_scene* s, const drtb_render_opts* o,
                   const double* seed_img, double* img, double* grad,
                   int n_threads, int rand_mode, uint64_t* draws_out)
{
    return drt_ref_render_mesh(s, nullptr, o, seed_img, img, grad, n_threads, rand_mode, draws_out);
}

// Single explicit ray through the reference's Pathtracer::trace, for the
// drtb_trace_rays parity test.  jac (n_params x 3) = d radiance_c / d param_kc,
// obtained with three one-hot backward seeds on a fresh tape each (same key).
int drt_ref_trace_ray(const drtb_scene* s, const drtb_render_opts* o,
                      const double* orig, const double* dir, uint64_t key,
                      double* radiance, double* jac)
{
    try {
        World w(*s);
        Pathtracer<T> tracer(o->absorb, size_t(o->min_bounces));
        Vector<T, 3> og{orig[0], orig[1], orig[2]}, d{dir[0], dir[1], dir[2]};
        g_libc = 0;
        for (int c = 0; c < 3; ++c) {
            g_key = key;
            g_ctr = 2;                       // slots 0,1 belong to the camera
            Vector<T, 3, true> L = tracer.trace(w.scene, og, d);
            if (c == 0)
                for (int j = 0; j < 3; ++j) radiance[j] = L.detach()[j];
            if (!jac) break;
            Vector<T, 3> e(0);
            e[c] = 1;
            for (auto& p : w.params) p.grad() = Vector<T, 3>(0);
            L.backward(e);
            for (int k = 0; k < s->n_params; ++k)
                jac[k * 3 + c] = w.params[k].grad()[c];
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

int drt_ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"
