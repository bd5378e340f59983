// Host-side checks of the source-compatible include/drt headers: value vectors,
// the reverse-mode tape, custom backward functions, integrate(), the host
// conveniences of Shape/Camera, and the flattening that feeds the C ABI.
// No GPU needed; the GPU-dependent calls are checked to fail loudly.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "drt/bxdf.hpp"
#include "drt/camera.hpp"
#include "drt/constants.hpp"
#include "drt/emitter.hpp"
#include "drt/integrate.hpp"
#include "drt/pathtracer.hpp"
#include "drt/render.hpp"
#include "drt/shape.hpp"
#include "drt/vector.hpp"

using namespace drt;
using V = Vector<double, 3>;
using D = Vector<double, 3, true>;

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)
static bool close(double a, double b, double tol = 1e-12) { return std::fabs(a - b) <= tol * (1 + std::fabs(b)); }

int main(int argc, char**)
{
    // ---- plain vectors
    V a{1, 2, 3}, b{4, 5, 6};
    CHECK(close(dot(a, b), 32));
    CHECK(close(norm(V{3, 4, 0}), 5));
    V c = cross(a, b);
    CHECK(c[0] == -3 && c[1] == 6 && c[2] == -3);
    V n = normalize(V{0, 0, 2});
    CHECK(n[2] == 1);
    V r = reflect(V{1, -1, 0}, V{0, 1, 0});
    CHECK(r[0] == -1 && r[1] == -1);
    CHECK(((a + b) * 2.0 / 2)[1] == 7 && (-a)[0] == -1 && (a / b)[0] == 0.25 && (2 * a - b)[2] == 0);
    bool threw = false;
    try { V bad{1, 2}; (void)bad; } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    CHECK(a.size() == 3 && V(7)[1] == 7);

    // ---- tape: y = ((p * q) / 2 + p) * 3 - q / p
    D p(V{1, 2, 4}, true), q(V{3, 5, 7}, true), k(V{1, 1, 1});
    CHECK(p.requires_grad() && !k.requires_grad());
    D y = ((p * q) / 2 + p) * 3 - q / p;
    CHECK(close(y[0], (1 * 3 / 2.0 + 1) * 3 - 3));
    y.backward(V{1, 1, 1});
    for (int i = 0; i < 3; ++i) {
        double pv = p.detach()[i], qv = q.detach()[i];
        CHECK(close(p.grad()[i], 3 * (qv / 2 + 1) + qv / (pv * pv)));
        CHECK(close(q.grad()[i], 3 * pv / 2 - 1 / pv));
    }
    y.backward(V{1, 0, 0});                       // gradients accumulate
    CHECK(close(p.grad()[0], 2 * (3 * (3 / 2.0 + 1) + 3)) && close(p.grad()[1], 3 * (5 / 2.0 + 1) + 5 / 4.0));
    threw = false;
    try { k.grad(); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    D alias = p;                                   // copies alias one accumulator
    CHECK(alias.id() == p.id());
    D plain = k + k;                               // nothing tracked -> constant
    CHECK(!plain.requires_grad());
    p += q;                                        // compound ops rebind the handle
    CHECK(p.requires_grad() && close(p[0], 4));

    // ---- custom backward function (the reference README's extension point)
    D w(V{2, 2, 2}, true);
    D sq(w.detach() * w.detach(), [w](const V& g) { w.backward(2.0 * w.detach() * g); });
    (sq * 0.5).backward(V{1, 1, 1});
    CHECK(close(w.grad()[2], 2.0));

    // ---- integrate(): E[f(x)/pdf], x ~ U(0,1), f = theta * x  => theta / 2, d/dtheta = 1/2
    std::srand(1);
    D theta(V{3, 3, 3}, true);
    auto sampler = [] { return std::make_tuple(random::uniform(), 1.0); };
    auto f = [theta](double x) { return theta * x; };
    D est = integrate<double, 3>(f, sampler, 20000) / 20000;
    CHECK(std::fabs(est[0] - 1.5) < 0.03);
    est.backward(V{1, 1, 1});
    CHECK(std::fabs(theta.grad()[0] - 0.5) < 0.01);
    D theta2(V{3, 3, 3}, true);
    auto f2 = [theta2](double x) { return theta2 * x; };
    D un = integrate<double, 3>(f2, sampler, 20000, true) / 20000;   // fresh samples in backward
    un.backward(V{1, 1, 1});
    CHECK(std::fabs(un[0] - 1.5) < 0.03 && std::fabs(theta2.grad()[0] - 0.5) < 0.01);

    // ---- shapes: the reference's acceptance rules
    double t = 0;
    Plane<double> wall(V{1, 0, 0.1}, -3);
    CHECK(wall.intersect(V{0, 0, 0}, V{-1, 0, 0}, t) && close(t, 3));
    CHECK(!wall.intersect(V{0, 0, 0}, V{1, 0, 0}, t));
    CHECK(wall.normal(V{0, 0, 0})[2] == 0.1);     // returned un-normalised
    Sphere<double> ball(V{0, 0, 3}, 1);
    CHECK(ball.intersect(V{0, 0, 0}, V{0, 0, 1}, t) && close(t, 2));
    CHECK(ball.intersect(V{0, 0, 3}, V{0, 0, 1}, t) && close(t, 1));      // from inside: far root
    CHECK(!ball.intersect(V{0, 0, 5}, V{0, 0, 1}, t) && !ball.intersect(V{0, 2, 0}, V{0, 0, 1}, t));
    CHECK(ball.intersect(V{0, 0, 0}, V{0, 0, 2}, t) && close(t, 6 - std::sqrt(28.0)));    // a == 1 even for |d| = 2
    CHECK(close(norm(ball.normal(V{0, 1, 3.5})), 1));

    // ---- camera: look_at basis of src/render.cpp:65 and the pixel -> ray map
    Camera<double> cam(640, 480);
    cam.look_at(V{0, 0, 0}, V{0, 0, 1});
    CHECK(cam.forward()[2] == 1 && cam.right()[0] == -1 && cam.up()[1] == 1 && close(cam.aspect(), 4 / 3.0));
    auto [dir, pdf] = cam.sample(0, 0);
    CHECK(pdf == 1 && close(norm(dir), 1) && dir[0] > 0 && dir[1] > 0);   // row 0 = top, image left = +x

    // ---- diffuse BRDF host conveniences
    D albedo(V{0.5, 0.25, 1}, true);
    DiffuseBxDF<double> lam(albedo);
    D fr = lam(V{0, 1, 0}, V{0, 1, 0}, V{0, 1, 0});
    CHECK(close(fr[1], 0.25 / pi));
    auto [wo, pw] = lam.sample(V{0, 1, 0}, V{0, 1, 0});
    CHECK(wo[1] >= 0 && close(pw, wo[1] / pi, 1e-9));

    // ---- flattening: aliasing, order, null BxDF
    auto white = std::make_shared<DiffuseBxDF<double>>(albedo);
    D glow(V{1, 1, 1}, true);
    auto lamp = std::make_shared<AreaEmitter<double>>(glow);
    Sphere<double> s0(V{0, 0, 3}, 1, white), s1(V{0, 3, 3}, 1, nullptr, lamp);
    Plane<double> p0(V{0, 1, 0}, -3, white);
    Scene<double> scene{&s0, &p0, &s1};
    auto flat = gpu::flatten(scene);
    CHECK(flat.prims.size() == 3 && flat.materials.size() == 1 && flat.handles.size() == 2);
    CHECK(flat.prims[0].type == DRTB_SPHERE && flat.prims[1].type == DRTB_PLANE && flat.prims[1].material == 0);
    CHECK(flat.prims[2].material == -1 && flat.prims[2].emission == 1 && flat.params[1] == 0.25);
    auto spec = std::make_shared<SpecularBxDF<double>>(albedo, 30);
    Sphere<double> s2(V{0, 0, 3}, 1, spec);
    Scene<double> glossy_scene{&s2, &p0};          // SpecularBxDF flattens to DRTB_SPECULAR + its exponent
    auto gflat = gpu::flatten(glossy_scene);
    CHECK(gflat.materials.size() == 2 && gflat.materials[0].type == DRTB_SPECULAR && gflat.materials[0].exponent == 30);
    CHECK(gflat.materials[1].type == DRTB_DIFFUSE && gflat.materials[0].color == gflat.materials[1].color);
    Scene<double> bad_scene{&s2, nullptr};
    threw = false;
    try { gpu::flatten(bad_scene); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);

    // ---- Triangle<T> (new shape): host intersect with the rules of include/drtb.h, flattening into one drtb_mesh
    {
        auto tri_mat = std::make_shared<DiffuseBxDF<double>>(albedo);
        Triangle<double> tri(V{0, 0, 2}, V{1, 0, 2}, V{0, 1, 2}, tri_mat);
        double tt = -1;
        CHECK(tri.intersect(V{0.25, 0.25, 0}, V{0, 0, 1}, tt) && tt == 2.0);
        CHECK(tri.intersect(V{0, 0, 0}, V{0, 0, 1}, tt));                      // a vertex: inclusive bounds
        CHECK(tri.intersect(V{0.5, 0.5, 0}, V{0, 0, 1}, tt));                  // the hypotenuse: u + v == 1
        CHECK(!tri.intersect(V{0.6, 0.6, 0}, V{0, 0, 1}, tt));
        CHECK(!tri.intersect(V{0.25, 0.25, 3}, V{0, 0, 1}, tt));               // behind the origin: t < 0
        CHECK(tri.intersect(V{0.25, 0.25, 3}, V{0, 0, -1}, tt) && tt == 1.0);  // both sides are hit
        CHECK(!tri.intersect(V{0.25, 0.25, 0}, V{1, 0, 0}, tt));               // parallel: det == 0
        const V tn = tri.normal(V{0, 0, 2});
        CHECK(tn[0] == 0 && tn[1] == 0 && tn[2] == 1);
        Triangle<double> lamp_tri(V{0, 0, 5}, V{1, 0, 5}, V{0, 1, 5}, nullptr, lamp);
        Scene<double> mesh_scene{&s0, &p0, &tri, &lamp_tri};
        auto mf = gpu::flatten(mesh_scene);
        CHECK(mf.prims.size() == 2 && mf.tri_indices.size() == 6 && mf.tri_vertices.size() == 18);
        CHECK(mf.tri_vertices[3] == 1.0 && mf.tri_vertices[7] == 1.0 && mf.tri_vertices[2] == 2.0 && mf.tri_vertices[11] == 5.0);
        CHECK(mf.tri_color[0] == mf.materials[0].color && mf.tri_color[1] == -1);
        CHECK(mf.tri_emission[0] == -1 && mf.tri_emission[1] >= 0 && mf.tri_indices[5] == 5);
        Scene<double> wrong_order{&tri, &s0};       // an analytic shape after a triangle would change the tie-break order
        threw = false;
        try { gpu::flatten(wrong_order); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
        Triangle<double> glossy_tri(V{0, 0, 2}, V{1, 0, 2}, V{0, 1, 2}, std::make_shared<SpecularBxDF<double>>(albedo, 10));
        Scene<double> glossy_mesh{&glossy_tri};
        threw = false;
        try { gpu::flatten(glossy_mesh); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
    }

    // ---- no GPU => a loud exception, never a silent CPU render
    if (argc > 1 || drtb_device_count() == 0) {
        Pathtracer<double> tracer(0.5, 1);
        Camera<double> small(8, 8);
        V img[64];
        threw = false;
        try { render(scene, small, tracer, 1, img); } catch (const std::runtime_error& e) {
            threw = std::string(e.what()).find("no CPU fallback") != std::string::npos;
        }
        CHECK(threw);
        threw = false;
        try { tracer.trace(scene, V{0, 0, 0}, V{0, 0, 1}); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
    }
    std::printf(failures ? "%d FAILURES\n" : "all header checks passed\n", failures);
    return failures ? 1 : 0;
}
