// drt/pathtracer.hpp — Scene<T> and Pathtracer<T> (reference pathtracer.hpp:12-136).
//
// Same constructor and trace() signature; the recursion, the per-segment tape
// and the virtual raycast of the reference are gone.  trace() is a batch of ONE
// through drtb_trace_rays (kept for API compatibility -- it costs a kernel
// launch per ray); whole images go through drt::render() (drt/render.hpp),
// which replaces the pixel loop of src/render.cpp:72-86 with one launch.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>
#include "gpu.hpp"
#include "random.hpp"
#include "vector.hpp"

namespace drt {

template <typename T>
class Pathtracer {
    double absorb_;
    std::size_t min_bounces_;

public:
    // Russian roulette: from depth >= min_bounces on, a path is absorbed with
    // probability `absorb` and survivors are weighted by 1/(1 - absorb).
    // "B bounces exactly" is Pathtracer(1.0, B).
    Pathtracer(double absorb, std::size_t min_bounces) : absorb_(absorb), min_bounces_(min_bounces) {}

    double absorb() const { return absorb_; }
    std::size_t min_bounces() const { return min_bounces_; }

    // Radiance arriving at `orig` from direction `dir` (used as given, not
    // normalised).  The result is differentiable: backward(g) adds
    // g . d(radiance)/d(param) to every parameter the scene references.
    // `depth` shifts the roulette threshold exactly as the reference's
    // recursion depth argument does.
    Vector<T, 3, true> trace(const Scene<T>& scene, Vector<T, 3> orig, Vector<T, 3> dir,
                             std::size_t depth = 0, int device_index = 0) const
    {
        gpu::FlatScene<T> flat = gpu::flatten(scene);
        drtb_camera cam{};                                 // unused by explicit rays
        cam.width = cam.height = 1;
        cam.vfov = 1.0;
        cam.forward[2] = 1.0; cam.right[0] = 1.0; cam.up[1] = 1.0;
        drtb_render_opts o{};
        o.spp = 1;
        o.min_bounces = int32_t(depth >= min_bounces_ ? 0 : min_bounces_ - depth);
        o.absorb = absorb_;
        o.precision = DRTB_F64;
        o.flags = DRTB_FLAG_IMAGE | DRTB_FLAG_GRAD;
        o.seed_scale = 1.0;
        const double og[3] = {double(orig[0]), double(orig[1]), double(orig[2])};
        const double dr[3] = {double(dir[0]), double(dir[1]), double(dir[2])};
        const std::uint64_t key = random::next_ray_key();
        double L[3] = {0, 0, 0};
        auto jac = std::make_shared<std::vector<double>>(flat.params.size(), 0.0);
        {
            gpu::Device& dev = gpu::device(device_index);
            std::lock_guard<std::mutex> g(dev.lock);
            dev.sync(flat, cam);
            dev.check("drtb_trace_rays",
                      drtb_trace_rays(dev.ctx(), &o, 1, og, dr, &key, L, jac->empty() ? nullptr : jac->data()));
        }
        Vector<T, 3> value{T(L[0]), T(L[1]), T(L[2])};
        auto handles = flat.handles;
        return Vector<T, 3, true>(value, [handles, jac](const Vector<T, 3>& g) {
            for (std::size_t k = 0; k < handles.size(); ++k) {
                Vector<T, 3> gk;                            // channels never mix
                for (int c = 0; c < 3; ++c) gk[c] = g[c] * T((*jac)[3 * k + c]);
                handles[k].backward(gk);
            }
        });
    }
};

} // namespace drt
