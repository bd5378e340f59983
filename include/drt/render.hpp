// drt/render.hpp — drt::render(): the pixel loop of the reference application
// (src/render.cpp:72-86: cam.sample -> tracer.trace -> accumulate ->
// radiance.backward(seed)) as ONE call into the CUDA path.
//
//   Vector<double,3>* img = new Vector<double,3>[w*h];
//   drt::render(scene, cam, tracer, samples, img);      // image + gradients
//   red.grad();                                          // as after .backward()
//
// NEW relative to the reference (which spells the loop out in main()).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <vector>
#include "camera.hpp"
#include "gpu.hpp"
#include "pathtracer.hpp"
#include "vector.hpp"

namespace drt {

struct RenderOptions {
    bool gradients = true;                         // run the adjoint (radiance.backward)
    Vector<double, 3> seed = Vector<double, 3>(1); // the per-sample cotangent, src/render.cpp:80
    const Vector<double, 3>* seed_image = nullptr; // optional per-pixel cotangent (w*h), multiplies `seed`
    double seed_scale = 1.0;                       // e.g. 1/samples for d(loss)/d(pixel) seeds
    std::uint64_t stream = 0;                      // independent sample streams
    std::uint64_t adjoint_stream = 0;              // != 0: decorrelated adjoint (fresh samples in backward)
    int precision = DRTB_F64;                      // DRTB_F64 (parity) | DRTB_F32 (throughput)
    int device = 0;
    // More than one entry: the image rows are spread over these GPUs of one box (bands of 8 rows), the image is
    // assembled on the first and the gradients are summed over all of them (drtb_multi_render).  Same results.
    std::vector<int> devices;
    drtb_stats* stats = nullptr;
    // Per-pixel gradient image (README.md:138-145): if grad_image != nullptr it receives, per
    // pixel, the share of grad_image_of.grad() that the pixel's samples contributed (w*h entries).
    Vector<double, 3>* grad_image = nullptr;
    const void* grad_image_of = nullptr;           // &parameter (any Vector<T,3,true> handle of it)
};

static_assert(sizeof(Vector<double, 3>) == 3 * sizeof(double), "Vector<double,3> must be 3 packed doubles");

// img: width*height, row-major, row 0 = top (may be nullptr to skip the image).
// Gradients are ADDED to every parameter created with requires_grad = true,
// unnormalised (sum over all samples of all pixels), like the reference's tape.
template <typename T>
void render(const Scene<T>& scene, const Camera<T>& cam, const Pathtracer<T>& tracer, std::size_t samples,
            Vector<double, 3>* img, const RenderOptions& opt = RenderOptions())
{
    gpu::FlatScene<T> flat = gpu::flatten(scene);
    const drtb_camera c = gpu::flatten(cam);
    drtb_render_opts o{};
    o.spp = int32_t(samples);
    o.min_bounces = int32_t(tracer.min_bounces());
    o.absorb = tracer.absorb();
    o.seed = opt.stream;
    o.precision = opt.precision;
    o.flags = (img ? DRTB_FLAG_IMAGE : 0u) | (opt.gradients ? DRTB_FLAG_GRAD : 0u);
    o.seed_scale = opt.seed_scale;
    o.adjoint_seed = opt.adjoint_stream;
    std::vector<double> grad(flat.params.size(), 0.0);
    int gparam = -1;
    if (opt.grad_image) {
        if (!opt.gradients || !opt.grad_image_of) throw std::runtime_error("drt::render: grad_image needs gradients and grad_image_of");
        const auto& h = *static_cast<const Vector<T, 3, true>*>(opt.grad_image_of);
        for (std::size_t k = 0; k < flat.handles.size(); ++k)
            if (flat.handles[k].id() == h.id()) gparam = int(k);
        if (gparam < 0) throw std::runtime_error("drt::render: grad_image_of is not a parameter of this scene");
    }
    if (opt.devices.size() > 1) {
        if (opt.grad_image) throw std::runtime_error("drt::render: grad_image is a single-device call");
        gpu::MultiDevice& md = gpu::multi_device(opt.devices);
        std::lock_guard<std::mutex> g(md.lock);
        md.sync(flat, c);
        md.check("drtb_multi_render",
                 drtb_multi_render(md.handle(), &o, reinterpret_cast<const double*>(opt.seed_image),
                                   reinterpret_cast<double*>(img), opt.gradients ? grad.data() : nullptr, opt.stats));
    } else {
        gpu::Device& dev = gpu::device(opt.devices.size() == 1 ? opt.devices[0] : opt.device);
        std::lock_guard<std::mutex> g(dev.lock);
        dev.sync(flat, c);
        if (opt.grad_image)
            dev.check("drtb_render_grad_image",
                      drtb_render_grad_image(dev.ctx(), &o, gparam, reinterpret_cast<const double*>(opt.seed_image),
                                             reinterpret_cast<double*>(img), grad.data(),
                                             reinterpret_cast<double*>(opt.grad_image), opt.stats));
        else
            dev.check("drtb_render",
                      drtb_render(dev.ctx(), &o, reinterpret_cast<const double*>(opt.seed_image),
                                  reinterpret_cast<double*>(img), opt.gradients ? grad.data() : nullptr, opt.stats));
    }
    if (!opt.gradients) return;
    for (std::size_t k = 0; k < flat.handles.size(); ++k) {
        if (!flat.handles[k].is_leaf()) continue;          // constants have no grad()
        Vector<T, 3> gk;
        for (int ch = 0; ch < 3; ++ch) gk[ch] = T(opt.seed[ch] * grad[3 * k + ch]);   // channels never mix
        flat.handles[k].backward(gk);                      // VariableNode-style +=
    }
}

} // namespace drt
