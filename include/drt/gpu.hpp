// drt/gpu.hpp — the bridge from the include/drt object graph to the C ABI
// (include/drtb.h -> libdrtb.so -> sm_100a kernels).  NEW relative to the
// reference, which has no device code; everything here is host plumbing:
// flatten Scene<T>/Camera<T> into PODs, own one drtb_ctx per device, turn
// status codes into exceptions.  There is no CPU fallback: without a GPU every
// call throws std::runtime_error carrying drtb_last_error().
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include "../drtb.h"
#include "camera.hpp"
#include "shape.hpp"
#include "vector.hpp"

namespace drt {

template <typename T>
using Scene = std::vector<Shape<T>*>;        // non-owning, caller order == tie-break priority

namespace gpu {

// Scene<T> as the C ABI wants it, plus the handles the gradients go back to.
template <typename T>
struct FlatScene {
    std::vector<drtb_prim> prims;
    std::vector<drtb_material> materials;
    std::vector<double> params;                          // n x 3
    std::vector<Vector<T, 3, true>> handles;             // one per unique parameter node
    // Triangle<T> shapes of the scene, as one drtb_mesh (three vertices per triangle, unshared)
    std::vector<double> tri_vertices;                    // n_triangles x 9
    std::vector<int32_t> tri_indices, tri_color, tri_emission;

    int param_index(const Vector<T, 3, true>& h)
    {
        for (std::size_t k = 0; k < handles.size(); ++k)
            if (handles[k].id() == h.id()) return int(k);  // aliases share one accumulator
        handles.push_back(h);
        for (int c = 0; c < 3; ++c) params.push_back(double(h.detach()[c]));
        return int(handles.size()) - 1;
    }
};

template <typename T>
FlatScene<T> flatten(const Scene<T>& scene)
{
    FlatScene<T> f;
    std::vector<const BxDF<T>*> seen;
    for (Shape<T>* s : scene) {
        if (!s) throw std::runtime_error("drt::gpu::flatten: null shape in scene");
        drtb_prim p{};
        s->describe(p);
        p.material = -1;
        p.emission = -1;
        double tv[9];
        if (s->describe_triangle(tv)) {
            // drtb_mesh: per-triangle DiffuseBxDF albedo and AreaEmitter, by parameter index
            const BxDF<T>* b = s->bxdf();
            if (b && b->kind() != BxDFKind::Diffuse) throw std::runtime_error("drt::gpu::flatten: triangles take a DiffuseBxDF (or none)");
            const int32_t base = int32_t(f.tri_indices.size());
            f.tri_vertices.insert(f.tri_vertices.end(), tv, tv + 9);
            for (int k = 0; k < 3; ++k) f.tri_indices.push_back(base + k);
            f.tri_color.push_back(b ? f.param_index(b->color()) : -1);
            f.tri_emission.push_back(s->emitter() ? f.param_index(s->emitter()->emission()) : -1);
            continue;
        }
        // scene order is the tie-break priority (pathtracer.hpp:80) and the library scans analytic primitives
        // before the mesh: a plane or sphere listed after a triangle would change who wins an exact tie
        if (!f.tri_indices.empty()) throw std::runtime_error("drt::gpu::flatten: list planes and spheres before triangles in the Scene");
        if (const BxDF<T>* b = s->bxdf()) {
            int m = -1;
            for (std::size_t i = 0; i < seen.size(); ++i)
                if (seen[i] == b) m = int(i);
            if (m < 0) {
                m = int(seen.size());
                seen.push_back(b);
                const bool spec = b->kind() == BxDFKind::Specular;
                f.materials.push_back(drtb_material{spec ? DRTB_SPECULAR : DRTB_DIFFUSE, f.param_index(b->color()),
                                                    spec ? b->exponent() : 0.0});
            }
            p.material = m;
        }
        if (const Emitter<T>* e = s->emitter()) p.emission = f.param_index(e->emission());
        f.prims.push_back(p);
    }
    return f;
}

template <typename T>
drtb_camera flatten(const Camera<T>& cam)
{
    drtb_camera c{};
    c.width = int32_t(cam.width());
    c.height = int32_t(cam.height());
    c.vfov = cam.vfov();
    for (int i = 0; i < 3; ++i) {
        c.eye[i] = double(cam.eye()[i]);
        c.forward[i] = double(cam.forward()[i]);
        c.right[i] = double(cam.right()[i]);
        c.up[i] = double(cam.up()[i]);
    }
    return c;
}

// One drtb_ctx per device, created on first use, shared by every call.
class Device {
    drtb_ctx* ctx_ = nullptr;
    std::vector<unsigned char> uploaded_;                 // bytes of the last uploaded geometry
    std::vector<double> uploaded_params_;
    std::vector<double> uploaded_tris_;                   // vertices of the attached mesh
    std::vector<int32_t> uploaded_tri_mat_;

    [[noreturn]] void raise(const char* what, int rc) const
    {
        const char* m = drtb_last_error(ctx_);
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (m ? m : ""));
    }

public:
    std::mutex lock;                                       // a ctx serves one host thread at a time

    explicit Device(int index)
    {
        int rc = drtb_create(index, &ctx_);
        if (rc != DRTB_OK) raise("drtb_create", rc);
    }
    ~Device() { drtb_destroy(ctx_); }
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
    drtb_ctx* ctx() { return ctx_; }
    void check(const char* what, int rc) const { if (rc != DRTB_OK) raise(what, rc); }

    // Upload only what changed: an optimisation loop that moves parameters
    // between renders pays for drtb_set_params, not for a scene rebuild.
    template <typename T>
    void sync(const FlatScene<T>& f, const drtb_camera& cam)
    {
        std::vector<unsigned char> bytes(sizeof cam + f.prims.size() * sizeof(drtb_prim) +
                                         f.materials.size() * sizeof(drtb_material));
        unsigned char* w = bytes.data();
        std::memcpy(w, &cam, sizeof cam); w += sizeof cam;
        if (!f.prims.empty()) std::memcpy(w, f.prims.data(), f.prims.size() * sizeof(drtb_prim));
        w += f.prims.size() * sizeof(drtb_prim);
        if (!f.materials.empty()) std::memcpy(w, f.materials.data(), f.materials.size() * sizeof(drtb_material));
        std::vector<int32_t> tri_mat(f.tri_color);
        tri_mat.insert(tri_mat.end(), f.tri_emission.begin(), f.tri_emission.end());
        const bool mesh_changed = f.tri_vertices != uploaded_tris_ || tri_mat != uploaded_tri_mat_;
        if (bytes != uploaded_ || f.params.size() != uploaded_params_.size() || mesh_changed) {
            drtb_scene s{};
            s.prims = f.prims.data(); s.n_prims = int32_t(f.prims.size());
            s.materials = f.materials.data(); s.n_materials = int32_t(f.materials.size());
            s.params = f.params.data(); s.n_params = int32_t(f.params.size() / 3);
            s.camera = cam;
            check("drtb_scene_upload", drtb_scene_upload(ctx_, &s));
            if (!f.tri_indices.empty()) {                    // Triangle<T> shapes: one mesh, BVH built on the GPU
                drtb_mesh mesh{};
                mesh.vertices = f.tri_vertices.data(); mesh.n_vertices = int64_t(f.tri_vertices.size() / 3);
                mesh.indices = f.tri_indices.data(); mesh.n_triangles = int64_t(f.tri_indices.size() / 3);
                mesh.color = f.tri_color.data(); mesh.emission = f.tri_emission.data();
                check("drtb_mesh_upload", drtb_mesh_upload(ctx_, &mesh));
            }
            uploaded_ = std::move(bytes);
            uploaded_params_ = f.params;
            uploaded_tris_ = f.tri_vertices;
            uploaded_tri_mat_ = std::move(tri_mat);
        } else if (f.params != uploaded_params_) {
            check("drtb_set_params", drtb_set_params(ctx_, f.params.data(), int32_t(f.params.size() / 3)));
            uploaded_params_ = f.params;
        }
    }
};

// The GPUs of one box behind one handle (drtb_multi_*): image bands are spread over the devices, the image is
// assembled on the first one by the render kernels themselves, gradients are summed over the devices.
class MultiDevice {
    drtb_multi* m_ = nullptr;

    [[noreturn]] void raise(const char* what, int rc) const
    {
        const char* msg = drtb_multi_last_error(m_);
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (msg ? msg : ""));
    }

public:
    std::mutex lock;
    explicit MultiDevice(const std::vector<int>& devices)
    {
        int rc = drtb_multi_create(devices.data(), int32_t(devices.size()), &m_);
        if (rc != DRTB_OK) raise("drtb_multi_create", rc);
    }
    ~MultiDevice() { drtb_multi_destroy(m_); }
    MultiDevice(const MultiDevice&) = delete;
    MultiDevice& operator=(const MultiDevice&) = delete;
    drtb_multi* handle() { return m_; }
    void check(const char* what, int rc) const { if (rc != DRTB_OK) raise(what, rc); }

    template <typename T>
    void sync(const FlatScene<T>& f, const drtb_camera& cam)
    {
        drtb_scene s{};
        s.prims = f.prims.data(); s.n_prims = int32_t(f.prims.size());
        s.materials = f.materials.data(); s.n_materials = int32_t(f.materials.size());
        s.params = f.params.data(); s.n_params = int32_t(f.params.size() / 3);
        s.camera = cam;
        check("drtb_multi_scene_upload", drtb_multi_scene_upload(m_, &s));
        if (!f.tri_indices.empty()) {
            drtb_mesh mesh{};
            mesh.vertices = f.tri_vertices.data(); mesh.n_vertices = int64_t(f.tri_vertices.size() / 3);
            mesh.indices = f.tri_indices.data(); mesh.n_triangles = int64_t(f.tri_indices.size() / 3);
            mesh.color = f.tri_color.data(); mesh.emission = f.tri_emission.data();
            check("drtb_multi_mesh_upload", drtb_multi_mesh_upload(m_, &mesh));
        }
    }
};

inline MultiDevice& multi_device(const std::vector<int>& devices)
{
    static std::mutex m;
    static std::map<std::vector<int>, std::unique_ptr<MultiDevice>> groups;
    std::lock_guard<std::mutex> g(m);
    auto& slot = groups[devices];
    if (!slot) slot.reset(new MultiDevice(devices));
    return *slot;
}

inline Device& device(int index = 0)
{
    static std::mutex m;
    static std::map<int, std::unique_ptr<Device>> devices;
    std::lock_guard<std::mutex> g(m);
    auto& slot = devices[index];
    if (!slot) slot.reset(new Device(index));
    return *slot;
}

// An image buffer in pinned host memory (drtb_host_alloc): drt::render / drtb_render let the kernel store the pixels
// straight into it, so no device->host copy follows the render.  NEW relative to the reference, which renders into
// a plain `new Vector<double, 3>[w * h]` (src/render.cpp:69) -- that works here as well and costs the copy.
class PinnedImage {
public:
    explicit PinnedImage(std::size_t pixels) : n_(pixels)
    {
        void* p = nullptr;
        if (drtb_host_alloc(std::max<std::size_t>(pixels, 1) * sizeof(Vector<double, 3>), &p) != DRTB_OK)
            throw std::runtime_error("drt::gpu::PinnedImage: drtb_host_alloc failed");
        data_ = static_cast<Vector<double, 3>*>(p);
        for (std::size_t i = 0; i < n_; ++i) new (data_ + i) Vector<double, 3>();
    }
    ~PinnedImage() { drtb_host_free(data_); }
    PinnedImage(const PinnedImage&) = delete;
    PinnedImage& operator=(const PinnedImage&) = delete;
    Vector<double, 3>* data() { return data_; }
    const Vector<double, 3>* data() const { return data_; }
    std::size_t size() const { return n_; }
    Vector<double, 3>& operator[](std::size_t i) { return data_[i]; }
    const Vector<double, 3>& operator[](std::size_t i) const { return data_[i]; }

private:
    Vector<double, 3>* data_ = nullptr;
    std::size_t n_ = 0;
};

} // namespace gpu
} // namespace drt
