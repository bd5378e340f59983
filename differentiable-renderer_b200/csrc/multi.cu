// multi.cu — the GPUs of one NVSwitch box behind ONE handle, in ONE process (SURVEY.md §8b/§8e:
// `drtb_create(const int* devices, int n, ...)`, gradients "summed over GPUs").  Host code only: it drives one
// drtb_ctx per device through the same C ABI a single-GPU caller uses.
//
// The pixel loop of src/render.cpp:72-86 is cut into bands of image rows (band b -> device b mod n, include/drtb.h
// shard_*).  Analytic scenes: every device's render kernel stores its finished pixels straight into the FULL image
// that lives on the first device (peer stores over NVLink, drtb_set_image_peers), so the image is assembled by the
// kernels and leaves the box in one device-to-host copy.  Mesh scenes (wavefront pipeline): every device renders its
// compact bands, which are copied to their rows of the host image.  Gradients: P x 3 doubles per device, added on
// the host in device order (the device-resident sum without a host round trip is drtb_set_grad_peers).
#include <new>

#include "host.hpp"

using namespace drtbh;

struct drtb_multi {
    int n = 0;
    int devices[drtb::kMaxPeers] = {};
    drtb_ctx* ctx[drtb::kMaxPeers] = {};
    cudaStream_t stream[drtb::kMaxPeers] = {};
    double* d_full = nullptr;         size_t full_cap = 0;       // full image on devices[0] (analytic scenes)
    double* d_img[drtb::kMaxPeers] = {};   size_t img_cap[drtb::kMaxPeers] = {};     // compact shard images (mesh scenes)
    double* d_seed[drtb::kMaxPeers] = {};  size_t seed_cap[drtb::kMaxPeers] = {};
    double* d_grad[drtb::kMaxPeers] = {};  size_t grad_cap[drtb::kMaxPeers] = {};
    drtb_stats* d_stats[drtb::kMaxPeers] = {};
    std::vector<double> h_seed, h_grad;
    bool peer_ok = true;              // every device can store into devices[0]
    bool has_mesh = false;
    int W = 0, H = 0, P3 = 0;
    std::string err;
};

namespace {

thread_local std::string g_multi_create_err;

int mfail(drtb_multi* m, int code, const std::string& msg)
{
    if (m) m->err = msg; else g_multi_create_err = msg;
    return code;
}

#define MCK(m, call)                                                                                         \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return mfail(m, DRTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
int grow(drtb_multi* m, T*& p, size_t& cap, size_t n)
{
    if (n <= cap && p) return DRTB_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (cudaMalloc((void**)&p, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return mfail(m, DRTB_ERR_NOMEM, "cudaMalloc failed"); }
    cap = n;
    return DRTB_OK;
}

int sub(drtb_multi* m, int i, int rc)                 // a device's context failed: carry its message up
{
    if (rc != DRTB_OK) m->err = "device " + std::to_string(m->devices[i]) + ": " + drtb_last_error(m->ctx[i]);
    return rc;
}

} // namespace

extern "C" {

int drtb_multi_create(const int* devices, int32_t n, drtb_multi** out)
{
    if (!out) return mfail(nullptr, DRTB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n < 1 || n > drtb::kMaxPeers) return mfail(nullptr, DRTB_ERR_INVALID, "between 1 and 8 devices");
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return mfail(nullptr, DRTB_ERR_INVALID, "a device is listed twice");
    drtb_multi* m = new (std::nothrow) drtb_multi;
    if (!m) return mfail(nullptr, DRTB_ERR_NOMEM, "out of host memory");
    m->n = n;
    for (int i = 0; i < n; ++i) {
        m->devices[i] = devices[i];
        int rc = drtb_create(devices[i], &m->ctx[i]);
        if (rc != DRTB_OK) {
            const std::string msg = drtb_last_error(nullptr);
            drtb_multi_destroy(m);
            return mfail(nullptr, rc, "device " + std::to_string(devices[i]) + ": " + msg);
        }
        if (cudaSetDevice(devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc((void**)&m->d_stats[i], sizeof(drtb_stats)) != cudaSuccess) {
            const std::string msg = cudaGetErrorString(cudaGetLastError());
            drtb_multi_destroy(m);
            return mfail(nullptr, DRTB_ERR_CUDA, "device setup failed: " + msg);
        }
        if (i > 0) {                                   // this device stores pixels into the first device's full image
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[0]);
            if (can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
                cudaGetLastError();
            }
            m->peer_ok = m->peer_ok && can != 0;
        }
    }
    *out = m;
    return DRTB_OK;
}

void drtb_multi_destroy(drtb_multi* m)
{
    if (!m) return;
    for (int i = 0; i < m->n; ++i) {
        cudaSetDevice(m->devices[i]);
        if (m->stream[i]) { cudaStreamSynchronize(m->stream[i]); cudaStreamDestroy(m->stream[i]); }
        cudaFree(m->d_img[i]); cudaFree(m->d_seed[i]); cudaFree(m->d_grad[i]); cudaFree(m->d_stats[i]);
        if (i == 0) cudaFree(m->d_full);
        if (m->ctx[i]) drtb_destroy(m->ctx[i]);
    }
    delete m;
}

const char* drtb_multi_last_error(const drtb_multi* m) { return m ? m->err.c_str() : g_multi_create_err.c_str(); }

int32_t drtb_multi_device_count(const drtb_multi* m) { return m ? m->n : 0; }

int drtb_multi_scene_upload(drtb_multi* m, const drtb_scene* scene)
{
    if (!m) return DRTB_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        const int rc = sub(m, i, drtb_scene_upload(m->ctx[i], scene));
        if (rc != DRTB_OK) return rc;
    }
    m->W = scene->camera.width; m->H = scene->camera.height; m->P3 = scene->n_params * 3;
    m->has_mesh = false;
    return DRTB_OK;
}

int drtb_multi_mesh_upload(drtb_multi* m, const drtb_mesh* mesh)
{
    if (!m) return DRTB_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {                  // every device builds its own copy of the BVH
        const int rc = sub(m, i, drtb_mesh_upload(m->ctx[i], mesh));
        if (rc != DRTB_OK) return rc;
    }
    m->has_mesh = mesh && mesh->n_triangles > 0;
    return DRTB_OK;
}

int drtb_multi_set_params(drtb_multi* m, const double* params, int32_t n_params)
{
    if (!m) return DRTB_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        const int rc = sub(m, i, drtb_set_params(m->ctx[i], params, n_params));
        if (rc != DRTB_OK) return rc;
    }
    return DRTB_OK;
}

int drtb_multi_render(drtb_multi* m, const drtb_render_opts* opts, const double* seed_img, double* img, double* grad,
                      drtb_stats* stats)
{
    if (!m) return DRTB_ERR_INVALID;
    if (!opts) return mfail(m, DRTB_ERR_INVALID, "opts is NULL");
    if (m->W < 1) return mfail(m, DRTB_ERR_INVALID, "no scene uploaded");
    const int n = m->n, W = m->W, H = m->H, P3 = m->P3;
    const bool want_img = (opts->flags & DRTB_FLAG_IMAGE) != 0, want_grad = (opts->flags & DRTB_FLAG_GRAD) != 0;
    if (want_img && !img) return mfail(m, DRTB_ERR_INVALID, "DRTB_FLAG_IMAGE set but img is NULL");
    if (want_grad && !grad) return mfail(m, DRTB_ERR_INVALID, "DRTB_FLAG_GRAD set but grad is NULL");
    const int band = opts->band_rows > 0 ? opts->band_rows : 8;
    const size_t row3 = size_t(W) * 3;
    // analytic scenes on peer-connected devices: the kernels assemble the image on the first device
    const bool fused = want_img && n > 1 && m->peer_ok && !m->has_mesh;
    int rc;
    if (want_img && (fused || n == 1)) {
        MCK(m, cudaSetDevice(m->devices[0]));
        if ((rc = grow(m, m->d_full, m->full_cap, std::max<size_t>(size_t(H) * row3, 3))) != DRTB_OK) return rc;
    }
    m->h_grad.assign(size_t(n) * std::max(P3, 1), 0.0);
    drtb_render_opts o[drtb::kMaxPeers];
    int rows[drtb::kMaxPeers];
    // enqueue everything, device by device; nothing below blocks until every device has its work
    for (int i = 0; i < n; ++i) {
        MCK(m, cudaSetDevice(m->devices[i]));
        o[i] = *opts;
        o[i].shard_index = i; o[i].shard_count = n; o[i].band_rows = band;
        if (stats) o[i].flags |= DRTB_FLAG_STATS;
        rows[i] = drtb_shard_rows(H, i, n, band);
        const size_t px3 = std::max<size_t>(size_t(rows[i]) * row3, 3);
        double* d_seed = nullptr;
        if (seed_img && want_grad) {                  // this device's bands of the per-pixel adjoint seed, compact
            if ((rc = grow(m, m->d_seed[i], m->seed_cap[i], px3)) != DRTB_OK) return rc;
            int r = 0;
            for (int b0 = i * band; b0 < H; b0 += n * band) {
                const int nb = std::min(band, H - b0);
                MCK(m, cudaMemcpyAsync(m->d_seed[i] + size_t(r) * row3, seed_img + size_t(b0) * row3, sizeof(double) * nb * row3,
                                       cudaMemcpyHostToDevice, m->stream[i]));
                r += nb;
            }
            d_seed = m->d_seed[i];
        }
        if (want_grad && (rc = grow(m, m->d_grad[i], m->grad_cap[i], std::max<size_t>(P3, 3))) != DRTB_OK) return rc;
        double* d_img = nullptr;
        if (want_img) {
            if (n == 1) d_img = m->d_full;
            else if (fused) {
                double* peers[1] = {m->d_full};
                if ((rc = sub(m, i, drtb_set_image_peers(m->ctx[i], peers, 1))) != DRTB_OK) return rc;
            } else {
                if ((rc = grow(m, m->d_img[i], m->img_cap[i], px3)) != DRTB_OK) return rc;
                d_img = m->d_img[i];
            }
        }
        if (!fused && (rc = sub(m, i, drtb_set_image_peers(m->ctx[i], nullptr, 0))) != DRTB_OK) return rc;
        if ((rc = sub(m, i, drtb_reserve(m->ctx[i], &o[i]))) != DRTB_OK) return rc;     // allocation / first-use work, off the clock
    }
    cudaEvent_t ev0[drtb::kMaxPeers] = {}, ev1[drtb::kMaxPeers] = {};
    for (int i = 0; i < n; ++i) {
        MCK(m, cudaSetDevice(m->devices[i]));
        if (stats) { MCK(m, cudaEventCreate(&ev0[i])); MCK(m, cudaEventCreate(&ev1[i])); MCK(m, cudaEventRecord(ev0[i], m->stream[i])); }
        double* d_img = !want_img ? nullptr : n == 1 ? m->d_full : fused ? nullptr : m->d_img[i];
        rc = sub(m, i, drtb_render_device(m->ctx[i], &o[i], (seed_img && want_grad) ? m->d_seed[i] : nullptr, d_img,
                                          want_grad ? m->d_grad[i] : nullptr, stats ? m->d_stats[i] : nullptr, m->stream[i]));
        if (rc != DRTB_OK) return rc;
        if (stats) MCK(m, cudaEventRecord(ev1[i], m->stream[i]));
        if (want_grad && P3)
            MCK(m, cudaMemcpyAsync(m->h_grad.data() + size_t(i) * P3, m->d_grad[i], sizeof(double) * P3, cudaMemcpyDeviceToHost, m->stream[i]));
        if (want_img && n > 1 && !fused) {            // compact bands -> their rows of the host image
            int r = 0;
            for (int b0 = i * band; b0 < H; b0 += n * band) {
                const int nb = std::min(band, H - b0);
                MCK(m, cudaMemcpyAsync(img + size_t(b0) * row3, m->d_img[i] + size_t(r) * row3, sizeof(double) * nb * row3,
                                       cudaMemcpyDeviceToHost, m->stream[i]));
                r += nb;
            }
        }
    }
    drtb_stats total{};
    for (int i = 0; i < n; ++i) {
        MCK(m, cudaSetDevice(m->devices[i]));
        MCK(m, cudaStreamSynchronize(m->stream[i]));
        if (stats) {
            drtb_stats s{};
            MCK(m, cudaMemcpy(&s, m->d_stats[i], sizeof s, cudaMemcpyDeviceToHost));
            float ms = 0.f;
            MCK(m, cudaEventElapsedTime(&ms, ev0[i], ev1[i]));
            cudaEventDestroy(ev0[i]); cudaEventDestroy(ev1[i]);
            total.segments += s.segments; total.lit_paths += s.lit_paths; total.truncated_paths += s.truncated_paths;
            total.bvh_nodes += s.bvh_nodes; total.tri_tests += s.tri_tests;
            total.paths += uint64_t(rows[i]) * W * opts->spp;
            total.kernel_ms = std::max(total.kernel_ms, double(ms));      // the devices run side by side
        }
    }
    // every device's pixels have landed in the first device's full image (all streams are synchronised)
    if (want_img && (fused || n == 1)) {
        MCK(m, cudaSetDevice(m->devices[0]));
        MCK(m, cudaMemcpy(img, m->d_full, sizeof(double) * size_t(H) * row3, cudaMemcpyDeviceToHost));
    }
    if (want_grad)
        for (int j = 0; j < P3; ++j) {
            double v = 0.0;
            for (int i = 0; i < n; ++i) v += m->h_grad[size_t(i) * P3 + j];     // device order: reproducible
            grad[j] = v;
        }
    if (stats) *stats = total;
    return DRTB_OK;
}

} // extern "C"
