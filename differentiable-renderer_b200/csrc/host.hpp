// host.hpp — what the translation units of libdrtb.so share on the host side: the context, error
// plumbing, the chunk planner and the entry points each .cu file offers the others.
//
//   drtb.cu        the C ABI of include/drtb.h, scene flattening, gradient reduction, explicit rays
//   render_f64.cu  render_kernel / render_regen_kernel in IEEE double (the parity instantiation)
//   render_f32.cu  the same kernels in float (throughput instantiation, and DRTB_MIXED's fast pass)
//   mesh.cu        triangle meshes: GPU BVH build, the wavefront integrator
//
// One .cu per group keeps every kernel's code generation independent of the others and lets the
// build compile them in parallel.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>          // header-only NVTX 3: ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached

#include "path.cuh"

// resident blocks per SM the render kernels are compiled for (build.py passes the measured choices)
#ifndef DRTB_MIN_BLOCKS
#define DRTB_MIN_BLOCKS 1
#endif
#ifndef DRTB_MESH_MIN_BLOCKS
#define DRTB_MESH_MIN_BLOCKS DRTB_MIN_BLOCKS
#endif
#ifndef DRTB_MIN_BLOCKS_F32
#define DRTB_MIN_BLOCKS_F32 DRTB_MIN_BLOCKS
#endif
#ifndef DRTB_MIN_BLOCKS_GEN
#define DRTB_MIN_BLOCKS_GEN (DRTB_MIN_BLOCKS < 5 ? DRTB_MIN_BLOCKS : 5)   // double GEN kernels (lobe code, gradient image): 96 registers
#endif

struct drtb_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    bool has_scene = false;
    bool has_specular = false;    // some primitive carries a DRTB_SPECULAR material
    bool const_weight = true;     // every plane's diffuse weight is a constant of the plane (drtb_scene_upload)
    std::vector<drtb_prim> prims;
    std::vector<drtb_material> materials;
    std::vector<double> params;
    drtb_camera camera{};
    drtb::DevScene<double> sc64{};
    drtb::DevScene<float> sc32{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double* d_params = nullptr;   size_t params_cap = 0;
    double* d_partial = nullptr;  size_t partial_cap = 0;
    double* d_ring = nullptr;     size_t ring_cap = 0;      // lit-path rings of the QUEUE >= 2 kernels
    int ring_policy = 0;          // DRTB_RING=global: lit-path ring in global memory at every depth (A/B aid)
    bool no_regen = false;        // DRTB_NO_REGEN=1: Russian-roulette renders without path regeneration (A/B aid)
    double* d_img = nullptr;      size_t img_cap = 0;
    double* d_seed = nullptr;     size_t seed_cap = 0;
    double* d_grad = nullptr;     size_t grad_cap = 0;
    double* d_gimg = nullptr;     size_t gimg_cap = 0;
    drtb_stats* d_stats = nullptr;
    unsigned long long launches = 0;
    // triangle mesh + BVH (device)
    int64_t n_tris = 0;
    float4* d_nodes = nullptr;    // raw 16-byte words of the wide nodes (bvh.cuh); lives in d_geom behind the triangles
    size_t geom_bytes = 0;        // d_tri32 (= the base of the one allocation) .. end of the nodes: what the traversal reads,
                                  // kept L2-resident by an access-policy window while the wavefront streams rays through
    double* d_tri64 = nullptr;
    float4* d_tri32 = nullptr;
    int32_t* d_tri_color = nullptr;
    int32_t* d_tri_emis = nullptr;
    double mesh_build_ms = 0.0;
    int mesh_nodes = 0;
    // wavefront buffers (mesh scenes), grown on demand
    void* wf_mem = nullptr;       size_t wf_cap = 0;
    unsigned long long* d_task_counter = nullptr;
    unsigned long long* d_retrace = nullptr;  size_t retrace_cap = 0;     // DRTB_MIXED: (pixel, sample) of the paths to re-trace
    unsigned int* d_retrace_count = nullptr;
    size_t partial_extra_rows = 0;            // gradient partial rows the re-trace will append behind the float pass's
    double* img_peers[drtb::kMaxPeers] = {};   // drtb_set_image_peers: full images the render kernel fills directly
    int n_img_peers = 0;
    double* grad_peers[drtb::kMaxPeers] = {};  // drtb_set_grad_peers: every rank's gradient exchange buffer
    int n_grad_peers = 0, grad_rank = 0;
    unsigned long long grad_epoch = 0;         // calls of the exchange so far (all ranks count alike)
    size_t l2_persist_max = 0;    // cudaDevAttrMaxPersistingL2CacheSize (0: not supported)
    bool l2_reserved = false;     // the persisting set-aside is currently reserved for this context's mesh
    size_t l2_window_max = 0;     // cudaDevAttrMaxAccessPolicyWindowSize
    // Preparation pass (drtb_reserve, and drtb_render before it starts its timer): every scratch buffer is
    // sized, every kernel attribute set and every kernel instantiation launched once on zero work (module load,
    // local-memory reservation), so that none of it lands between the events that drtb_stats.kernel_ms reports.
    bool dry = false;
    std::vector<const void*> warmed;     // kernel instantiations this context has launched before
};

namespace drtbh {

// NVTX range for the lifetime of the object: forward / adjoint render, gradient reduction, the cross-GPU exchange,
// the BVH build and the wavefront's batches show up as named spans in a timeline.
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range&) = delete;
    Range& operator=(const Range&) = delete;
};

int fail(drtb_ctx* ctx, int code, const std::string& msg);       // records the message, returns code

#define CK(ctx, call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return drtbh::fail(ctx, DRTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
int ensure(drtb_ctx* ctx, T*& p, size_t& cap, size_t n)
{
    if (n <= cap && p) return DRTB_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, DRTB_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cap = n;
    return DRTB_OK;
}

// true the first time `fn` (a kernel instantiation) is seen by this context
inline bool first_use(drtb_ctx* ctx, const void* fn)
{
    for (const void* f : ctx->warmed) if (f == fn) return false;
    ctx->warmed.push_back(fn);
    return true;
}

inline int shard_rows_impl(int H, int idx, int cnt, int band)
{
    if (H <= 0) return 0;
    if (cnt <= 1) return H;
    if (band < 1) band = 1;
    int rows = 0;
    const int nb = (H + band - 1) / band;
    for (int b = idx; b < nb; b += cnt) rows += std::min(band, H - b * band);
    return rows;
}

inline int effective_max_depth(const drtb_render_opts* o)
{
    if (o->max_depth > 0) return o->max_depth;
    return o->absorb == 1.0 ? std::max(1, o->min_bounces) : drtb::kMaxDepth;
}

inline drtb::MeshView mesh_view(const drtb_ctx* ctx)
{
    drtb::MeshView m{};
    m.nodes = ctx->d_nodes; m.tri64 = ctx->d_tri64; m.tri32 = ctx->d_tri32;
    m.color = ctx->d_tri_color; m.emis = ctx->d_tri_emis;
    m.n_tris = int32_t(ctx->n_tris); m.n_prims = int32_t(ctx->prims.size());
    return m;
}

inline void free_mesh(drtb_ctx* ctx)
{
    cudaFree(ctx->d_tri64); cudaFree(ctx->d_tri32);           // d_nodes points into d_tri32's allocation
    if (ctx->geom_bytes > 0 && ctx->l2_reserved) {            // give the L2 set-aside back to everybody else
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        ctx->l2_reserved = false;
    }
    ctx->geom_bytes = 0;
    cudaFree(ctx->d_tri_color); cudaFree(ctx->d_tri_emis);
    ctx->d_nodes = nullptr; ctx->d_tri64 = nullptr; ctx->d_tri32 = nullptr;
    ctx->d_tri_color = nullptr; ctx->d_tri_emis = nullptr;
    ctx->n_tris = 0;
}

// Optional per-pixel gradient image of one parameter (drtb_render_grad_image).
struct GradImage {
    int32_t param = -1;
    double* d_out = nullptr;             // shard_rows x W x 3 (device)
};

// How a render's units of work are cut into the chunks that warps claim from the global counter
// (render_kernel: units = warp tasks; render_regen_kernel: units = pixels).  The first n_big chunks hold
// `big` units each, the rest `small` units each (the last one possibly fewer): big chunks keep the claim and
// the per-chunk gradient row cheap, the small ones of the last round keep the tail of the kernel short.
// Pure host arithmetic, exported as drtb_chunk_plan so that the CPU tests can check that every unit is
// covered exactly once for any size.
constexpr int kRegenPixels = 64;         // pixels per chunk of the regenerating kernel (its shared accumulators)
struct ChunkPlan { long long big, small, n_big, n_chunks; };
ChunkPlan plan_chunks(long long n_units, int spp, long long resident_warps, bool regen, long long forced_big);

// grad[j] = sum over `rows` partial rows, in an order fixed by `rows` alone.  Up to 4096 rows:
// one block.  More (a large render leaves one row per chunk of warp tasks): a first pass of
// 1024-row blocks into the scratch rows behind the partials, then one block over those.
constexpr int kReduceDirectRows = 4096, kReduceBlockRows = 1024;
inline size_t reduce_scratch_rows(size_t rows) { return rows > kReduceDirectRows ? (rows + kReduceBlockRows - 1) / kReduceBlockRows : 0; }
int reduce_partials(drtb_ctx* ctx, double* partial, size_t rows, int P3, double* d_grad, cudaStream_t stream);   // drtb.cu

// ---- what the kernel translation units export ---------------------------------------------------
// Which instantiation of the analytic-scene kernels one render needs (chosen in drtb.cu).
struct AnalyticLaunch {
    bool smallp;            // <= kSmallP parameters: per-thread shared gradient columns
    int  queue;             // 0 no lit-path ring, 1 ring in shared memory, 2 ring in global memory
    bool gen;               // SpecularBxDF materials and / or a gradient image
    bool regen;             // Russian roulette on an all-diffuse scene: render_regen_kernel
    bool mixed;             // DRTB_MIXED's float pass: close calls go on the re-trace list (float launcher only)
    size_t smem;            // dynamic shared memory of render_kernel
    long long n_tasks;      // warp tasks (render_kernel)
    long long npix;         // pixels of the shard (render_regen_kernel)
    int  P3;
    bool want_grad;
};
// Enqueue (or, with ctx->dry, prepare) the analytic-scene render; `rows` = gradient partial rows written.
int launch_analytic_f64(drtb_ctx* ctx, drtb::RenderArgs& a, const AnalyticLaunch& l, cudaStream_t stream, size_t& rows);   // render_f64.cu
int launch_analytic_f32(drtb_ctx* ctx, drtb::RenderArgs& a, const AnalyticLaunch& l, cudaStream_t stream, size_t& rows);   // render_f32.cu
// DRTB_MIXED's second pass (render_f64.cu): re-traces the listed paths in double; its gradient rows follow row0.
int launch_retrace_f64(drtb_ctx* ctx, drtb::RenderArgs& a, int P3, bool want_grad, size_t row0, cudaStream_t stream, size_t& rows_added);
// Mesh scenes: the wavefront, batch by batch (mesh.cu).
int launch_wavefront(drtb_ctx* ctx, const drtb_render_opts* o, const double* d_seed, double* d_img, double* d_grad,
                     drtb_stats* d_stats, const GradImage& gi, cudaStream_t stream);
int mesh_upload(drtb_ctx* ctx, const drtb_mesh* mesh);             // mesh.cu: copy + GPU BVH build
// Every translation unit owns a copy of the device tables of real.cuh (no relocatable device code):
cudaError_t init_tables_render_f64();
cudaError_t init_tables_mesh();

} // namespace drtbh
