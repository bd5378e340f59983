// mesh.cu — config 4 of BASELINE.json: triangle meshes.  The GPU BVH build (bvh.cuh), the wavefront
// integrator's stage kernels (wavefront.cuh) in both precisions, its adjoint stage and the host loop
// that drives a render batch by batch.  NEW functionality: the reference scans a std::vector<Shape*>
// linearly and has no triangle (pathtracer.hpp:72-89).
#define DRTB_BVH_BUILD_KERNELS
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "host.hpp"
#include "sinks.cuh"
#include "wavefront.cuh"

namespace drtb {
namespace {

using drtbh::fail;
using drtbh::ensure;
using drtbh::free_mesh;
using drtbh::mesh_view;
using drtbh::shard_rows_impl;
using drtbh::effective_max_depth;
using drtbh::GradImage;

// DRTB_FLAG_DETERMINISTIC: the gradient buffer held 64-bit fixed-point sums (AtomicSink); back to doubles, in place
__global__ void fixed_to_double_kernel(double* __restrict__ grad, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) grad[i] = double(reinterpret_cast<const long long*>(grad)[i]) * (1.0 / kFixedScale);
}

__global__ void iota_kernel(int* __restrict__ v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// Wavefront stage 4: radiance recurrence + adjoint over the records one batch
// left in HBM, the per-pixel sums of src/render.cpp:78-82 and the gradient sums.
//
// Only the paths that reached a light carry anything (10 % of them in config 4), so a warp that swept its own 32
// slots kept 3 lanes busy.  A warp therefore owns a CHUNK of consecutive pixels (<= kAdjPixels, ~kAdjSlots path slots),
// reads the state words of the chunk 32 at a time, queues the lit slots in a small shared-memory ring and sweeps
// 32 queued records at once.  The ring is in slot order, so a pixel's samples are summed one after the other in
// sample order (add_to_pixels: lanes holding the same pixel add in lane order), which fixes the order of every
// floating-point sum whatever the chunk or the launch geometry.
constexpr int kAdjPixels = 32;
#ifndef DRTB_ADJ_SLOTS
#define DRTB_ADJ_SLOTS 512
#endif
constexpr int kAdjSlots = DRTB_ADJ_SLOTS;       // a 4 M-path batch is 8192 chunks for the 4736 warps of the launch
constexpr int kAdjRing = 64;
template <typename R, bool SMALLP, int CAP>
__global__ void __launch_bounds__(kBlock)
wf_adjoint(const __grid_constant__ DevScene<R> sc, const __grid_constant__ WfArgs a, const WfBuffers<R> b, int partial_row0)
{
    extern __shared__ double s_dyn[];
    __shared__ BlockScene<R> bs;
    __shared__ double s_red[kSmallP * 3][kWarpsPerBlock];
    __shared__ double s_px[kWarpsPerBlock][2][kAdjPixels * 3];       // radiance and gradient-image sums of the chunk's pixels
    __shared__ int s_ring[kWarpsPerBlock][kAdjRing];                  // queued lit slots: slot within the chunk | n << 20
    const bool want_grad = (a.flags & DRTB_FLAG_GRAD) != 0;
    const int P3 = sc.n_params * 3;
    double* s_acc = s_dyn;
    const int acc_doubles = (SMALLP && want_grad) ? P3 * kBlock : 0;
    load_block_scene(bs, sc, a.params);
    for (int i = threadIdx.x; i < acc_doubles; i += kBlock) s_acc[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int spp = a.spp;
    const long long npix = a.n_paths / spp, pix0 = a.first_path / spp;
    const int C = max(1, min(kAdjPixels, kAdjSlots / spp));           // pixels per chunk
    const long long n_tasks = (npix + C - 1) / C;
    const long long n_warps = (long long)gridDim.x * kWarpsPerBlock;
    const R inv_p = a.absorb < 1.0 ? R(1.0 / (1.0 - a.absorb)) : R(0);
    SmemSink ssink{s_acc + threadIdx.x};
    AtomicSink asink{a.grad_atomic, (a.flags & DRTB_FLAG_DETERMINISTIC) != 0};
    Materials<R, true> mat;
    mat.bs = &bs; mat.mesh = a.mesh; mat.params = a.params;
    double* pxacc = s_px[warp][0];
    double* pxg = s_px[warp][1];
    int* ring = s_ring[warp];
    uint32_t n_lit = 0;
    for (long long task = (long long)blockIdx.x * kWarpsPerBlock + warp; task < n_tasks; task += n_warps) {
        const long long cpix0 = task * C;                     // first pixel of the chunk, within the batch
        const int K = int(min((long long)C, npix - cpix0));
        const long long slot0 = cpix0 * spp;
        const int n_slots = K * spp;
        for (int k = lane; k < K * 3; k += 32) { pxacc[k] = 0.0; pxg[k] = 0.0; }
        __syncwarp();
        int q_head = 0, q_count = 0;                          // warp-uniform
        // lanes holding the same pixel add one after the other, in lane order (px < 0: nothing to add)
        auto add_to_pixels = [&](int px, const R* L0, const double* g) {
            const unsigned grp = __match_any_sync(0xffffffffu, px);
            const int rank = __popc(grp & lt_mask);
            const int rounds = __reduce_max_sync(0xffffffffu, px < 0 ? 0 : __popc(grp));
            for (int r = 0; r < rounds; ++r) {
                if (px >= 0 && rank == r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) { pxacc[3 * px + c] += double(L0[c]); pxg[3 * px + c] += g[c]; }
                }
                __syncwarp();
            }
        };
        auto drain = [&](int m) {
            __syncwarp();
            R L0[3] = {R(0), R(0), R(0)};
            double gacc[3] = {0.0, 0.0, 0.0};                 // gradient image (gimg_param == -1: stays zero)
            int px = -1;
            if (lane < m) {
                const int e = ring[(q_head + lane) & (kAdjRing - 1)];
                const int s = e & 0xfffff, n = e >> 20;
                px = s / spp;
                const long long p = slot0 + s;
                R g0[3] = {R(0), R(0), R(0)};
                if (want_grad) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) g0[c] = R(a.seed_scale * (a.seed_img ? a.seed_img[(pix0 + cpix0 + px) * 3 + c] : 1.0));
                }
                PixelSink<SmemSink> ps{ssink, a.gimg_param, gacc};
                PixelSink<AtomicSink> pa{asink, a.gimg_param, gacc};
                const WfRecordView<R, CAP> rec{b.rec_w + p, b.rec_prim + p, a.batch};
                if (SMALLP) radiance_and_adjoint(mat, rec, n, a.min_bounces, inv_p, want_grad, g0, L0, ps);
                else        radiance_and_adjoint(mat, rec, n, a.min_bounces, inv_p, want_grad, g0, L0, pa);
                n_lit += (L0[0] != R(0)) | (L0[1] != R(0)) | (L0[2] != R(0));
            }
            add_to_pixels(px, L0, gacc);
            q_head = (q_head + m) & (kAdjRing - 1);
            q_count -= m;
        };
        for (int s0 = 0; s0 < n_slots; s0 += 32) {
            const int s = s0 + lane;
            uint32_t st = 0;
            if (s < n_slots) st = b.state[slot0 + s];
            const bool lit = (st & kStLit) != 0;
            const unsigned m = __ballot_sync(0xffffffffu, lit);
            if (lit) ring[(q_head + q_count + __popc(m & lt_mask)) & (kAdjRing - 1)] = s | int((st >> 16) & 0xffu) << 20;
            q_count += __popc(m);
            if (q_count >= 32) drain(32);
        }
        if (q_count > 0) drain(q_count);
        __syncwarp();
        for (int k = lane; k < K * 3; k += 32) {
            const long long at = (pix0 + cpix0) * 3 + k;      // pixel within the shard
            if (a.img) a.img[at] = pxacc[k] / double(spp);
            if (a.gimg) a.gimg[at] = pxg[k];
        }
        __syncwarp();
    }
    if (SMALLP && want_grad) {
        for (int j = 0; j < P3; ++j) {
            double v = warp_sum(s_acc[j * kBlock + threadIdx.x]);
            if (lane == 0) s_red[j][warp] = v;
        }
        __syncthreads();
        if (threadIdx.x < P3) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kWarpsPerBlock; ++w) v += s_red[threadIdx.x][w];
            a.grad_partial[((size_t)partial_row0 + blockIdx.x) * P3 + threadIdx.x] = v;
        }
    }
    if (a.stats) {
        unsigned long long t = n_lit;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && t) atomicAdd((unsigned long long*)&a.stats->lit_paths, t);
    }
}

// Mesh scenes: the wavefront of wavefront.cuh, batch by batch.
template <typename R>
int launch_wavefront(drtb_ctx* ctx, const DevScene<R>& sc, const drtb_render_opts* o, const double* d_seed, double* d_img,
                     double* d_grad, drtb_stats* d_stats, const GradImage& gi, cudaStream_t stream)
{
    const int W = ctx->camera.width, H = ctx->camera.height;
    const int P = int(ctx->params.size() / 3), P3 = P * 3;
    const bool want_grad = (o->flags & DRTB_FLAG_GRAD) != 0, want_img = (o->flags & DRTB_FLAG_IMAGE) != 0;
    const int cnt = o->shard_count > 1 ? o->shard_count : 1;
    const int rows = shard_rows_impl(H, o->shard_index, cnt, o->band_rows);
    const long long npix = (long long)rows * W;
    const int D = effective_max_depth(o);
    const bool smallp = P <= kSmallP;
    // whole pixels per batch, a multiple of 32 so that warps of wf_adjoint never straddle batches
    long long pix_per_batch = std::max<long long>(32, (kBatchPaths / o->spp) / 32 * 32);
    pix_per_batch = std::min<long long>(pix_per_batch, (npix + 31) / 32 * 32);
    const long long batch = pix_per_batch * o->spp;
    if (batch > (1ll << 30)) return fail(ctx, DRTB_ERR_UNSUPPORTED, "spp too large for one wavefront batch");
    const int n_batches = int((npix + pix_per_batch - 1) / pix_per_batch);

    // carve the buffers out of one allocation
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t sz_ray = up(size_t(batch) * sizeof(R4<R>)), sz_i = up(size_t(batch) * 4);
    const size_t sz_rw = up(size_t(batch) * D * sizeof(R)), sz_rp = up(size_t(batch) * D * 4), sz_cnt = up(size_t(D + 2) * 4);
    const size_t total = 2 * sz_ray + 2 * sz_i + sz_rw + sz_rp + 2 * sz_cnt;
    if (total > ctx->wf_cap) {
        cudaFree(ctx->wf_mem); ctx->wf_mem = nullptr; ctx->wf_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->wf_mem, total);
        if (e != cudaSuccess) return fail(ctx, DRTB_ERR_NOMEM, std::string("cudaMalloc (wavefront buffers): ") + cudaGetErrorString(e));
        ctx->wf_cap = total;
    }
    char* mem = static_cast<char*>(ctx->wf_mem);
    WfBuffers<R> b{};
    b.ray_a = reinterpret_cast<R4<R>*>(mem); mem += sz_ray;
    b.ray_b = reinterpret_cast<R4<R>*>(mem); mem += sz_ray;
    b.hit = reinterpret_cast<int32_t*>(mem); mem += sz_i;
    b.state = reinterpret_cast<uint32_t*>(mem); mem += sz_i;
    b.rec_w = reinterpret_cast<R*>(mem); mem += sz_rw;
    b.rec_prim = reinterpret_cast<int32_t*>(mem); mem += sz_rp;
    b.alive_count = reinterpret_cast<int32_t*>(mem); mem += sz_cnt;
    b.fetch = reinterpret_cast<uint32_t*>(mem);

    WfArgs a{};
    a.spp = o->spp; a.min_bounces = o->min_bounces; a.max_depth = D; a.flags = o->flags; a.absorb = o->absorb;
    a.key0 = o->seed * kSeedMul;
    a.shard_index = o->shard_index; a.shard_count = cnt; a.band_rows = o->band_rows > 0 ? o->band_rows : 1;
    a.batch = int(batch); a.seed_scale = o->seed_scale;
    a.params = ctx->d_params; a.seed_img = d_seed; a.img = want_img ? d_img : nullptr;
    a.stats = (o->flags & DRTB_FLAG_STATS) ? d_stats : nullptr;
    a.mesh = mesh_view(ctx);
    a.gimg = (want_grad && gi.d_out) ? gi.d_out : nullptr;
    a.gimg_param = a.gimg ? gi.param : -1;
    a.specular = ctx->has_specular ? 1 : 0;
    if (a.stats && !ctx->dry) CK(ctx, cudaMemsetAsync(d_stats, 0, sizeof(drtb_stats), stream));
    if (want_grad && !smallp) {
        if (!ctx->dry) CK(ctx, cudaMemsetAsync(d_grad, 0, sizeof(double) * P3, stream));
        a.grad_atomic = d_grad;
    }
    // grids
    int trav_per_sm = 0;
    CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&trav_per_sm, wf_traverse<R>, 128, 0));
    const int trav_grid = ctx->sm_count * std::max(1, trav_per_sm);
    const size_t adj_smem = (smallp && want_grad) ? size_t(P3) * kBlock * sizeof(double) : 0;
    const int adj_grid = ctx->sm_count * 8;
    if (want_grad && smallp) {
        const size_t rows = size_t(adj_grid) * n_batches;
        int rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (rows + drtbh::reduce_scratch_rows(rows)) * P3);
        if (rc != DRTB_OK) return rc;
        a.grad_partial = ctx->d_partial;
    }
    const bool no_bvh = (o->flags & DRTB_FLAG_NO_BVH) != 0;
    const bool deep = D > kQueueDepth;
    if (smallp) {
        if (deep) CK(ctx, cudaFuncSetAttribute(wf_adjoint<R, true, kMaxDepth>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(adj_smem)));
        else      CK(ctx, cudaFuncSetAttribute(wf_adjoint<R, true, kQueueDepth>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(adj_smem)));
    }
    if (ctx->dry) {
        // buffers are sized; load the stage kernels (lazy module loading) without running them
        cudaFuncAttributes fa;
        CK(ctx, cudaFuncGetAttributes(&fa, wf_generate<R>));
        CK(ctx, cudaFuncGetAttributes(&fa, wf_traverse<R>));
        CK(ctx, cudaFuncGetAttributes(&fa, wf_shade<R>));
        return DRTB_OK;
    }
    // "Scene geometry and the BVH staged in L2" (north_star): the rays of a batch stream through L2 once per depth
    // (4 M x 72 B), which evicts the 85 MB of triangles and nodes that EVERY ray reads.  A persisting access-policy
    // window over the geometry keeps it resident; ray traffic is left to the normal (evict-first for misses) policy.
    const bool l2_window = ctx->l2_persist_max > 0 && ctx->geom_bytes > 0 && !no_bvh && std::getenv("DRTB_NO_L2_WINDOW") == nullptr;
    if (l2_window && !ctx->l2_reserved) {                    // reserved while the mesh is attached (free_mesh gives it back)
        const size_t want = std::min(ctx->l2_persist_max, (ctx->geom_bytes + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1));
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) ctx->l2_reserved = true;
        else cudaGetLastError();
    }
    if (l2_window && ctx->l2_reserved) {
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = ctx->d_tri32;
        av.accessPolicyWindow.num_bytes = std::min(ctx->geom_bytes, ctx->l2_window_max);
        av.accessPolicyWindow.hitRatio = float(std::min(1.0, double(ctx->l2_persist_max) / double(av.accessPolicyWindow.num_bytes)));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(ctx, cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    for (int bi = 0; bi < n_batches; ++bi) {
        drtbh::Range batch_range("drtb: wavefront batch (generate, traverse + shade per depth, adjoint)");
        const long long p0 = (long long)bi * pix_per_batch;
        a.first_path = p0 * o->spp;
        a.n_paths = int(std::min<long long>(pix_per_batch, npix - p0) * o->spp);
        const int g256 = (a.n_paths + 255) / 256;
        CK(ctx, cudaMemsetAsync(b.alive_count, 0, 2 * sz_cnt, stream));          // alive_count and fetch
        a.depth = 0;
        wf_generate<R><<<g256, 256, 0, stream>>>(sc, a, b);
        for (int depth = 0; depth < D; ++depth) {
            a.depth = depth;
            if (no_bvh) wf_traverse_brute<R><<<(a.n_paths + 127) / 128, 128, 0, stream>>>(a, b);
            else        wf_traverse<R><<<trav_grid, 128, 0, stream>>>(a, b);
            wf_shade<R><<<g256, 256, 0, stream>>>(sc, a, b);
        }
        if (smallp) {
            if (deep) wf_adjoint<R, true, kMaxDepth><<<adj_grid, kBlock, adj_smem, stream>>>(sc, a, b, bi * adj_grid);
            else      wf_adjoint<R, true, kQueueDepth><<<adj_grid, kBlock, adj_smem, stream>>>(sc, a, b, bi * adj_grid);
        } else {
            if (deep) wf_adjoint<R, false, kMaxDepth><<<adj_grid, kBlock, 0, stream>>>(sc, a, b, 0);
            else      wf_adjoint<R, false, kQueueDepth><<<adj_grid, kBlock, 0, stream>>>(sc, a, b, 0);
        }
        CK(ctx, cudaGetLastError());
        ctx->launches += 2 + 2 * D;
    }
    if (l2_window && ctx->l2_reserved) {
        cudaStreamAttrValue av{};                            // num_bytes = 0: window off for whatever the caller enqueues next
        CK(ctx, cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    if (want_grad && !smallp && (o->flags & DRTB_FLAG_DETERMINISTIC) && !ctx->dry) {
        fixed_to_double_kernel<<<(P3 + 255) / 256, 256, 0, stream>>>(d_grad, P3);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
    }
    if (want_grad && smallp) {
        // at most 148 * 8 * n_batches rows: the fixed-order reduction of drtb.cu (one block up to 4096 rows)
        const size_t rows = size_t(adj_grid) * n_batches;
        const int rc = drtbh::reduce_partials(ctx, ctx->d_partial, rows, P3, d_grad, stream);
        if (rc != DRTB_OK) return rc;
    }
    return DRTB_OK;
}

int mesh_upload_impl(drtb_ctx* ctx, const drtb_mesh* mesh)
{
    drtbh::Range whole("drtb_mesh_upload: copy + GPU BVH build");
    if (!ctx->has_scene) return fail(ctx, DRTB_ERR_INVALID, "upload a scene before attaching a mesh");
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    free_mesh(ctx);
    if (!mesh || mesh->n_triangles == 0) return DRTB_OK;
    const int64_t n = mesh->n_triangles, nv = mesh->n_vertices;
    if (n < 0 || nv <= 0 || !mesh->vertices || !mesh->indices) return fail(ctx, DRTB_ERR_INVALID, "mesh has NULL arrays or bad counts");
    if (n > (int64_t(1) << 28)) return fail(ctx, DRTB_ERR_UNSUPPORTED, "more than 2^28 triangles");
    const int P = int(ctx->params.size() / 3);
    for (int64_t i = 0; i < 3 * n; ++i)
        if (mesh->indices[i] < 0 || mesh->indices[i] >= nv) return fail(ctx, DRTB_ERR_INVALID, "mesh vertex index out of range");
    for (int64_t i = 0; i < n; ++i) {
        if (mesh->color && (mesh->color[i] < -1 || mesh->color[i] >= P)) return fail(ctx, DRTB_ERR_INVALID, "triangle colour parameter index out of range");
        if (mesh->emission && (mesh->emission[i] < -1 || mesh->emission[i] >= P)) return fail(ctx, DRTB_ERR_INVALID, "triangle emission parameter index out of range");
    }
    // ---- device buffers: persistent mesh data + build temporaries
    const bool use_lbvh = [] { const char* e = std::getenv("DRTB_BVH"); return e && std::string(e) == "lbvh"; }();
    double* d_vert = nullptr; int32_t* d_idx = nullptr;
    float4 *d_lo = nullptr, *d_hi = nullptr, *d_blo = nullptr, *d_bhi = nullptr; uint4* d_wide = nullptr; uint32_t* d_bounds = nullptr;
    uint64_t *d_keys = nullptr, *d_keys2 = nullptr, *d_flags = nullptr, *d_scan = nullptr;
    uint32_t *d_vals = nullptr, *d_vals2 = nullptr;
    int2 *d_children = nullptr, *d_tasks = nullptr, *d_tasks2 = nullptr;
    int *d_parent = nullptr, *d_arrive = nullptr, *d_clusters = nullptr, *d_clusters2 = nullptr, *d_nearest = nullptr;
    int32_t* d_leaf_order = nullptr; CollapseCounters* d_cnt = nullptr; void *d_tmp = nullptr, *d_tmp2 = nullptr;
    BuildState* d_state = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_vert); cudaFree(d_idx); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_blo); cudaFree(d_bhi); cudaFree(d_wide);
        cudaFree(d_bounds); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_flags); cudaFree(d_scan); cudaFree(d_vals);
        cudaFree(d_vals2); cudaFree(d_children); cudaFree(d_tasks); cudaFree(d_tasks2); cudaFree(d_parent); cudaFree(d_arrive);
        cudaFree(d_clusters); cudaFree(d_clusters2); cudaFree(d_nearest); cudaFree(d_leaf_order); cudaFree(d_cnt);
        cudaFree(d_tmp); cudaFree(d_tmp2); cudaFree(d_state);
    };
#define CKM(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); free_mesh(ctx); return fail(ctx, e_ == cudaErrorMemoryAllocation ? DRTB_ERR_NOMEM : DRTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    cudaStream_t st = ctx->stream;
    const size_t nn = size_t(n), n_int = nn > 1 ? nn - 1 : 1;
    CKM(cudaMalloc((void**)&ctx->d_tri64, nn * kTri64Stride * sizeof(double)));
    CKM(cudaMalloc((void**)&ctx->d_tri_color, nn * sizeof(int32_t)));
    CKM(cudaMalloc((void**)&ctx->d_tri_emis, nn * sizeof(int32_t)));
    CKM(cudaMalloc((void**)&d_vert, size_t(nv) * 3 * sizeof(double)));
    CKM(cudaMalloc((void**)&d_idx, nn * 3 * sizeof(int32_t)));
    CKM(cudaMalloc((void**)&d_lo, nn * sizeof(float4)));      CKM(cudaMalloc((void**)&d_hi, nn * sizeof(float4)));
    CKM(cudaMalloc((void**)&d_blo, 2 * nn * sizeof(float4))); CKM(cudaMalloc((void**)&d_bhi, 2 * nn * sizeof(float4)));
    CKM(cudaMalloc((void**)&d_wide, nn * kNodeStride * sizeof(uint4)));       // a wide node has >= 2 children: < n nodes
    CKM(cudaMalloc((void**)&d_bounds, 6 * sizeof(uint32_t)));
    CKM(cudaMalloc((void**)&d_keys, nn * sizeof(uint64_t)));  CKM(cudaMalloc((void**)&d_keys2, nn * sizeof(uint64_t)));
    CKM(cudaMalloc((void**)&d_vals, nn * sizeof(uint32_t)));  CKM(cudaMalloc((void**)&d_vals2, nn * sizeof(uint32_t)));
    CKM(cudaMalloc((void**)&d_children, n_int * sizeof(int2)));
    CKM(cudaMalloc((void**)&d_tasks, nn * sizeof(int2)));     CKM(cudaMalloc((void**)&d_tasks2, nn * sizeof(int2)));
    CKM(cudaMalloc((void**)&d_leaf_order, nn * sizeof(int32_t)));
    CKM(cudaMalloc((void**)&d_cnt, sizeof(CollapseCounters)));
    CKM(cudaMalloc((void**)&d_state, sizeof(BuildState)));
    const BuildState init_state{int(n), 0, 1, 0};
    CKM(cudaMemcpyAsync(d_state, &init_state, sizeof init_state, cudaMemcpyHostToDevice, ctx->stream));
    CKM(cudaMemcpyAsync(d_vert, mesh->vertices, size_t(nv) * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CKM(cudaMemcpyAsync(d_idx, mesh->indices, nn * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (mesh->color) CKM(cudaMemcpyAsync(ctx->d_tri_color, mesh->color, nn * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    else CKM(cudaMemsetAsync(ctx->d_tri_color, 0xff, nn * sizeof(int32_t), st));
    if (mesh->emission) CKM(cudaMemcpyAsync(ctx->d_tri_emis, mesh->emission, nn * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    else CKM(cudaMemsetAsync(ctx->d_tri_emis, 0xff, nn * sizeof(int32_t), st));
    // scene bounds start at (+max, -max) in the ordered-uint encoding
    const uint32_t init_bounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CKM(cudaMemcpyAsync(d_bounds, init_bounds, sizeof init_bounds, cudaMemcpyHostToDevice, st));
    CKM(cudaEventRecord(ctx->ev0, st));
    const int T = 256, G = int((nn + T - 1) / T);
    const BinTree bt{d_blo, d_bhi, d_children};
    // 1. bounds, Morton codes, sort
    mesh_prepare_kernel<<<G, T, 0, st>>>(d_vert, d_idx, int(n), ctx->d_tri64, d_lo, d_hi, d_bounds);
    CKM(cudaGetLastError());
    mesh_morton_kernel<<<G, T, 0, st>>>(d_lo, d_hi, d_bounds, int(n), d_keys, d_vals);
    CKM(cudaGetLastError());
    size_t tmp_bytes = 0;
    CKM(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, int(n), 0, 63, st));
    CKM(cudaMalloc(&d_tmp, tmp_bytes));
    CKM(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, int(n), 0, 63, st));
    bin_leaves_kernel<<<G, T, 0, st>>>(d_vals2, d_lo, d_hi, d_bounds, int(n), bt);
    CKM(cudaGetLastError());
    ctx->launches += 5;                                      // prepare, morton, sort (>= 2), leaves
    // 2. binary tree
    int root = 0;
    if (n > 1 && use_lbvh) {
        CKM(cudaMalloc((void**)&d_parent, 2 * nn * sizeof(int)));
        CKM(cudaMalloc((void**)&d_arrive, n_int * sizeof(int)));
        CKM(cudaMemsetAsync(d_arrive, 0, n_int * sizeof(int), st));
        lbvh_hierarchy_kernel<<<G, T, 0, st>>>(d_keys2, int(n), d_children, d_parent);
        CKM(cudaGetLastError());
        lbvh_refit_kernel<<<G, T, 0, st>>>(int(n), d_parent, d_arrive, bt);
        CKM(cudaGetLastError());
        ctx->launches += 2;
        root = int(n);                                       // Karras: internal node 0 is the root
    } else if (n > 1) {
        CKM(cudaMalloc((void**)&d_clusters, nn * sizeof(int)));  CKM(cudaMalloc((void**)&d_clusters2, nn * sizeof(int)));
        CKM(cudaMalloc((void**)&d_nearest, nn * sizeof(int)));
        CKM(cudaMalloc((void**)&d_flags, (nn + 1) * sizeof(uint64_t)));
        CKM(cudaMalloc((void**)&d_scan, (nn + 1) * sizeof(uint64_t)));
        size_t scan_bytes = 0;
        CKM(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_flags, d_scan, int(n) + 1, st));
        CKM(cudaMalloc(&d_tmp2, scan_bytes));
        iota_kernel<<<G, T, 0, st>>>(d_clusters, int(n));
        CKM(cudaGetLastError());
        CKM(cudaMemsetAsync(d_flags, 0, (nn + 1) * sizeof(uint64_t), st));
        // device-paced (bvh.cuh, BuildState): batches of kPlocBatch iterations between two reads of the state
        constexpr int kPlocBatch = 8;
        const int g1 = int((nn + 1 + T - 1) / T);
        int m = int(n), made = 0;
        for (long long it = 0; m > 1; ) {
            for (int k = 0; k < kPlocBatch; ++k, ++it) {
                int* src = (it & 1) ? d_clusters2 : d_clusters;
                int* dst = (it & 1) ? d_clusters : d_clusters2;
                ploc_nearest_kernel<<<G, 256, 0, st>>>(src, d_state, bt, d_nearest);
                ploc_flag_kernel<<<g1, T, 0, st>>>(d_nearest, d_state, int(n), d_flags);
                CKM(cub::DeviceScan::ExclusiveSum(d_tmp2, scan_bytes, d_flags, d_scan, int(n) + 1, st));   // scan[m] = totals
                ploc_merge_kernel<<<G, T, 0, st>>>(src, d_nearest, d_flags, d_scan, d_state, int(n), bt, dst);
                ploc_advance_kernel<<<1, 1, 0, st>>>(d_scan, d_state);
                CKM(cudaGetLastError());
                ctx->launches += 5;
            }
            BuildState h{};
            CKM(cudaMemcpyAsync(&h, d_state, sizeof h, cudaMemcpyDeviceToHost, st));
            CKM(cudaStreamSynchronize(st));
            if (h.error || (h.m >= m && h.m > 1)) { cleanup(); free_mesh(ctx); return fail(ctx, DRTB_ERR_CUDA, "PLOC iteration made no progress"); }
            m = h.m; made = h.made;
        }
        root = int(n) + made - 1;                            // the last node created
    }
    // 3. collapse to the 8-wide compressed BVH, one level per launch
    const CollapseCounters init_cnt{1, 0, 0, 0};
    const int2 root_task = make_int2(root, 0);
    CKM(cudaMemcpyAsync(d_cnt, &init_cnt, sizeof init_cnt, cudaMemcpyHostToDevice, st));
    CKM(cudaMemcpyAsync(d_tasks, &root_task, sizeof root_task, cudaMemcpyHostToDevice, st));
    // device-paced as well: a level's grid is sized for an upper bound of its tasks (8 x the level above, at most n),
    // the task count itself stays on the device; the host looks once per kCollapseBatch levels
    constexpr int kCollapseBatch = 4;
    CollapseCounters h_cnt = init_cnt;
    long long bound = 1;
    for (long long lvl = 0, n_tasks = 1; n_tasks > 0; ) {
        for (int k = 0; k < kCollapseBatch; ++k, ++lvl) {
            int2* src = (lvl & 1) ? d_tasks2 : d_tasks;
            int2* dst = (lvl & 1) ? d_tasks : d_tasks2;
            collapse8_kernel<<<int((bound + T - 1) / T), T, 0, st>>>(src, d_state, int(n), bt, d_vals2, d_wide, d_leaf_order, d_cnt, dst);
            collapse_advance_kernel<<<1, 1, 0, st>>>(d_cnt, d_state);
            CKM(cudaGetLastError());
            bound = std::min<long long>(bound * 8, (long long)nn);
            ctx->launches += 2;
        }
        BuildState h{};
        CKM(cudaMemcpyAsync(&h, d_state, sizeof h, cudaMemcpyDeviceToHost, st));
        CKM(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
        CKM(cudaStreamSynchronize(st));
        n_tasks = h.n_tasks;
    }
    if (h_cnt.tris != int(n)) { cleanup(); free_mesh(ctx); return fail(ctx, DRTB_ERR_CUDA, "BVH collapse lost triangles"); }
    // what the traversal reads -- float triangles (leaf order), then the wide nodes -- in ONE allocation, so that a
    // single L2 access-policy window can keep it resident (launch_wavefront)
    const size_t tri_bytes = (nn * kTri32Stride * sizeof(float4) + 255) & ~size_t(255);
    const size_t node_bytes = size_t(h_cnt.nodes) * kNodeStride * sizeof(uint4);
    CKM(cudaMalloc((void**)&ctx->d_tri32, tri_bytes + node_bytes));
    ctx->d_nodes = reinterpret_cast<float4*>(reinterpret_cast<char*>(ctx->d_tri32) + tri_bytes);
    ctx->geom_bytes = tri_bytes + node_bytes;
    CKM(cudaMemcpyAsync(ctx->d_nodes, d_wide, node_bytes, cudaMemcpyDeviceToDevice, st));
    leaf_triangles_kernel<<<G, T, 0, st>>>(d_leaf_order, ctx->d_tri64, int(n), ctx->d_tri32);
    CKM(cudaGetLastError());
    ctx->launches++;
    CKM(cudaEventRecord(ctx->ev1, st));
    CKM(cudaStreamSynchronize(st));
    float ms = 0.f;
    CKM(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->mesh_build_ms = ms;
    ctx->mesh_nodes = h_cnt.nodes;
#undef CKM
    cleanup();
    ctx->n_tris = n;
    return DRTB_OK;
}

} // namespace
} // namespace drtb

namespace drtbh {

int launch_wavefront(drtb_ctx* ctx, const drtb_render_opts* o, const double* d_seed, double* d_img, double* d_grad,
                     drtb_stats* d_stats, const GradImage& gi, cudaStream_t stream)
{
    return o->precision == DRTB_F32 ? drtb::launch_wavefront<float>(ctx, ctx->sc32, o, d_seed, d_img, d_grad, d_stats, gi, stream)
                                    : drtb::launch_wavefront<double>(ctx, ctx->sc64, o, d_seed, d_img, d_grad, d_stats, gi, stream);
}

int mesh_upload(drtb_ctx* ctx, const drtb_mesh* mesh) { return drtb::mesh_upload_impl(ctx, mesh); }

cudaError_t init_tables_mesh() { return drtb::upload_sincos_tab(); }

} // namespace drtbh
