// drtb.cu — the C ABI of include/drtb.h and the host side of a render: scene flattening, the choice of
// kernel instantiation, gradient reduction, explicit rays.  The render kernels themselves live in
// render_f64.cu / render_f32.cu (render_kernels.cuh) and mesh.cu (bvh.cuh, wavefront.cuh); host.hpp is
// what the translation units share.
//
// Kernels in this file
//   reduce_grad_kernel         fixed-order sum of the per-chunk gradient partials
//   trace_rays_kernel<R>       Pathtracer::trace on explicit rays (+ Jacobian)
//   fma_peak_kernel<R>         FMA issue-rate micro-benchmark (roofline denominator)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see build.py).
#include <new>

#include "host.hpp"
#include "sinks.cuh"

using namespace drtb;
using namespace drtbh;

// ===========================================================================
// device code
// ===========================================================================
namespace {

// out[blockIdx.x][j] = sum of partial[r][j] over this block's rows
// [blockIdx.x * rows_per_block, ...).  Thread t owns column t % P3 and every
// (256 / P3)-th row, so a warp reads consecutive doubles of the row-major
// partials; the row lanes of a column are then added in ascending order by one
// thread.  The order depends on (n_rows, rows_per_block, P3) alone.  One block
// over all rows gives grad[j]; many rows take two passes (reduce_partials).
// P3 <= 256 (kMaxParams * 3 = 192).
__global__ void __launch_bounds__(256)
reduce_grad_kernel(const double* __restrict__ partial, int n_rows, int rows_per_block, int P3, double* __restrict__ out)
{
    __shared__ double s[256];
    const int lanes = 256 / P3;                      // row lanes per column
    const int col = threadIdx.x % P3, rl = threadIdx.x / P3;
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(n_rows, r0 + rows_per_block);
    double v = 0.0;
    if (rl < lanes)
        for (int r = r0 + rl; r < r1; r += lanes) v += partial[(size_t)r * P3 + col];
    s[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x < P3) {
        double t = 0.0;
        for (int l = 0; l < lanes; ++l) t += s[l * P3 + threadIdx.x];
        out[(size_t)blockIdx.x * P3 + threadIdx.x] = t;
    }
}

// Sum of the gradients of the GPUs of one NVSwitch box WITHOUT a collective library: every rank stores its P3
// doubles into slot `rank` of the exchange buffer of EVERY rank (peer stores over NVLink), then a flag; waits until
// the flags of all ranks have arrived in its OWN buffer; and adds the slots in rank order -- the same bits on every
// rank.  One 1-block launch behind reduce_grad_kernel, where ncclAllReduce of these 96 bytes cost ~0.1 ms of
// latency per render at 8 GPUs.
//   exchange buffer of a rank: [2 parities][n][P3] doubles, then n 64-bit flags (zero-initialised)
// Slots alternate between two parities per call (`epoch`), so a fast rank's stores of call e + 1 never land on
// the slots a slow rank is still adding for call e; it cannot reach call e + 2 before the slow rank has set its
// flag for e + 1, i.e. finished e.  A rank that never arrives (its render failed) is given ~2 s, then the sums
// are poisoned with NaN instead of hanging the GPU.
struct GradPeers { double* buf[kMaxPeers]; };
__global__ void __launch_bounds__(256)
grad_allreduce_kernel(const __grid_constant__ GradPeers peers, int n, int rank, int P3, unsigned long long epoch, double* grad)
{
    __shared__ int s_timeout;
    if (threadIdx.x == 0) s_timeout = 0;
    const size_t parity = size_t(epoch & 1ull) * n * P3;
    for (int i = threadIdx.x; i < n * P3; i += blockDim.x) {
        const int peer = i / P3, j = i - peer * P3;
        peers.buf[peer][parity + size_t(rank) * P3 + j] = grad[j];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < n) {
        unsigned long long* theirs = reinterpret_cast<unsigned long long*>(peers.buf[threadIdx.x] + size_t(2) * n * P3);
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(theirs + rank), "l"(epoch) : "memory");
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers.buf[rank] + size_t(2) * n * P3) + threadIdx.x;
        const long long t0 = clock64();
        unsigned long long seen = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
            if (seen >= epoch) break;
            if (clock64() - t0 > 4000000000ll) { s_timeout = 1; break; }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < P3; j += blockDim.x) {
        double v = 0.0;
        for (int r = 0; r < n; ++r) v += peers.buf[rank][parity + size_t(r) * P3 + j];
        grad[j] = s_timeout ? __longlong_as_double(0x7ff8000000000000ll) : v;
    }
}

// Pathtracer<T>::trace(scene, orig, dir) for user-supplied rays (pathtracer.hpp:121-136).
template <typename R, bool MESH>
__global__ void __launch_bounds__(kBlock)
trace_rays_kernel(const __grid_constant__ DevScene<R> sc, const double* __restrict__ params, const MeshView mesh,
                  uint32_t flags, int min_bounces, double absorb, int max_depth, long long n,
                  const double* __restrict__ orig, const double* __restrict__ dir,
                  const uint64_t* __restrict__ keys, double* __restrict__ radiance, double* jac)
{
    __shared__ BlockScene<R> bs;
    load_block_scene(bs, sc, params);
    __syncthreads();
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    Materials<R, MESH> mat;
    mat.bs = &bs;
    if constexpr (MESH) { mat.mesh = mesh; mat.params = params; }
    const R inv_p = absorb < 1.0 ? R(1.0 / (1.0 - absorb)) : R(0);
    V3<R> o = {R(orig[3 * i]), R(orig[3 * i + 1]), R(orig[3 * i + 2])};
    V3<R> d = {R(dir[3 * i]), R(dir[3 * i + 1]), R(dir[3 * i + 2])};
    PathRecord<R, MESH, kMaxDepth> rec;
    bool lit;
    TraceCounters cnt;
    int nv = trace_path<R, MESH, kMaxDepth, true>(sc, bs, mat, (flags & DRTB_FLAG_NO_BVH) != 0, keys[i] * kKeyMul, 2u, o, d,
                                                  min_bounces, absorb, max_depth, rec, lit, cnt);
    R L0[3] = {R(0), R(0), R(0)};
    if (lit) {
        const R one[3] = {R(1), R(1), R(1)};
        JacSink sink{jac ? jac + (size_t)i * sc.n_params * 3 : nullptr};
        radiance_and_adjoint(mat, rec, nv, min_bounces, inv_p, jac != nullptr, one, L0, sink);
    }
    radiance[3 * i] = double(L0[0]); radiance[3 * i + 1] = double(L0[1]); radiance[3 * i + 2] = double(L0[2]);
}

// 8 independent FMA chains per thread, operands in registers: the non-tensor
// FMA pipe at its issue limit.
template <typename R>
__global__ void __launch_bounds__(256) fma_peak_kernel(R* out, int iters, R a, R b)
{
    R x0 = R(threadIdx.x), x1 = x0 + R(1), x2 = x0 + R(2), x3 = x0 + R(3);
    R x4 = x0 + R(4), x5 = x0 + R(5), x6 = x0 + R(6), x7 = x0 + R(7);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = Real<R>::fma(x0, a, b); x1 = Real<R>::fma(x1, a, b); x2 = Real<R>::fma(x2, a, b); x3 = Real<R>::fma(x3, a, b);
            x4 = Real<R>::fma(x4, a, b); x5 = Real<R>::fma(x5, a, b); x6 = Real<R>::fma(x6, a, b); x7 = Real<R>::fma(x7, a, b);
        }
    }
    R s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == R(-1)) out[0] = s;                    // never true; keeps the chains alive
}

} // namespace

// ===========================================================================
// host side
// ===========================================================================
namespace {

thread_local std::string g_create_err;
constexpr int kMixedMaxDepth = 8;         // DRTB_MIXED's float pass: deeper paths drift too far from the double path (path.cuh)
constexpr int kMaxExchangeP3 = 4096;      // gradient scalars the peer exchange (grad_allreduce_kernel) serves

template <typename R>
void fill_dev_scene(DevScene<R>& d, const drtb_ctx& c)
{
    memset(&d, 0, sizeof d);
    d.n_prims = int(c.prims.size());
    d.n_params = int(c.params.size() / 3);
    // scan slots (see DevScene): axis-aligned unit planes go to the per-axis windows (at most
    // kAxisFast each, the rest are scanned as general planes -- same t bit for bit); the first
    // kFast other planes / spheres sit right-aligned in the straight-line windows, the rest are
    // appended in scene order
    std::vector<int> planes, spheres, axis_planes[3];
    for (int i = 0; i < d.n_prims; ++i) {
        const drtb_prim& p = c.prims[i];
        if (p.type != DRTB_PLANE) { spheres.push_back(i); continue; }
        int axis = -1, nz = 0;
        for (int j = 0; j < 3; ++j) if (p.v[j] != 0.0) { ++nz; axis = j; }
        if (nz == 1 && std::fabs(p.v[axis]) == 1.0 && std::isfinite(p.v[3]) && int(axis_planes[axis].size()) < kAxisFast)
            axis_planes[axis].push_back(i);
        else
            planes.push_back(i);
    }
    d.n_fast_planes = std::min<int>(kFast, int(planes.size()));
    d.n_fast_spheres = std::min<int>(kFast, int(spheres.size()));
    d.n_over_planes = int(planes.size()) - d.n_fast_planes;
    d.n_over_spheres = int(spheres.size()) - d.n_fast_spheres;
    auto put = [&](int slot, int i) {
        for (int j = 0; j < 4; ++j) d.prim[slot][j] = R(c.prims[i].v[j]);
        for (int j = 0; j < 3; ++j) d.sph[slot][j] = R(c.prims[i].v[j]);
        d.sph[slot][3] = R(c.prims[i].v[3] * c.prims[i].v[3]);        // spheres: m_radius * m_radius, shape.hpp:85
        d.id[slot] = i;
        d.slot[i] = int8_t(slot);
    };
    for (int j = 0; j < int(planes.size()); ++j)
        put(j < d.n_fast_planes ? kFast - d.n_fast_planes + j : 2 * kFast + (j - d.n_fast_planes), planes[j]);
    for (int j = 0; j < int(spheres.size()); ++j)
        put(j < d.n_fast_spheres ? 2 * kFast - d.n_fast_spheres + j
                                 : 2 * kFast + d.n_over_planes + (j - d.n_fast_spheres), spheres[j]);
    int store = 2 * kFast + d.n_over_planes + d.n_over_spheres;       // unscanned storage: n, off by scene index
    for (int ax = 0; ax < 3; ++ax) {
        const int n = int(axis_planes[ax].size());
        d.n_aa[ax] = n;
        for (int j = 0; j < n; ++j) {
            const int i = axis_planes[ax][j];
            d.aa[ax][kAxisFast - n + j].c = R(c.prims[i].v[ax] * c.prims[i].v[3]);  // s * off, exact (s = +-1)
            d.aa[ax][kAxisFast - n + j].id = i;
            put(store++, i);
        }
    }
    for (int i = 0; i < d.n_prims; ++i) {
        const drtb_prim& p = c.prims[i];
        d.type[i] = int8_t(p.type);
        d.color[i] = p.material >= 0 ? c.materials[p.material].color : -1;
        d.mtype[i] = int8_t(p.material >= 0 ? c.materials[p.material].type : DRTB_DIFFUSE);
        d.expo[i] = R(p.material >= 0 ? c.materials[p.material].exponent : 0.0);
        d.emis[i] = p.emission;
        if (p.type == DRTB_PLANE) {
            // make_frame(normal), bxdf.hpp:29-41, in double with the reference's own operation order
            const double n[3] = {p.v[0], p.v[1], p.v[2]};
            const bool ex = std::fabs(n[0]) < std::fabs(n[1]);
            const double e[3] = {ex ? 1.0 : 0.0, ex ? 0.0 : 1.0, 0.0};
            const double en = ex ? n[0] : n[1];
            double t[3], b[3];
            for (int j = 0; j < 3; ++j) t[j] = e[j] - n[j] * en;
            double len = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
            for (int j = 0; j < 3; ++j) t[j] /= len;
            b[0] = n[1] * t[2] - n[2] * t[1]; b[1] = n[2] * t[0] - n[0] * t[2]; b[2] = n[0] * t[1] - n[1] * t[0];
            len = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
            for (int j = 0; j < 3; ++j) { d.frame[i][j] = R(t[j]); d.frame[i][3 + j] = R(b[j] / len); }
            d.wtab[i] = R(3.14159265358979323846 * (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]));   // diffuse_sample: pi |n|^2
        } else {
            d.wtab[i] = R(3.14159265358979323846);
        }
    }
    const drtb_camera& cam = c.camera;
    for (int j = 0; j < 3; ++j) {
        d.eye[j] = R(cam.eye[j]); d.fwd[j] = R(cam.forward[j]); d.right[j] = R(cam.right[j]);
        d.nup[j] = R(-1.0 * cam.up[j]);                       // operator-: -1*v (vector.hpp:320-325)
    }
    d.aspect = R(double(cam.width) / cam.height);            // camera.hpp:48-49
    d.tan_half = R(std::tan(cam.vfov / 2.));                 // camera.hpp:56-57, host libm
    d.inv_w = R(1.0 / cam.width); d.inv_h = R(1.0 / cam.height);
    d.width = cam.width; d.height = cam.height;
    // eye_clear (path.cuh, ClosestMargin): camera rays are unit vectors, so |t| >= |h| / |n| for a plane and
    // >= | |eye - c| - r | for a sphere
    d.eye_clear = 1;
    for (int i = 0; i < d.n_prims; ++i) {
        const drtb_prim& p = c.prims[i];
        if (p.type == DRTB_PLANE) {
            const double h = cam.eye[0] * p.v[0] + cam.eye[1] * p.v[1] + cam.eye[2] * p.v[2] - p.v[3];
            const double nn = std::sqrt(p.v[0] * p.v[0] + p.v[1] * p.v[1] + p.v[2] * p.v[2]);
            bool exact = h == 0.0;
            for (int j = 0; j < 3; ++j)                     // ... and exactly so in float as well
                exact = exact && double(float(cam.eye[j])) == cam.eye[j] && double(float(p.v[j])) == p.v[j];
            exact = exact && double(float(p.v[3])) == p.v[3];
            if (!exact && !(std::fabs(h) >= 2e-4 * nn)) d.eye_clear = 0;
        } else {
            const double ox = cam.eye[0] - p.v[0], oy = cam.eye[1] - p.v[1], oz = cam.eye[2] - p.v[2];
            if (!(std::fabs(std::sqrt(ox * ox + oy * oy + oz * oz) - p.v[3]) >= 2e-4)) d.eye_clear = 0;
        }
    }
}

// How a render's units of work are cut into the chunks that warps claim from the global counter
// (render_kernel: units = warp tasks; render_regen_kernel: units = pixels).  The first n_big chunks hold
// `big` units each, the rest `small` units each (the last one possibly fewer): big chunks keep the claim and
// the per-chunk gradient row cheap, the small ones of the last round keep the tail of the kernel short.
// Pure host arithmetic, exported as drtb_chunk_plan so that the CPU tests can check that every unit is
// covered exactly once for any size.
} // namespace

drtbh::ChunkPlan drtbh::plan_chunks(long long n_units, int spp, long long resident_warps, bool regen, long long forced_big)
{
    ChunkPlan p{1, 1, 0, 0};
    if (n_units <= 0) return p;
    if (resident_warps < 1) resident_warps = 1;
    if (!regen) {
        // about 1024 paths per chunk, but never so large that a warp gets fewer than ~8 chunks; then single tasks
        const long long paths_per_task = spp >= 32 ? spp : 32;
        long long big = std::max<long long>(1, 1024 / paths_per_task);
        big = std::max<long long>(1, std::min(big, n_units / (8 * resident_warps)));
        if (forced_big > 0) big = forced_big;
        const long long small_units = big > 1 ? std::min(n_units, resident_warps * big) : 0;
        p.big = big; p.small = 1;
        p.n_big = (n_units - small_units) / big;
        p.n_chunks = p.n_big + (n_units - p.n_big * big);
    } else {
        // about 1024 samples per chunk, at most kRegenPixels pixels, at least one warp of samples; the last round
        // in chunks of about 64 samples
        const long long full_warp = std::min<long long>(kRegenPixels, (32 + spp - 1) / spp);   // pixels that fill 32 lanes once
        long long big = std::max<long long>(1, std::min<long long>(kRegenPixels, 1024 / spp));
        big = std::max(full_warp, std::min(big, n_units / (8 * resident_warps)));
        if (forced_big > 0) big = std::max(full_warp, std::min<long long>(kRegenPixels, forced_big));
        const long long small = std::max(full_warp, std::min<long long>(big, 64 / spp));
        const long long small_units = big > small ? std::min(n_units, resident_warps * big) : 0;
        p.big = big; p.small = small;
        p.n_big = (n_units - small_units) / big;
        const long long rest = n_units - p.n_big * big;
        p.n_chunks = p.n_big + (rest + small - 1) / small;
    }
    return p;
}

// grad[j] = sum over `rows` partial rows, in an order fixed by `rows` alone.  Up to 4096 rows:
// one block.  More (a large render leaves one row per chunk of warp tasks): a first pass of
// 1024-row blocks into the scratch rows behind the partials, then one block over those.
int drtbh::reduce_partials(drtb_ctx* ctx, double* partial, size_t rows, int P3, double* d_grad, cudaStream_t stream)
{
    if (P3 <= 0 || rows == 0) return DRTB_OK;
    Range r("drtb: gradient reduction");
    if (rows <= kReduceDirectRows) {
        reduce_grad_kernel<<<1, 256, 0, stream>>>(partial, int(rows), int(rows), P3, d_grad);
        ctx->launches++;
    } else {
        double* scratch = partial + rows * P3;
        const int nb = int(reduce_scratch_rows(rows));
        reduce_grad_kernel<<<nb, 256, 0, stream>>>(partial, int(rows), kReduceBlockRows, P3, scratch);
        reduce_grad_kernel<<<1, 256, 0, stream>>>(scratch, nb, nb, P3, d_grad);
        ctx->launches += 2;
    }
    CK(ctx, cudaGetLastError());
    return DRTB_OK;
}

int drtbh::fail(drtb_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return code;
}

namespace {

int validate_opts(drtb_ctx* ctx, const drtb_render_opts* o)
{
    if (!ctx->has_scene) return fail(ctx, DRTB_ERR_INVALID, "no scene uploaded");
    if (!o) return fail(ctx, DRTB_ERR_INVALID, "opts is NULL");
    if (o->spp < 1) return fail(ctx, DRTB_ERR_INVALID, "spp must be >= 1");
    if (o->min_bounces < 0) return fail(ctx, DRTB_ERR_INVALID, "min_bounces must be >= 0");
    if (!(o->absorb >= 0.0 && o->absorb <= 1.0)) return fail(ctx, DRTB_ERR_INVALID, "absorb must be in [0, 1]");
    if (o->precision != DRTB_F64 && o->precision != DRTB_F32 && o->precision != DRTB_MIXED) return fail(ctx, DRTB_ERR_INVALID, "unknown precision");
    if (o->shard_count > 1 && (o->shard_index < 0 || o->shard_index >= o->shard_count || o->band_rows < 1))
        return fail(ctx, DRTB_ERR_INVALID, "bad shard (index, count, band_rows)");
    if (o->max_depth < 0 || o->max_depth > kMaxDepth)
        return fail(ctx, DRTB_ERR_UNSUPPORTED, "max_depth must be in [0, 64]");
    if (ctx->n_tris == 0 && ctx->params.size() > size_t(kMaxParams) * 3)
        return fail(ctx, DRTB_ERR_UNSUPPORTED, "more than 64 RGB parameters needs a mesh scene (drtb_mesh_upload)");
    if (o->absorb == 1.0 && o->min_bounces > kMaxDepth)
        return fail(ctx, DRTB_ERR_UNSUPPORTED, "min_bounces > 64 with absorb == 1 exceeds the vertex record");
    return DRTB_OK;
}

// Enqueue one render (+ gradient reduction) on `stream`; all pointers device.
int launch_render_once(drtb_ctx* ctx, const drtb_render_opts* o, const double* d_seed, double* d_img,
                       double* d_grad, drtb_stats* d_stats, const GradImage& gi, cudaStream_t stream)
{
    const int W = ctx->camera.width, H = ctx->camera.height;
    const int P = int(ctx->params.size() / 3), P3 = P * 3;
    const bool want_grad = (o->flags & DRTB_FLAG_GRAD) != 0;
    const bool want_img = (o->flags & DRTB_FLAG_IMAGE) != 0;
    if (want_grad && !d_grad) return fail(ctx, DRTB_ERR_INVALID, "DRTB_FLAG_GRAD set but grad is NULL");
    const bool peers = want_img && ctx->n_img_peers > 0;
    if (want_img && !d_img && !peers) return fail(ctx, DRTB_ERR_INVALID, "DRTB_FLAG_IMAGE set but img is NULL");
    if (peers && ctx->n_tris > 0) return fail(ctx, DRTB_ERR_UNSUPPORTED, "peer images are filled by the analytic-scene kernel only");
    {
        // A shard that owns no rows (more shards than bands) renders nothing: the gradients it reports are
        // zeros (drtb.h: grad is OVERWRITTEN), its statistics are zeros, and no kernel is launched.
        const int cnt0 = o->shard_count > 1 ? o->shard_count : 1;
        if ((long long)shard_rows_impl(H, o->shard_index, cnt0, o->band_rows) * W == 0) {
            if (ctx->dry) return DRTB_OK;
            if (want_grad && P3) CK(ctx, cudaMemsetAsync(d_grad, 0, sizeof(double) * P3, stream));
            if ((o->flags & DRTB_FLAG_STATS) && d_stats) CK(ctx, cudaMemsetAsync(d_stats, 0, sizeof(drtb_stats), stream));
            return DRTB_OK;
        }
    }
    if (ctx->n_tris > 0) return launch_wavefront(ctx, o, d_seed, d_img, d_grad, d_stats, gi, stream);
    const int cnt = o->shard_count > 1 ? o->shard_count : 1;
    const int rows = shard_rows_impl(H, o->shard_index, cnt, o->band_rows);

    RenderArgs a{};
    a.spp = o->spp; a.min_bounces = o->min_bounces; a.max_depth = effective_max_depth(o);
    a.flags = o->flags; a.absorb = o->absorb; a.key0 = o->seed * kSeedMul;
    a.shard_index = o->shard_index; a.shard_count = cnt; a.band_rows = o->band_rows > 0 ? o->band_rows : 1;
    a.shard_rows = rows; a.seed_scale = o->seed_scale;
    a.params = ctx->d_params; a.seed_img = d_seed; a.img = want_img ? d_img : nullptr;
    a.n_peer_img = peers ? ctx->n_img_peers : 0;
    for (int p = 0; p < a.n_peer_img; ++p) a.peer_img[p] = ctx->img_peers[p];
    a.stats = (o->flags & DRTB_FLAG_STATS) ? d_stats : nullptr;
    const bool want_gimg = want_grad && gi.d_out != nullptr;
    a.gimg = want_gimg ? gi.d_out : nullptr;
    a.gimg_param = want_gimg ? gi.param : -1;
    // general kernels: SpecularBxDF materials, a gradient image, or a plane whose diffuse weight is not a constant
    // (tilted non-unit normal: diffuse_sample in path.cuh) -- the all-diffuse kernels keep no weights in their records
    const bool gen = ctx->has_specular || want_gimg || !ctx->const_weight;

    const bool smallp = P <= kSmallP;
    // DRTB_MIXED: float pass + double re-trace of the close calls, where both kernels exist and the float path stays
    // close to the double one -- fixed-length paths (absorb == 1) of at most kMixedMaxDepth bounces on an all-diffuse
    // analytic scene with <= kSmallP parameters, compact image.  Anything else renders in double, which meets the
    // same promise (parity on every pixel) without the speed-up.
    const bool mixed = o->precision == DRTB_MIXED && o->absorb >= 1.0 && a.max_depth <= kMixedMaxDepth && !gen && smallp && !peers;
    const bool f32 = o->precision == DRTB_F32 || mixed;
    // lit-path compaction needs whole-pixel warp tasks and records that fit the ring
    const bool queue = o->spp >= 32 && a.max_depth <= kQueueDepth;
    constexpr bool mesh = false;                  // mesh scenes took the wavefront above
    a.mesh = mesh_view(ctx);
    size_t smem = (smallp && want_grad) ? size_t(P3) * kBlock * sizeof(double) : 0;
    // the all-diffuse kernels queue primitive lists only, the weights are constants of the primitives (PathRecord)
    const size_t ring_real = !gen ? 0 : f32 ? sizeof(float) : sizeof(double);
    const size_t ring_bytes = queue ? kWarpsPerBlock * queue_bytes_per_warp(a.max_depth, ring_real,
                                                                            mesh ? sizeof(int32_t) : sizeof(uint8_t)) : 0;
    // analytic scenes with 9 .. 64 parameters: shared atomic columns, as many as keep 5 blocks on an SM
    const bool shared_atomic = !smallp && !mesh;
    a.sink_cols = 1;
    if (shared_atomic && want_grad) {
        const size_t room = 44 * 1024 - 6 * 1024;                         // 227 KB / 5 blocks, minus static smem
        const size_t budget = ring_bytes + 8 * 1024 < room ? room - ring_bytes : room;   // a large ring goes to global memory below
        int cols = kBlock;
        while (cols > 1 && size_t(P3) * cols * sizeof(double) > budget) cols >>= 1;
        a.sink_cols = cols;
        smem = size_t(P3) * cols * sizeof(double);
    }
    // The shared ring of a deep record costs resident blocks (double, max_depth 16: 37 KB of ring, 3 blocks instead
    // of 5, -36 % throughput): past the budget the ring moves to a global scratch buffer (QUEUE == 2).
    int queue_kind = queue ? 1 : 0;
    if (queue && !mesh) {
        const int want_blocks = f32 ? DRTB_MIN_BLOCKS_F32 : gen ? DRTB_MIN_BLOCKS_GEN : DRTB_MIN_BLOCKS;
        const size_t static_smem = (f32 ? sizeof(BlockScene<float>) : sizeof(BlockScene<double>) + 3 * kBlock * sizeof(double)) + 1024;   // (+ the double kernels' pixel columns) + 1 KB the system reserves per block
        if ((static_smem + smem + ring_bytes) * want_blocks > size_t(228) * 1024) queue_kind = 2;
    }
    // Measured and NOT adopted for records that fit (B <= 8, double): the global ring at every depth is 0.9 % faster
    // (42.51 against 42.88 ms: the freed shared memory goes to the L1 that holds the local-memory records, hit rate
    // 44 -> 81 %) but the L2 writes the constantly rewritten rings back to HBM, 0.9 GB per render against 0.15 GB --
    // a per-launch persisting access-policy window over the rings did not change that -- and keeping the lanes' own
    // records in the freed shared memory instead of local memory was slower (43.71 ms).  DRTB_RING forces either.
    if (queue_kind == 1 && !mesh && ctx->ring_policy == 2) queue_kind = 2;
    // records of at most 8 vertices (the headline's 8 bounces) in the global ring: the short-record instantiation
    // (render_kernels.cuh, QUEUE == 3), compiled for the all-diffuse kernels with per-thread gradient columns
    if (a.max_depth <= 8 && smallp && !gen && !mixed && queue_kind) queue_kind = queue_kind == 2 ? 3 : 4;
    if (queue_kind == 1 || queue_kind == 4) smem += ring_bytes;       // the shared-memory rings
    smem = (smem + 15) & ~size_t(15);
    const long long npix = (long long)rows * W;
    const int ppw = o->spp >= 32 ? 1 : 32 / o->spp;
    const long long n_tasks = (npix + ppw - 1) / ppw;

    if (a.stats && !ctx->dry) {
        CK(ctx, cudaMemsetAsync(d_stats, 0, sizeof(drtb_stats), stream));
    }
    if (want_grad && !smallp && !shared_atomic) {
        if (!ctx->dry) CK(ctx, cudaMemsetAsync(d_grad, 0, sizeof(double) * P3, stream));
        a.grad_atomic = d_grad;
    }
    size_t partial_rows = 0;
    int rc;
    // Russian roulette on an all-diffuse analytic scene: the path-regenerating kernel
    const bool regen = o->absorb < 1.0 && !mesh && !gen && !ctx->no_regen;
    ctx->partial_extra_rows = 0;
    if (mixed) {
        const size_t cap = std::max<size_t>(size_t(1) << 16, size_t(npix) * size_t(o->spp) / 8);
        if (cap > 0xffffffffull) return fail(ctx, DRTB_ERR_UNSUPPORTED, "DRTB_MIXED: more than 2^35 paths in one render");
        if ((rc = ensure(ctx, ctx->d_retrace, ctx->retrace_cap, cap)) != DRTB_OK) return rc;
        if (!ctx->d_retrace_count) CK(ctx, cudaMalloc((void**)&ctx->d_retrace_count, sizeof(unsigned int)));
        a.retrace_list = ctx->d_retrace; a.retrace_count = ctx->d_retrace_count; a.retrace_cap = (unsigned int)cap;
        if (!ctx->dry) CK(ctx, cudaMemsetAsync(ctx->d_retrace_count, 0, sizeof(unsigned int), stream));
        ctx->partial_extra_rows = want_grad ? size_t(ctx->sm_count) * 4 : 0;
    }
    const AnalyticLaunch l{smallp, queue_kind, gen, regen, mixed, smem, n_tasks, npix, P3, want_grad};
    rc = f32 ? launch_analytic_f32(ctx, a, l, stream, partial_rows) : launch_analytic_f64(ctx, a, l, stream, partial_rows);
    if (rc != DRTB_OK) return rc;
    if (mixed) {
        size_t extra = 0;
        if ((rc = launch_retrace_f64(ctx, a, P3, want_grad, partial_rows, stream, extra)) != DRTB_OK) return rc;
        partial_rows += extra;
    }
    if (ctx->dry) return DRTB_OK;
    if (want_grad && (smallp || shared_atomic)) return reduce_partials(ctx, ctx->d_partial, partial_rows, P3, d_grad, stream);
    return DRTB_OK;
}

// Decorrelated adjoint (the reference's integrate_unbiased idea, integrate.hpp:39-52:
// fresh samples for the backward pass): the image comes from stream `seed`, the
// gradients from stream `adjoint_seed`.  With a counter-based RNG that is simply
// a second, gradient-only launch keyed differently.
int launch_render_local(drtb_ctx* ctx, const drtb_render_opts* o, const double* d_seed, double* d_img,
                        double* d_grad, drtb_stats* d_stats, const GradImage& gi, cudaStream_t stream);

// One render of this GPU's shard and, with drtb_set_grad_peers, the sum of the gradients over the GPUs of the job.
int launch_render(drtb_ctx* ctx, const drtb_render_opts* o, const double* d_seed, double* d_img,
                  double* d_grad, drtb_stats* d_stats, const GradImage& gi, cudaStream_t stream)
{
    const int P3 = int(ctx->params.size());
    const bool exchange = ctx->n_grad_peers > 1 && (o->flags & DRTB_FLAG_GRAD) && P3 > 0;
    if (exchange && P3 > kMaxExchangeP3)
        return fail(ctx, DRTB_ERR_UNSUPPORTED, "the peer gradient exchange serves up to 4096 gradient scalars; use a collective library for more");
    int rc;
    {
        Range r(ctx->dry ? "drtb: prepare (scratch, kernel attributes, first-use launches)"
                         : (o->flags & DRTB_FLAG_GRAD) ? "drtb: render forward + adjoint" : "drtb: render forward");
        rc = launch_render_local(ctx, o, d_seed, d_img, d_grad, d_stats, gi, stream);
    }
    if (rc != DRTB_OK || ctx->dry || !exchange) return rc;
    Range r("drtb: gradient exchange over NVLink peers");
    GradPeers gp{};
    for (int p = 0; p < ctx->n_grad_peers; ++p) gp.buf[p] = ctx->grad_peers[p];
    grad_allreduce_kernel<<<1, 256, 0, stream>>>(gp, ctx->n_grad_peers, ctx->grad_rank, P3, ++ctx->grad_epoch, d_grad);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    return DRTB_OK;
}

int launch_render_local(drtb_ctx* ctx, const drtb_render_opts* o, const double* d_seed, double* d_img,
                        double* d_grad, drtb_stats* d_stats, const GradImage& gi, cudaStream_t stream)
{
    if (gi.d_out) {
        if (!(o->flags & DRTB_FLAG_GRAD)) return fail(ctx, DRTB_ERR_INVALID, "a gradient image needs DRTB_FLAG_GRAD");
        if (gi.param < 0 || size_t(gi.param) * 3 >= ctx->params.size()) return fail(ctx, DRTB_ERR_INVALID, "gradient-image parameter index out of range");
    }
    const bool split = o->adjoint_seed != 0 && (o->flags & DRTB_FLAG_GRAD) && (o->flags & DRTB_FLAG_IMAGE);
    if (!split) {
        drtb_render_opts one = *o;
        if (o->adjoint_seed != 0 && (o->flags & DRTB_FLAG_GRAD)) one.seed = o->adjoint_seed;   // gradient-only call
        return launch_render_once(ctx, &one, d_seed, d_img, d_grad, d_stats, gi, stream);
    }
    drtb_render_opts fwd = *o, adj = *o;
    fwd.flags &= ~DRTB_FLAG_GRAD;
    adj.flags &= ~(DRTB_FLAG_IMAGE | DRTB_FLAG_STATS);       // stats describe the image pass
    adj.seed = o->adjoint_seed;
    int rc = launch_render_once(ctx, &fwd, nullptr, d_img, nullptr, d_stats, GradImage{}, stream);
    if (rc != DRTB_OK) return rc;
    return launch_render_once(ctx, &adj, d_seed, nullptr, d_grad, nullptr, gi, stream);
}

} // namespace

extern "C" {

int drtb_abi_version(void) { return DRTB_ABI_VERSION; }

int drtb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

size_t drtb_struct_size(int which)
{
    switch (which) {
        case 0: return sizeof(drtb_prim);
        case 1: return sizeof(drtb_material);
        case 2: return sizeof(drtb_camera);
        case 3: return sizeof(drtb_scene);
        case 4: return sizeof(drtb_render_opts);
        case 5: return sizeof(drtb_stats);
        case 6: return sizeof(drtb_mesh);
        default: return 0;
    }
}

int drtb_create(int device, drtb_ctx** out)
{
    if (!out) return fail(nullptr, DRTB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, DRTB_ERR_NO_DEVICE,
                    std::string("no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                        "); this library has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(nullptr, DRTB_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, DRTB_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, DRTB_ERR_UNSUPPORTED,
                    std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                        "; this library carries sm_100a code only");
    drtb_ctx* ctx = new (std::nothrow) drtb_ctx;
    if (!ctx) return fail(nullptr, DRTB_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_window_max = size_t(std::max(0, prop.accessPolicyMaxWindowSize));
    if (const char* e = std::getenv("DRTB_NO_REGEN")) ctx->no_regen = std::atoi(e) != 0;
    if (const char* e = std::getenv("DRTB_RING")) ctx->ring_policy = std::string(e) == "global" ? 2 : 0;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_stats, sizeof(drtb_stats)) != cudaSuccess) {
        std::string m = cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return fail(nullptr, DRTB_ERR_CUDA, "context setup failed: " + m);
    }
    // the L2 set-aside for a mesh's geometry (mesh.cu) is reserved when a mesh is attached, not here: a persisting
    // carve-out shrinks the L2 every other kernel sees (the analytic megakernel's 52 MB of vertex records then spill
    // to HBM: 6.1 GB of write-backs per render against 0.15 GB, profiles/README.md round 2)
    ctx->l2_persist_max = size_t(std::max(0, prop.persistingL2CacheMaxSize));
    // (sin, cos)(2 pi i 2^23 / M) for Real<double>::sincos_tab: one copy per translation unit that samples in double
    if (drtb::upload_sincos_tab() != cudaSuccess || init_tables_render_f64() != cudaSuccess || init_tables_mesh() != cudaSuccess) {
        std::string m = cudaGetErrorString(cudaGetLastError());
        drtb_destroy(ctx);
        return fail(nullptr, DRTB_ERR_CUDA, "context setup failed: " + m);
    }
    *out = ctx;
    return DRTB_OK;
}

void drtb_destroy(drtb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    cudaFree(ctx->d_params); cudaFree(ctx->d_partial); cudaFree(ctx->d_img); cudaFree(ctx->d_task_counter); cudaFree(ctx->d_ring);
    cudaFree(ctx->d_retrace); cudaFree(ctx->d_retrace_count);
    cudaFree(ctx->d_seed); cudaFree(ctx->d_grad); cudaFree(ctx->d_gimg); cudaFree(ctx->d_stats); cudaFree(ctx->wf_mem);
    free_mesh(ctx);
    delete ctx;
}

const char* drtb_last_error(const drtb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int drtb_scene_upload(drtb_ctx* ctx, const drtb_scene* s)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (!s || s->n_prims < 0 || s->n_materials < 0 || s->n_params < 0 || (s->n_prims && !s->prims) ||
        (s->n_materials && !s->materials) || (s->n_params && !s->params))
        return fail(ctx, DRTB_ERR_INVALID, "scene has NULL arrays or negative counts");
    if (s->camera.width < 1 || s->camera.height < 1) return fail(ctx, DRTB_ERR_INVALID, "camera width/height must be >= 1");
    if (s->n_prims > kMaxPrims)
        return fail(ctx, DRTB_ERR_UNSUPPORTED, "more than 32 analytic primitives is not supported by this build");
    for (int m = 0; m < s->n_materials; ++m) {
        if (s->materials[m].type != DRTB_DIFFUSE && s->materials[m].type != DRTB_SPECULAR) return fail(ctx, DRTB_ERR_INVALID, "unknown material type");
        if (s->materials[m].color < 0 || s->materials[m].color >= s->n_params) return fail(ctx, DRTB_ERR_INVALID, "material colour index out of range");
    }
    for (int i = 0; i < s->n_prims; ++i) {
        const drtb_prim& p = s->prims[i];
        if (p.type != DRTB_SPHERE && p.type != DRTB_PLANE) return fail(ctx, DRTB_ERR_INVALID, "unknown primitive type");
        if (p.material < -1 || p.material >= s->n_materials) return fail(ctx, DRTB_ERR_INVALID, "primitive material index out of range");
        if (p.emission < -1 || p.emission >= s->n_params) return fail(ctx, DRTB_ERR_INVALID, "primitive emission index out of range");
    }
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    free_mesh(ctx);                                   // a mesh belongs to the scene it was attached to
    ctx->prims.assign(s->prims, s->prims + s->n_prims);
    ctx->materials.assign(s->materials, s->materials + s->n_materials);
    ctx->params.assign(s->params, s->params + size_t(s->n_params) * 3);
    ctx->camera = s->camera;
    ctx->has_specular = false;
    for (const drtb_prim& p : ctx->prims)
        if (p.material >= 0 && ctx->materials[p.material].type == DRTB_SPECULAR) ctx->has_specular = true;
    fill_dev_scene(ctx->sc64, *ctx);
    fill_dev_scene(ctx->sc32, *ctx);
    // the diffuse weight of a plane is pi |n|^2 + pi (n.tg) x / cos(theta) (diffuse_sample, path.cuh): a constant unless
    // n.tg != 0, i.e. unless the normal is neither a unit vector nor orthogonal to make_frame's helper axis
    ctx->const_weight = true;
    for (size_t i = 0; i < ctx->prims.size(); ++i) {
        const drtb_prim& p = ctx->prims[i];
        if (p.type != DRTB_PLANE) continue;
        const double nt = p.v[0] * ctx->sc64.frame[i][0] + p.v[1] * ctx->sc64.frame[i][1] + p.v[2] * ctx->sc64.frame[i][2];
        const double nn = std::sqrt(p.v[0] * p.v[0] + p.v[1] * p.v[1] + p.v[2] * p.v[2]);
        if (!(std::fabs(nt) <= 1e-14 * nn)) ctx->const_weight = false;
    }
    int rc = ensure(ctx, ctx->d_params, ctx->params_cap, std::max<size_t>(3, ctx->params.size()));
    if (rc != DRTB_OK) return rc;
    if (!ctx->params.empty())
        CK(ctx, cudaMemcpy(ctx->d_params, ctx->params.data(), sizeof(double) * ctx->params.size(), cudaMemcpyHostToDevice));
    ctx->has_scene = true;
    return DRTB_OK;
}

int drtb_mesh_upload(drtb_ctx* ctx, const drtb_mesh* mesh)
{
    if (!ctx) return DRTB_ERR_INVALID;
    return mesh_upload(ctx, mesh);
}

double drtb_mesh_build_ms(const drtb_ctx* ctx) { return ctx && ctx->n_tris > 0 ? ctx->mesh_build_ms : 0.0; }

int drtb_set_params(drtb_ctx* ctx, const double* params, int32_t n_params)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, DRTB_ERR_INVALID, "no scene uploaded");
    if (!params || size_t(n_params) * 3 != ctx->params.size()) return fail(ctx, DRTB_ERR_INVALID, "n_params does not match the uploaded scene");
    CK(ctx, cudaSetDevice(ctx->device));
    ctx->params.assign(params, params + size_t(n_params) * 3);
    // stream-ordered so that it cannot overtake a render still in flight
    CK(ctx, cudaMemcpyAsync(ctx->d_params, ctx->params.data(), sizeof(double) * ctx->params.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DRTB_OK;
}

int drtb_set_params_device(drtb_ctx* ctx, const double* d_params, int32_t n_params, void* stream)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, DRTB_ERR_INVALID, "no scene uploaded");
    if (!d_params || size_t(n_params) * 3 != ctx->params.size()) return fail(ctx, DRTB_ERR_INVALID, "n_params does not match the uploaded scene");
    CK(ctx, cudaSetDevice(ctx->device));
    // ordered on the caller's stream with the renders enqueued there; no host round trip
    CK(ctx, cudaMemcpyAsync(ctx->d_params, d_params, sizeof(double) * ctx->params.size(), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DRTB_OK;
}

int drtb_chunk_plan(int64_t n_units, int32_t spp, int64_t resident_warps, int32_t regen, int64_t out[4])
{
    if (!out || n_units < 0 || spp < 1) return DRTB_ERR_INVALID;
    const ChunkPlan p = plan_chunks(n_units, spp, resident_warps, regen != 0, 0);
    out[0] = p.big; out[1] = p.small; out[2] = p.n_big; out[3] = p.n_chunks;
    return DRTB_OK;
}

int32_t drtb_shard_rows(int32_t height, int32_t shard_index, int32_t shard_count, int32_t band_rows)
{
    if (shard_count > 1 && (shard_index < 0 || shard_index >= shard_count)) return 0;
    return shard_rows_impl(height, shard_index, shard_count, band_rows);
}

namespace {

// drtb_render / drtb_render_grad_image with host buffers
int render_host(drtb_ctx* ctx, const drtb_render_opts* o, int32_t gparam, const double* seed_img, double* img,
                double* grad, double* grad_img, drtb_stats* stats)
{
    if (!ctx) return DRTB_ERR_INVALID;
    Range whole("drtb_render (host buffers)");
    int rc = validate_opts(ctx, o);
    if (rc != DRTB_OK) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    const int W = ctx->camera.width, H = ctx->camera.height;
    const int P3 = int(ctx->params.size());
    const int rows = shard_rows_impl(H, o->shard_index, o->shard_count, o->band_rows);
    const size_t npx3 = size_t(rows) * W * 3;
    const bool want_grad = (o->flags & DRTB_FLAG_GRAD) != 0, want_img = (o->flags & DRTB_FLAG_IMAGE) != 0;
    if (want_img && !img) return fail(ctx, DRTB_ERR_INVALID, "DRTB_FLAG_IMAGE set but img is NULL");
    if (want_grad && !grad) return fail(ctx, DRTB_ERR_INVALID, "DRTB_FLAG_GRAD set but grad is NULL");
    if (want_img && (rc = ensure(ctx, ctx->d_img, ctx->img_cap, std::max<size_t>(npx3, 3))) != DRTB_OK) return rc;
    if (want_grad && (rc = ensure(ctx, ctx->d_grad, ctx->grad_cap, std::max<size_t>(P3, 3))) != DRTB_OK) return rc;
    GradImage gi;
    if (grad_img) {
        if ((rc = ensure(ctx, ctx->d_gimg, ctx->gimg_cap, std::max<size_t>(npx3, 3))) != DRTB_OK) return rc;
        gi.param = gparam; gi.d_out = ctx->d_gimg;
    }
    cudaStream_t st = ctx->stream;
    if (seed_img) {
        if ((rc = ensure(ctx, ctx->d_seed, ctx->seed_cap, std::max<size_t>(npx3, 3))) != DRTB_OK) return rc;
        CK(ctx, cudaMemcpyAsync(ctx->d_seed, seed_img, sizeof(double) * npx3, cudaMemcpyHostToDevice, st));
    }
    drtb_render_opts oo = *o;
    if (stats) oo.flags |= DRTB_FLAG_STATS;
    // A PINNED host image (cudaHostAlloc / cudaHostRegister: it has a device address) is written by the kernel
    // itself: the analytic-scene kernels store every pixel exactly once with plain stores, 24 bytes per ~10^5
    // instructions of tracing, so the image crosses PCIe under the compute and no copy follows the kernel (0.46 ms
    // of a 36 ms step at 1024^2).  Mesh scenes and DRTB_MIXED sum into the image with atomics and keep the copy.
    double* img_target = ctx->d_img;
    bool img_direct = false;
    if (want_img && npx3 && ctx->n_tris == 0 && o->precision != DRTB_MIXED) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, img) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            img_target = static_cast<double*>(at.devicePointer);
            img_direct = true;
        } else {
            (void)cudaGetLastError();
        }
    }
    // all host-side preparation (scratch sizing, kernel attributes, first-use launches) happens before the timer
    ctx->dry = true;
    rc = launch_render(ctx, &oo, seed_img ? ctx->d_seed : nullptr, img_target, ctx->d_grad, ctx->d_stats, gi, st);
    ctx->dry = false;
    if (rc != DRTB_OK) return rc;
    CK(ctx, cudaEventRecord(ctx->ev0, st));
    rc = launch_render(ctx, &oo, seed_img ? ctx->d_seed : nullptr, img_target, ctx->d_grad, ctx->d_stats, gi, st);
    if (rc != DRTB_OK) return rc;
    CK(ctx, cudaEventRecord(ctx->ev1, st));
    if (want_img && npx3 && !img_direct) CK(ctx, cudaMemcpyAsync(img, ctx->d_img, sizeof(double) * npx3, cudaMemcpyDeviceToHost, st));
    if (want_grad && P3) CK(ctx, cudaMemcpyAsync(grad, ctx->d_grad, sizeof(double) * P3, cudaMemcpyDeviceToHost, st));
    if (grad_img && npx3) CK(ctx, cudaMemcpyAsync(grad_img, ctx->d_gimg, sizeof(double) * npx3, cudaMemcpyDeviceToHost, st));
    if (stats) CK(ctx, cudaMemcpyAsync(stats, ctx->d_stats, sizeof(drtb_stats), cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    if (o->precision == DRTB_MIXED && ((want_img && npx3 && std::isnan(img[0])) || (want_grad && P3 && std::isnan(grad[0])))) {
        // more close calls than the re-trace list holds (the kernels poison the outputs rather than return them
        // incomplete): render in double instead
        drtb_render_opts exact = *o;
        exact.precision = DRTB_F64;
        return render_host(ctx, &exact, gparam, seed_img, img, grad, grad_img, stats);
    }
    if (stats) {
        float ms = 0.f;
        CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        stats->kernel_ms = ms;
#ifndef DRTB_TRAV_DEBUG                               // (the debug build of the traversal borrows this field)
        stats->paths = uint64_t(rows) * W * o->spp;
#endif
    }
    return DRTB_OK;
}

int render_device(drtb_ctx* ctx, const drtb_render_opts* o, const GradImage& gi, const double* d_seed_img, double* d_img,
                  double* d_grad, drtb_stats* d_stats, void* stream)
{
    if (!ctx) return DRTB_ERR_INVALID;
    int rc = validate_opts(ctx, o);
    if (rc != DRTB_OK) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    if ((o->flags & DRTB_FLAG_STATS) && !d_stats) return fail(ctx, DRTB_ERR_INVALID, "DRTB_FLAG_STATS set but d_stats is NULL");
    return launch_render(ctx, o, d_seed_img, d_img, d_grad, d_stats, gi, (cudaStream_t)stream);
}

} // namespace

int drtb_render(drtb_ctx* ctx, const drtb_render_opts* o, const double* seed_img, double* img,
                double* grad, drtb_stats* stats)
{
    return render_host(ctx, o, -1, seed_img, img, grad, nullptr, stats);
}

int drtb_reserve(drtb_ctx* ctx, const drtb_render_opts* o)
{
    if (!ctx) return DRTB_ERR_INVALID;
    int rc = validate_opts(ctx, o);
    if (rc != DRTB_OK) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    if ((rc = ensure(ctx, ctx->d_grad, ctx->grad_cap, std::max<size_t>(ctx->params.size(), 3))) != DRTB_OK) return rc;
    // the preparation pass never dereferences the output pointers; any non-NULL device address will do
    ctx->dry = true;
    rc = launch_render(ctx, o, nullptr, ctx->d_grad, ctx->d_grad, ctx->d_stats, GradImage{}, ctx->stream);
    ctx->dry = false;
    if (rc != DRTB_OK) return rc;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DRTB_OK;
}

int drtb_render_grad_image(drtb_ctx* ctx, const drtb_render_opts* o, int32_t param, const double* seed_img, double* img,
                           double* grad, double* grad_img, drtb_stats* stats)
{
    if (ctx && !grad_img) return fail(ctx, DRTB_ERR_INVALID, "grad_img is NULL");
    return render_host(ctx, o, param, seed_img, img, grad, grad_img, stats);
}

int drtb_render_device(drtb_ctx* ctx, const drtb_render_opts* o, const double* d_seed_img, double* d_img,
                       double* d_grad, drtb_stats* d_stats, void* stream)
{
    return render_device(ctx, o, GradImage{}, d_seed_img, d_img, d_grad, d_stats, stream);
}

int drtb_render_grad_image_device(drtb_ctx* ctx, const drtb_render_opts* o, int32_t param, const double* d_seed_img,
                                  double* d_img, double* d_grad, double* d_grad_img, drtb_stats* d_stats, void* stream)
{
    if (ctx && !d_grad_img) return fail(ctx, DRTB_ERR_INVALID, "d_grad_img is NULL");
    GradImage gi;
    gi.param = param; gi.d_out = d_grad_img;
    return render_device(ctx, o, gi, d_seed_img, d_img, d_grad, d_stats, stream);
}

int drtb_set_image_peers(drtb_ctx* ctx, double* const* full_images, int32_t n)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (n < 0 || n > kMaxPeers) return fail(ctx, DRTB_ERR_INVALID, "between 0 and 8 peer images");
    if (n > 0 && !full_images) return fail(ctx, DRTB_ERR_INVALID, "full_images is NULL");
    for (int p = 0; p < n; ++p)
        if (!full_images[p]) return fail(ctx, DRTB_ERR_INVALID, "a peer image pointer is NULL");
    for (int p = 0; p < kMaxPeers; ++p) ctx->img_peers[p] = p < n ? full_images[p] : nullptr;
    ctx->n_img_peers = n;
    return DRTB_OK;
}

size_t drtb_grad_exchange_bytes(int32_t n_ranks, int32_t n_params)
{
    if (n_ranks < 1 || n_ranks > kMaxPeers || n_params < 0) return 0;
    return (size_t(2) * n_ranks * n_params * 3 + n_ranks) * sizeof(double);
}

int drtb_set_grad_peers(drtb_ctx* ctx, void* const* exchange, int32_t n, int32_t rank)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (n < 0 || n > kMaxPeers) return fail(ctx, DRTB_ERR_INVALID, "between 0 and 8 gradient peers");
    if (n > 0 && (!exchange || rank < 0 || rank >= n)) return fail(ctx, DRTB_ERR_INVALID, "exchange is NULL or rank out of range");
    for (int p = 0; p < n; ++p)
        if (!exchange[p]) return fail(ctx, DRTB_ERR_INVALID, "an exchange buffer pointer is NULL");
    for (int p = 0; p < kMaxPeers; ++p) ctx->grad_peers[p] = p < n ? static_cast<double*>(exchange[p]) : nullptr;
    ctx->n_grad_peers = n;
    ctx->grad_rank = rank;
    ctx->grad_epoch = 0;
    return DRTB_OK;
}

int drtb_host_alloc(size_t bytes, void** ptr)
{
    if (!ptr || bytes == 0) return DRTB_ERR_INVALID;
    *ptr = nullptr;
    const cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *ptr = nullptr;
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? DRTB_ERR_NO_DEVICE : DRTB_ERR_CUDA;
    }
    return DRTB_OK;
}

int drtb_host_free(void* ptr)
{
    if (!ptr) return DRTB_OK;
    if (cudaFreeHost(ptr) != cudaSuccess) { (void)cudaGetLastError(); return DRTB_ERR_CUDA; }
    return DRTB_OK;
}

int drtb_ipc_alloc(drtb_ctx* ctx, size_t bytes, void** d_ptr, void* handle)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (!d_ptr || !handle || bytes == 0) return fail(ctx, DRTB_ERR_INVALID, "drtb_ipc_alloc: NULL argument or zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == DRTB_IPC_HANDLE_BYTES, "IPC handle size");
    CK(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    CK(ctx, cudaMalloc(&p, bytes));
    if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaFree(p); return fail(ctx, DRTB_ERR_CUDA, "cudaMemset of the shared allocation failed"); }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(ctx, DRTB_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    std::memcpy(handle, &h, sizeof(h));
    *d_ptr = p;
    return DRTB_OK;
}

int drtb_ipc_open(drtb_ctx* ctx, const void* handle, void** d_ptr)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (!d_ptr || !handle) return fail(ctx, DRTB_ERR_INVALID, "drtb_ipc_open: NULL argument");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CK(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DRTB_OK;
}

int drtb_ipc_close(drtb_ctx* ctx, void* d_ptr)
{
    if (!ctx) return DRTB_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaIpcCloseMemHandle(d_ptr));
    return DRTB_OK;
}

int drtb_ipc_free(drtb_ctx* ctx, void* d_ptr)
{
    if (!ctx) return DRTB_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaFree(d_ptr));
    return DRTB_OK;
}

int drtb_trace_rays(drtb_ctx* ctx, const drtb_render_opts* o, int64_t n, const double* orig, const double* dir,
                    const uint64_t* keys, double* radiance, double* jac)
{
    if (!ctx) return DRTB_ERR_INVALID;
    int rc = validate_opts(ctx, o);
    if (rc != DRTB_OK) return rc;
    if (n < 0 || (n && (!orig || !dir || !keys || !radiance))) return fail(ctx, DRTB_ERR_INVALID, "NULL ray buffers");
    if (n == 0) return DRTB_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    const int P3 = int(ctx->params.size());
    double *d_o = nullptr, *d_d = nullptr, *d_r = nullptr, *d_j = nullptr;
    uint64_t* d_k = nullptr;
    auto cleanup = [&]() { cudaFree(d_o); cudaFree(d_d); cudaFree(d_r); cudaFree(d_j); cudaFree(d_k); };
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, DRTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    const size_t b3 = sizeof(double) * 3 * size_t(n);
    CKF(cudaMalloc((void**)&d_o, b3)); CKF(cudaMalloc((void**)&d_d, b3)); CKF(cudaMalloc((void**)&d_r, b3));
    CKF(cudaMalloc((void**)&d_k, sizeof(uint64_t) * size_t(n)));
    if (jac && P3) { CKF(cudaMalloc((void**)&d_j, sizeof(double) * P3 * size_t(n))); CKF(cudaMemsetAsync(d_j, 0, sizeof(double) * P3 * size_t(n), ctx->stream)); }
    CKF(cudaMemcpyAsync(d_o, orig, b3, cudaMemcpyHostToDevice, ctx->stream));
    CKF(cudaMemcpyAsync(d_d, dir, b3, cudaMemcpyHostToDevice, ctx->stream));
    CKF(cudaMemcpyAsync(d_k, keys, sizeof(uint64_t) * size_t(n), cudaMemcpyHostToDevice, ctx->stream));
    const int grid = int((n + kBlock - 1) / kBlock);
    const int md = effective_max_depth(o);
    const MeshView mv = mesh_view(ctx);
    const bool f32 = o->precision == DRTB_F32;
    if (ctx->n_tris > 0) {
        if (f32) trace_rays_kernel<float, true><<<grid, kBlock, 0, ctx->stream>>>(ctx->sc32, ctx->d_params, mv, o->flags, o->min_bounces, o->absorb, md, n, d_o, d_d, d_k, d_r, d_j);
        else     trace_rays_kernel<double, true><<<grid, kBlock, 0, ctx->stream>>>(ctx->sc64, ctx->d_params, mv, o->flags, o->min_bounces, o->absorb, md, n, d_o, d_d, d_k, d_r, d_j);
    } else {
        if (f32) trace_rays_kernel<float, false><<<grid, kBlock, 0, ctx->stream>>>(ctx->sc32, ctx->d_params, mv, o->flags, o->min_bounces, o->absorb, md, n, d_o, d_d, d_k, d_r, d_j);
        else     trace_rays_kernel<double, false><<<grid, kBlock, 0, ctx->stream>>>(ctx->sc64, ctx->d_params, mv, o->flags, o->min_bounces, o->absorb, md, n, d_o, d_d, d_k, d_r, d_j);
    }
    CKF(cudaGetLastError());
    ctx->launches++;
    CKF(cudaMemcpyAsync(radiance, d_r, b3, cudaMemcpyDeviceToHost, ctx->stream));
    if (jac && P3) CKF(cudaMemcpyAsync(jac, d_j, sizeof(double) * P3 * size_t(n), cudaMemcpyDeviceToHost, ctx->stream));
    CKF(cudaStreamSynchronize(ctx->stream));
#undef CKF
    cleanup();
    return DRTB_OK;
}

int drtb_fma_peak(drtb_ctx* ctx, int32_t precision, double* tflops)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (!tflops) return fail(ctx, DRTB_ERR_INVALID, "tflops is NULL");
    if (precision != DRTB_F64 && precision != DRTB_F32) return fail(ctx, DRTB_ERR_INVALID, "unknown precision");
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = ensure(ctx, ctx->d_grad, ctx->grad_cap, 3);
    if (rc != DRTB_OK) return rc;
    const int grid = ctx->sm_count * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        if (precision == DRTB_F64) fma_peak_kernel<double><<<grid, 256, 0, ctx->stream>>>(ctx->d_grad, iters, 0.999999, 1e-7);
        else fma_peak_kernel<float><<<grid, 256, 0, ctx->stream>>>((float*)ctx->d_grad, iters, 0.999999f, 1e-7f);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
        CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flop = double(grid) * 256.0 * iters * 64.0 * 2.0;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    return DRTB_OK;
}

uint64_t drtb_launch_count(const drtb_ctx* ctx) { return ctx ? ctx->launches : 0; }

uint32_t drtb_stream_draw(uint64_t key, uint32_t slot) { return drtb::stream_draw(key, slot); }

} // extern "C"
