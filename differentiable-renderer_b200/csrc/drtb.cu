// drtb.cu — kernels + the C ABI of include/drtb.h.
//
// Kernel inventory
//   render_kernel<R, SMALLP>   persistent megakernel: one lane = one path
//                              (camera sample -> trace -> radiance -> adjoint),
//                              lanes of a warp = consecutive samples of one pixel;
//                              warps claim chunks of pixels from a global counter
//   render_regen_kernel<R, SMALLP>  the same pixel loop for Russian-roulette renders, with path
//                              regeneration over a chunk of pixels
//   reduce_grad_kernel         fixed-order sum of the per-chunk gradient partials
//   trace_rays_kernel<R>       Pathtracer::trace on explicit rays (+ Jacobian)
//   fma_peak_kernel<R>         FMA issue-rate micro-benchmark (roofline denominator)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see build.py).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "path.cuh"
#include "wavefront.cuh"

using namespace drtb;

// ===========================================================================
// device code
// ===========================================================================
namespace {

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Gradient sinks ------------------------------------------------------------
// Small parameter sets (the Cornell box has 4): every thread owns one column of
// a [n_params*3][kBlock] shared array -- no atomics, no bank conflicts, and a
// fixed summation order, so gradients are bit-reproducible run to run.
struct SmemSink {
    double* col;                                   // &acc[threadIdx.x]
    template <typename R> __device__ __forceinline__ void add(int p, int c, R v)
    {
        col[(3 * p + c) * kBlock] += double(v);
    }
};
// Medium parameter sets (9 .. kMaxParams, analytic scenes): the columns no longer fit per
// thread, so `cols` (a power of two, chosen by the launcher to fit shared memory) columns are
// shared by the threads with equal (threadIdx.x mod cols) and updated with shared-memory
// atomics (a CAS loop, ATOMS.CAST.SPIN.64).  Contention stays inside the block and is spread
// over P3 x cols words; the block reduction and reduce_grad_kernel are the small-set ones.
// (Global atomics here cost 8x the whole render at 9 parameters: every lit path of the grid
// hammers the same 27 words.)
struct SmemAtomicSink {
    double* col;                                   // &acc[threadIdx.x & (cols - 1)]
    int cols;
    template <typename R> __device__ __forceinline__ void add(int p, int c, R v)
    {
        atomicAdd(col + (3 * p + c) * cols, double(v));
    }
};
// Large parameter sets (mesh scenes, per-triangle albedos): one red.global.add.f64 per contribution.
struct AtomicSink {
    double* grad;
    template <typename R> __device__ __forceinline__ void add(int p, int c, R v)
    {
        atomicAdd(grad + 3 * p + c, double(v));
    }
};
// Gradient image (drtb_render_grad_image): parameter kp's contributions are
// additionally summed into the lane's per-pixel accumulator g[3].
template <typename Inner>
struct PixelSink {
    Inner inner;
    int kp;
    double* g;
    template <typename R> __device__ __forceinline__ void add(int p, int c, R v)
    {
        inner.add(p, c, v);
        if (p == kp) g[c] += double(v);
    }
};
struct JacSink {
    double* row;                                   // this ray's n_params x 3 block
    template <typename R> __device__ __forceinline__ void add(int p, int c, R v)
    {
        row[3 * p + c] += double(v);
    }
};

// The pixel loop of src/render.cpp:72-86.
//   SMALLP: <= kSmallP parameters, gradients in per-thread shared columns
//   QUEUE : spp >= 32 and max_depth <= kQueueDepth: lit paths are compacted through a per-warp
//           ring before the sweeps.  0: no ring; 1: the ring lives in shared memory; 2: in a global scratch
//           buffer (L1/L2 resident), chosen when the shared ring of a deep record (max_depth > 8
//           in double) would cost resident blocks -- the ring carries only the ~16-21 % of the
//           paths that are lit, so its latency does not matter, the occupancy does
#ifndef DRTB_MIN_BLOCKS
#define DRTB_MIN_BLOCKS 1
#endif
#ifndef DRTB_MESH_MIN_BLOCKS
#define DRTB_MESH_MIN_BLOCKS DRTB_MIN_BLOCKS
#endif
#ifndef DRTB_MIN_BLOCKS_F32
#define DRTB_MIN_BLOCKS_F32 DRTB_MIN_BLOCKS
#endif
//   MESH  : a triangle mesh + BVH is attached (ids are 32-bit, parameters in global memory)
//   GEN   : the general variant -- SpecularBxDF materials (bxdf.hpp:85-124) and the
//           per-pixel gradient image; the all-diffuse kernels do not carry that code
template <typename R, bool SMALLP, int QUEUE, bool MESH, bool GEN>
__global__ void __launch_bounds__(kBlock, MESH ? DRTB_MESH_MIN_BLOCKS : sizeof(R) == 4 ? DRTB_MIN_BLOCKS_F32 : DRTB_MIN_BLOCKS)
render_kernel(const __grid_constant__ DevScene<R> sc, const __grid_constant__ RenderArgs a)
{
    using Id = typename PrimId<MESH>::type;
    extern __shared__ double s_dyn[];              // [acc: n_params*3*kBlock doubles][rings]
    __shared__ BlockScene<R> bs;

    // gradient sink: per-thread columns (SMALLP), shared atomic columns (analytic scenes with
    // more parameters) or global atomics (mesh scenes with more parameters)
    constexpr bool kSharedAtomic = !SMALLP && !MESH;
    const bool want_grad = (a.flags & DRTB_FLAG_GRAD) != 0;
    const int P3 = sc.n_params * 3;
    double* s_acc = s_dyn;
    const int acc_doubles = !want_grad ? 0 : SMALLP ? P3 * kBlock : kSharedAtomic ? P3 * a.sink_cols : 0;
    load_block_scene(bs, sc, a.params);
    for (int i = threadIdx.x; i < acc_doubles; i += kBlock) s_acc[i] = 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = sc.width, spp = a.spp;
    const long long npix = (long long)a.shard_rows * W;
    // spp >= 32: one pixel per warp task, ceil(spp/32) passes over its samples;
    // spp <  32: floor(32/spp) pixels per warp task, one pass.
    const int ppw = spp >= 32 ? 1 : 32 / spp;
    const int passes = spp >= 32 ? (spp + 31) / 32 : 1;
    const R inv_p = a.absorb < 1.0 ? R(1.0 / (1.0 - a.absorb)) : R(0);

    // this warp's ring (QUEUE only)
    const int qdepth = a.max_depth;
    unsigned char* ring = reinterpret_cast<unsigned char*>(s_dyn + acc_doubles) +
                          (QUEUE ? size_t(warp) * queue_bytes_per_warp(qdepth, sizeof(R), sizeof(Id)) : 0);
    if constexpr (QUEUE == 2)
        ring = a.ring_scratch + (size_t(blockIdx.x) * kWarpsPerBlock + warp) * queue_bytes_per_warp(qdepth, sizeof(R), sizeof(Id));
    R* ring_w = reinterpret_cast<R*>(ring);
    Id* ring_prim = reinterpret_cast<Id*>(ring + size_t(qdepth) * kQueueSlots * sizeof(R));
    uint8_t* ring_n = reinterpret_cast<uint8_t*>(ring_prim + size_t(qdepth) * kQueueSlots);
    int q_head = 0, q_count = 0;                   // warp-uniform

    SmemSink ssink{s_acc + threadIdx.x};
    AtomicSink asink{a.grad_atomic};
    SmemAtomicSink msink{s_acc + (threadIdx.x & (a.sink_cols - 1)), a.sink_cols};
    Materials<R, MESH> mat;
    mat.bs = &bs;
    if constexpr (MESH) { mat.mesh = a.mesh; mat.params = a.params; }
    const bool no_bvh = (a.flags & DRTB_FLAG_NO_BVH) != 0;
    TraceCounters cnt;
    uint32_t n_lit = 0;

    // Dynamic distribution.  Warps that own equal shares of the image still finish up to ~20 % apart
    // (the schedulers do not serve resident warps evenly), and an SM whose warps have started to
    // retire issues less: with a static round-robin the last tenth of the kernel ran on a
    // half-empty machine.  So a warp claims the next CHUNK of `chunk_tasks` consecutive tasks from
    // a global counter until none are left.  A chunk's gradient partial is flushed by the warp
    // that ran it (SMALLP), so the sums do not depend on who ran what: results stay bit-reproducible.
    for (;;) {
        unsigned long long claimed = 0;
        if (lane == 0) claimed = atomicAdd(a.task_counter, 1ull);
        const long long chunk = (long long)__shfl_sync(0xffffffffu, claimed, 0);
        if (chunk >= a.n_chunks) break;
        // the first n_big chunks hold chunk_tasks tasks each, the rest a single task: the tail of
        // the kernel is then one task long, not one chunk
        const long long task0 = chunk < a.n_big_chunks ? chunk * a.chunk_tasks
                                                       : a.n_big_chunks * a.chunk_tasks + (chunk - a.n_big_chunks);
        const long long task1 = chunk < a.n_big_chunks ? task0 + a.chunk_tasks : task0 + 1;
        for (long long task = task0; task < task1; ++task) {
            const int sub = spp >= 32 ? 0 : lane / spp;           // pixel within the task
            const int i0 = spp >= 32 ? lane : lane % spp;         // first sample of this lane
            const long long pix = task * ppw + sub;
            const bool lane_ok = sub < ppw && pix < npix;
            int x = 0, y = 0;
            R g0[3] = {R(0), R(0), R(0)};
            if (lane_ok) {
                const int r = int(pix / W);
                x = int(pix - (long long)r * W);
                y = a.shard_count > 1 ? ((r / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + r % a.band_rows
                                      : r;
                if (want_grad) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        g0[c] = R(a.seed_scale * (a.seed_img ? a.seed_img[pix * 3 + c] : 1.0));
                }
            }
            double acc[3] = {0.0, 0.0, 0.0};
            double gacc[3] = {0.0, 0.0, 0.0};          // GEN: this lane's share of the pixel's gradient-image value

            // sweeps over one record; accumulates this lane's share of the pixel and the gradients
            auto sweep = [&](const auto& rec, int n) {
                R L0[3];
                auto run = [&](auto& sink) { radiance_and_adjoint(mat, rec, n, a.min_bounces, inv_p, want_grad, g0, L0, sink); };
                if constexpr (GEN) {
                    if constexpr (SMALLP)             { PixelSink<SmemSink> s{ssink, a.gimg_param, gacc}; run(s); }
                    else if constexpr (kSharedAtomic) { PixelSink<SmemAtomicSink> s{msink, a.gimg_param, gacc}; run(s); }
                    else                              { PixelSink<AtomicSink> s{asink, a.gimg_param, gacc}; run(s); }
                } else {
                    if constexpr (SMALLP)             run(ssink);
                    else if constexpr (kSharedAtomic) run(msink);
                    else                              run(asink);
                }
                acc[0] += double(L0[0]); acc[1] += double(L0[1]); acc[2] += double(L0[2]);       // render.cpp:78
                n_lit += (L0[0] != R(0)) | (L0[1] != R(0)) | (L0[2] != R(0));
            };
            // run the sweeps on the first m queued records, one per lane
            auto drain = [&](int m) {
                __syncwarp();
                if (lane < m) {
                    const int slot = (q_head + lane) & (kQueueSlots - 1);
                    QueueView<R, MESH> qv{ring_w + slot, ring_prim + slot};
                    sweep(qv, ring_n[slot]);
                }
                __syncwarp();
                q_head = (q_head + m) & (kQueueSlots - 1);
                q_count -= m;
            };

            for (int pass = 0; pass < passes; ++pass) {
                const int i = i0 + pass * 32;
                bool lit = false;
                int n = 0;
                PathRecord<R, MESH, QUEUE ? kQueueDepth : kMaxDepth> rec;
                if (lane_ok && i < spp) {
                    const uint64_t key = a.key0 + ((uint64_t)y * W + x) * (uint64_t)spp + (uint64_t)i;
                    const uint64_t base = key * kKeyMul;
                    V3<R> o = {sc.eye[0], sc.eye[1], sc.eye[2]};
                    V3<R> d = camera_ray(sc, x, y, base);
                    n = trace_path<R, MESH, QUEUE ? kQueueDepth : kMaxDepth, GEN>(sc, bs, mat, no_bvh, base, 2u, o, d, a.min_bounces,
                                                                                  a.absorb, a.max_depth, rec, lit, cnt);
                    if (!QUEUE && lit) sweep(rec, n);
                }
                if (QUEUE) {
                    const unsigned m = __ballot_sync(0xffffffffu, lit);
                    if (lit) {
                        const int slot = (q_head + q_count + __popc(m & ((1u << lane) - 1u))) & (kQueueSlots - 1);
                        for (int v = 0; v < n; ++v) {
                            ring_w[v * kQueueSlots + slot] = rec.w_[v];
                            ring_prim[v * kQueueSlots + slot] = rec.prim_[v];
                        }
                        ring_n[slot] = uint8_t(n);
                    }
                    q_count += __popc(m);
                    if (q_count >= 32) drain(32);
                }
            }
            // every queued record belongs to this task's pixel: finish them before the pixel is written
            if (QUEUE && q_count > 0) drain(q_count);

            // pixel_radiance / samples (render.cpp:82): sum the lanes of each pixel
            auto write_pixel = [&](double* dst, double* v, bool mean) {
                if (spp >= 32) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[c] = warp_sum(v[c]);
                    if (lane == 0 && lane_ok) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) dst[pix * 3 + c] = mean ? v[c] / double(spp) : v[c];
                    }
                } else {
                    double tot[3] = {v[0], v[1], v[2]};
                    for (int j = 1; j < spp; ++j) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            double o = __shfl_down_sync(0xffffffffu, v[c], j);
                            if (i0 + j < spp) tot[c] += o;
                        }
                    }
                    if (lane_ok && i0 == 0) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) dst[pix * 3 + c] = mean ? tot[c] / double(spp) : tot[c];
                    }
                }
            };
            if (a.img) write_pixel(a.img, acc, true);
            if (a.n_peer_img > 0) {
                // Image all-gather fused into the render: the pixel goes straight into the full image
                // of every GPU of the job (peer stores over NVLink, lane p -> peer p), at its image
                // row.  ~24 B per pixel and peer against ~10^5 instructions of tracing: the exchange
                // hides completely behind the compute and no gather step follows the kernel.
                const size_t at = ((size_t)y * W + x) * 3;
                if (spp >= 32) {
                    if (!a.img) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) acc[c] = warp_sum(acc[c]);
                    }
                    if (lane < a.n_peer_img && lane_ok) {
                        double* dst = a.peer_img[lane] + at;
#pragma unroll
                        for (int c = 0; c < 3; ++c) dst[c] = acc[c] / double(spp);
                    }
                } else {
                    double tot[3] = {acc[0], acc[1], acc[2]};
                    for (int j = 1; j < spp; ++j) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const double o = __shfl_down_sync(0xffffffffu, acc[c], j);
                            if (i0 + j < spp) tot[c] += o;
                        }
                    }
                    if (lane_ok && i0 == 0)
                        for (int p = 0; p < a.n_peer_img; ++p)
#pragma unroll
                            for (int c = 0; c < 3; ++c) a.peer_img[p][at + c] = tot[c] / double(spp);
                }
            }
            if constexpr (GEN) { if (a.gimg) write_pixel(a.gimg, gacc, false); }
        }
        if (SMALLP && want_grad) {
            // this chunk's gradient: the lanes' columns summed by an xor tree, one row per chunk
            double mine = 0.0;
            for (int j = 0; j < P3; ++j) {
                const double v = warp_sum(s_acc[j * kBlock + threadIdx.x]);
                s_acc[j * kBlock + threadIdx.x] = 0.0;
                if (lane == j) mine = v;
            }
            if (lane < P3) a.grad_partial[(size_t)chunk * P3 + lane] = mine;
        }
    }
    if (kSharedAtomic && want_grad) {
        __syncthreads();                           // every warp's atomics have landed
        for (int j = threadIdx.x; j < P3; j += kBlock) {
            double v = 0.0;
            for (int c = 0; c < a.sink_cols; ++c) v += s_acc[j * a.sink_cols + c];
            a.grad_partial[(size_t)blockIdx.x * P3 + j] = v;
        }
    }
    if (a.stats) {
        // 64-bit warp totals: a warp can see far more than 2^32 node visits on a large mesh
        auto total = [](uint32_t v) {
            unsigned long long t = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            return t;
        };
        const unsigned long long seg = total(cnt.segments), litp = total(n_lit), tr = total(cnt.truncated),
                                 nodes = total(cnt.bvh_nodes), tests = total(cnt.tri_tests);
        if (lane == 0) {
            atomicAdd((unsigned long long*)&a.stats->segments, seg);
            atomicAdd((unsigned long long*)&a.stats->lit_paths, litp);
            if (tr) atomicAdd((unsigned long long*)&a.stats->truncated_paths, tr);
            if (MESH) {
                atomicAdd((unsigned long long*)&a.stats->bvh_nodes, nodes);
                atomicAdd((unsigned long long*)&a.stats->tri_tests, tests);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// The pixel loop for Russian-roulette renders (absorb < 1) of all-diffuse analytic scenes, with
// PATH REGENERATION.  Path lengths are geometric there (mean 1.9 segments at the reference's
// defaults -b 1 -p 0.5, the longest of 32 about 6.5), so a warp that traces 32 samples to the
// end keeps a third of its lanes busy.  Here a warp owns a CHUNK of pixels (<= kRegenPixels,
// about 1024 samples) as one flat list of samples; every lane steps ONE segment per iteration
// (trace_segment), and as soon as kRefillLanes lanes are free they take the next samples of the
// list (ballot order, hence deterministic), whichever pixel those belong to -- only the last
// iterations of a chunk run on thinning lanes.  Lit records go through the per-warp ring (in
// global memory, kQueueDepth deep) tagged with their pixel; the sweeps add a pixel's radiance
// into the warp's shared accumulators in ring order (match_any groups, rank by rank), so the
// image is bit-reproducible.  Gradients, chunk distribution and reduction as in render_kernel.
// ---------------------------------------------------------------------------
constexpr int kRegenPixels = 64;
#ifndef DRTB_REFILL_LANES
#define DRTB_REFILL_LANES 8
#endif
__host__ __device__ constexpr size_t regen_smem_per_warp() { return size_t(kRegenPixels) * (3 * sizeof(double) + sizeof(int2)); }
__host__ __device__ constexpr size_t regen_ring_per_warp(size_t real_size) { return queue_bytes_per_warp(kQueueDepth, real_size, 1) + kQueueSlots; }

template <typename R, bool SMALLP>
__global__ void __launch_bounds__(kBlock, sizeof(R) == 4 ? DRTB_MIN_BLOCKS_F32 : DRTB_MIN_BLOCKS)
render_regen_kernel(const __grid_constant__ DevScene<R> sc, const __grid_constant__ RenderArgs a)
{
    extern __shared__ double s_dyn[];              // [acc][per warp: pxacc[kRegenPixels][3] | pxy[kRegenPixels]]
    __shared__ BlockScene<R> bs;
    const bool want_grad = (a.flags & DRTB_FLAG_GRAD) != 0;
    const int P3 = sc.n_params * 3;
    double* s_acc = s_dyn;
    const int acc_doubles = !want_grad ? 0 : SMALLP ? P3 * kBlock : P3 * a.sink_cols;
    load_block_scene(bs, sc, a.params);
    for (int i = threadIdx.x; i < acc_doubles; i += kBlock) s_acc[i] = 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int W = sc.width, spp = a.spp;
    const long long npix = (long long)a.shard_rows * W;
    const R inv_p = R(1.0 / (1.0 - a.absorb));
    double* pxacc = s_dyn + acc_doubles + size_t(warp) * (regen_smem_per_warp() / sizeof(double));
    int2* pxy = reinterpret_cast<int2*>(pxacc + kRegenPixels * 3);
    unsigned char* ring = a.ring_scratch + (size_t(blockIdx.x) * kWarpsPerBlock + warp) * regen_ring_per_warp(sizeof(R));
    R* ring_w = reinterpret_cast<R*>(ring);
    uint8_t* ring_prim = ring + size_t(kQueueDepth) * kQueueSlots * sizeof(R);
    uint8_t* ring_n = ring_prim + size_t(kQueueDepth) * kQueueSlots;
    uint8_t* ring_px = ring_n + kQueueSlots;
    int q_head = 0, q_count = 0;                   // warp-uniform

    SmemSink ssink{s_acc + threadIdx.x};
    SmemAtomicSink msink{s_acc + (threadIdx.x & (a.sink_cols - 1)), a.sink_cols};
    Materials<R, false> mat;
    mat.bs = &bs;
    TraceCounters cnt;
    uint32_t n_lit = 0;

    for (;;) {
        unsigned long long claimed = 0;
        if (lane == 0) claimed = atomicAdd(a.task_counter, 1ull);
        const long long chunk = (long long)__shfl_sync(0xffffffffu, claimed, 0);
        if (chunk >= a.n_chunks) break;
        // big chunks of chunk_tasks pixels, then a last round of small ones (render_kernel's tail rule)
        const long long big_end = a.n_big_chunks * a.chunk_tasks;
        const long long pix0 = chunk < a.n_big_chunks ? chunk * a.chunk_tasks : big_end + (chunk - a.n_big_chunks) * a.small_chunk;
        const long long pix1 = min(npix, pix0 + (chunk < a.n_big_chunks ? a.chunk_tasks : a.small_chunk));
        const int K = int(pix1 - pix0);
        for (int k = lane; k < K; k += 32) {        // this chunk's pixels: image coordinates, cleared sums
            const long long pix = pix0 + k;
            const int r = int(pix / W);
            const int x = int(pix - (long long)r * W);
            const int y = a.shard_count > 1 ? ((r / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + r % a.band_rows : r;
            pxy[k] = make_int2(x, y);
            pxacc[3 * k] = 0.0; pxacc[3 * k + 1] = 0.0; pxacc[3 * k + 2] = 0.0;
        }
        __syncwarp();

        // a pixel's radiance: lanes holding the same pixel add one after the other, in lane order
        auto add_to_pixels = [&](int px, const R* L0) {       // px < 0: nothing to add
            const unsigned grp = __match_any_sync(0xffffffffu, px);
            const int rank = __popc(grp & lt_mask);
            const int rounds = __reduce_max_sync(0xffffffffu, px < 0 ? 0 : __popc(grp));
            for (int r = 0; r < rounds; ++r) {
                if (px >= 0 && rank == r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) pxacc[3 * px + c] += double(L0[c]);
                }
                __syncwarp();
            }
        };
        // both sweeps over one record of pixel px (render.cpp:78-80)
        auto sweep = [&](const auto& rec, int n, int px, R* L0) {
            R g0[3] = {R(0), R(0), R(0)};
            if (want_grad) {
#pragma unroll
                for (int c = 0; c < 3; ++c) g0[c] = R(a.seed_scale * (a.seed_img ? a.seed_img[(pix0 + px) * 3 + c] : 1.0));
            }
            if constexpr (SMALLP) radiance_and_adjoint(mat, rec, n, a.min_bounces, inv_p, want_grad, g0, L0, ssink);
            else                  radiance_and_adjoint(mat, rec, n, a.min_bounces, inv_p, want_grad, g0, L0, msink);
            n_lit += (L0[0] != R(0)) | (L0[1] != R(0)) | (L0[2] != R(0));
        };
        auto drain = [&](int m) {
            __syncwarp();
            R L0[3] = {R(0), R(0), R(0)};
            int px = -1;
            if (lane < m) {
                const int slot = (q_head + lane) & (kQueueSlots - 1);
                QueueView<R, false> qv{ring_w + slot, ring_prim + slot};
                px = ring_px[slot];
                sweep(qv, ring_n[slot], px, L0);
            }
            add_to_pixels(px, L0);
            q_head = (q_head + m) & (kQueueSlots - 1);
            q_count -= m;
        };

        const int n_samples = K * spp;                        // <= kRegenPixels * spp
        int next_s = 0;                                       // warp-uniform: next unassigned sample of the chunk
        bool alive = false, lit = false;
        int depth = 0, n = 0, my_px = 0;
        uint64_t ctr = 0;
        V3<R> o = {R(0), R(0), R(0)}, d = o;
        PathRecord<R, false, kMaxDepth> rec;
        for (;;) {
            const unsigned dead = __ballot_sync(0xffffffffu, !alive);
            if (next_s < n_samples && (__popc(dead) >= DRTB_REFILL_LANES || dead == 0xffffffffu)) {
                const int mine = next_s + __popc(dead & lt_mask);
                if (!alive && mine < n_samples) {
                    my_px = mine / spp;
                    const int i = mine - my_px * spp;
                    const int2 xy = pxy[my_px];
                    const uint64_t key = a.key0 + ((uint64_t)xy.y * W + xy.x) * (uint64_t)spp + (uint64_t)i;
                    const uint64_t base = key * kKeyMul;
                    o = {sc.eye[0], sc.eye[1], sc.eye[2]};
                    d = camera_ray(sc, xy.x, xy.y, base);
                    ctr = base + kGolden + 2u;
                    depth = 0; n = 0; lit = false;
                    alive = !roulette_absorbs(ctr, 0, a.min_bounces, a.absorb);     // min_bounces == 0: trace() may return 0 at once
                }
                next_s = min(n_samples, next_s + __popc(dead));
            }
            if (__ballot_sync(0xffffffffu, alive) == 0u) {
                if (next_s >= n_samples) break;
                continue;                                      // every fresh sample was absorbed at once (min_bounces == 0)
            }
            bool done = false;
            if (alive) done = trace_segment(sc, bs, mat, ctr, o, d, depth, n, lit, a.min_bounces, a.absorb, a.max_depth, rec, cnt);
            if (done) alive = false;
            // a record too deep for the ring (p ~ 1e-5 at the reference's defaults) is swept by its own lane
            const bool deep = done && lit && n > kQueueDepth;
            if (__any_sync(0xffffffffu, deep)) {
                R L0[3] = {R(0), R(0), R(0)};
                if (deep) sweep(rec, n, my_px, L0);
                add_to_pixels(deep ? my_px : -1, L0);
            }
            const bool queued = done && lit && n <= kQueueDepth;
            const unsigned m = __ballot_sync(0xffffffffu, queued);
            if (queued) {
                const int slot = (q_head + q_count + __popc(m & lt_mask)) & (kQueueSlots - 1);
                for (int v = 0; v < n; ++v) {
                    ring_w[v * kQueueSlots + slot] = rec.w_[v];
                    ring_prim[v * kQueueSlots + slot] = rec.prim_[v];
                }
                ring_n[slot] = uint8_t(n);
                ring_px[slot] = uint8_t(my_px);
            }
            q_count += __popc(m);
            if (q_count >= 32) drain(32);
        }
        if (q_count > 0) drain(q_count);
        __syncwarp();

        // pixel_radiance / samples (render.cpp:82), compact shard image and/or every peer's full image
        for (int k = lane; k < K; k += 32) {
            const int2 xy = pxy[k];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double v = pxacc[3 * k + c] / double(spp);
                if (a.img) a.img[(pix0 + k) * 3 + c] = v;
                for (int p = 0; p < a.n_peer_img; ++p) a.peer_img[p][((size_t)xy.y * W + xy.x) * 3 + c] = v;
            }
        }
        __syncwarp();
        if (SMALLP && want_grad) {                            // this chunk's gradient row (see render_kernel)
            double mine = 0.0;
            for (int j = 0; j < P3; ++j) {
                const double v = warp_sum(s_acc[j * kBlock + threadIdx.x]);
                s_acc[j * kBlock + threadIdx.x] = 0.0;
                if (lane == j) mine = v;
            }
            if (lane < P3) a.grad_partial[(size_t)chunk * P3 + lane] = mine;
        }
    }
    if (!SMALLP && want_grad) {
        __syncthreads();                           // every warp's atomics have landed
        for (int j = threadIdx.x; j < P3; j += kBlock) {
            double v = 0.0;
            for (int c = 0; c < a.sink_cols; ++c) v += s_acc[j * a.sink_cols + c];
            a.grad_partial[(size_t)blockIdx.x * P3 + j] = v;
        }
    }
    if (a.stats) {
        auto total = [](uint32_t v) {
            unsigned long long t = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            return t;
        };
        const unsigned long long seg = total(cnt.segments), litp = total(n_lit), tr = total(cnt.truncated);
        if (lane == 0) {
            atomicAdd((unsigned long long*)&a.stats->segments, seg);
            atomicAdd((unsigned long long*)&a.stats->lit_paths, litp);
            if (tr) atomicAdd((unsigned long long*)&a.stats->truncated_paths, tr);
        }
    }
}

__global__ void iota_kernel(int* __restrict__ v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

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

// Wavefront stage 4: radiance recurrence + adjoint over the records one batch
// left in HBM, the per-pixel sums of src/render.cpp:78-82 and the gradient sums.
// Same warp-task shape as render_kernel (a warp owns whole pixels, lanes own
// samples), so the image is summed in a fixed order.
template <typename R, bool SMALLP, int CAP>
__global__ void __launch_bounds__(kBlock)
wf_adjoint(const __grid_constant__ DevScene<R> sc, const __grid_constant__ WfArgs a, const WfBuffers<R> b, int partial_row0)
{
    extern __shared__ double s_dyn[];
    __shared__ BlockScene<R> bs;
    __shared__ double s_red[kSmallP * 3][kWarpsPerBlock];
    const bool want_grad = (a.flags & DRTB_FLAG_GRAD) != 0;
    const int P3 = sc.n_params * 3;
    double* s_acc = s_dyn;
    const int acc_doubles = (SMALLP && want_grad) ? P3 * kBlock : 0;
    load_block_scene(bs, sc, a.params);
    for (int i = threadIdx.x; i < acc_doubles; i += kBlock) s_acc[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int spp = a.spp;
    const long long npix = a.n_paths / spp, pix0 = a.first_path / spp;
    const int ppw = spp >= 32 ? 1 : 32 / spp;
    const int passes = spp >= 32 ? (spp + 31) / 32 : 1;
    const long long n_tasks = (npix + ppw - 1) / ppw;
    const long long n_warps = (long long)gridDim.x * kWarpsPerBlock;
    const R inv_p = a.absorb < 1.0 ? R(1.0 / (1.0 - a.absorb)) : R(0);
    SmemSink ssink{s_acc + threadIdx.x};
    AtomicSink asink{a.grad_atomic};
    Materials<R, true> mat;
    mat.bs = &bs; mat.mesh = a.mesh; mat.params = a.params;
    uint32_t n_lit = 0;
    for (long long task = (long long)blockIdx.x * kWarpsPerBlock + warp; task < n_tasks; task += n_warps) {
        const int sub = spp >= 32 ? 0 : lane / spp;
        const int i0 = spp >= 32 ? lane : lane % spp;
        const long long lp = task * ppw + sub;                // pixel within the batch
        const bool lane_ok = sub < ppw && lp < npix;
        const long long pix = pix0 + lp;                      // pixel within the shard
        R g0[3] = {R(0), R(0), R(0)};
        if (lane_ok && want_grad) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g0[c] = R(a.seed_scale * (a.seed_img ? a.seed_img[pix * 3 + c] : 1.0));
        }
        double acc[3] = {0.0, 0.0, 0.0};
        double gacc[3] = {0.0, 0.0, 0.0};                     // gradient image (gimg_param == -1: stays zero)
        PixelSink<SmemSink> ps{ssink, a.gimg_param, gacc};
        PixelSink<AtomicSink> pa{asink, a.gimg_param, gacc};
        for (int pass = 0; pass < passes; ++pass) {
            const int i = i0 + pass * 32;
            if (!(lane_ok && i < spp)) continue;
            const long long p = lp * spp + i;
            const uint32_t st = b.state[p];
            if (!(st & kStLit)) continue;
            const int n = int((st >> 16) & 0xffu);
            const WfRecordView<R, CAP> rec{b.rec_w + p, b.rec_prim + p, a.batch};
            R L0[3];
            if (SMALLP) radiance_and_adjoint(mat, rec, n, a.min_bounces, inv_p, want_grad, g0, L0, ps);
            else        radiance_and_adjoint(mat, rec, n, a.min_bounces, inv_p, want_grad, g0, L0, pa);
            acc[0] += double(L0[0]); acc[1] += double(L0[1]); acc[2] += double(L0[2]);
            n_lit += (L0[0] != R(0)) | (L0[1] != R(0)) | (L0[2] != R(0));
        }
        auto write_pixel = [&](double* dst, double* v, bool mean) {
            if (spp >= 32) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = warp_sum(v[c]);
                if (lane == 0 && lane_ok) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) dst[pix * 3 + c] = mean ? v[c] / double(spp) : v[c];
                }
            } else {
                double tot[3] = {v[0], v[1], v[2]};
                for (int j = 1; j < spp; ++j) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        double o = __shfl_down_sync(0xffffffffu, v[c], j);
                        if (i0 + j < spp) tot[c] += o;
                    }
                }
                if (lane_ok && i0 == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) dst[pix * 3 + c] = mean ? tot[c] / double(spp) : tot[c];
                }
            }
        };
        if (a.img) write_pixel(a.img, acc, true);
        if (a.gimg) write_pixel(a.gimg, gacc, false);
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
struct drtb_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    bool has_scene = false;
    bool has_specular = false;    // some primitive carries a DRTB_SPECULAR material
    std::vector<drtb_prim> prims;
    std::vector<drtb_material> materials;
    std::vector<double> params;
    drtb_camera camera{};
    DevScene<double> sc64{};
    DevScene<float> sc32{};
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
    float4* d_nodes = nullptr;
    double* d_tri64 = nullptr;
    float4* d_tri32 = nullptr;
    int32_t* d_tri_color = nullptr;
    int32_t* d_tri_emis = nullptr;
    double mesh_build_ms = 0.0;
    int mesh_nodes = 0;
    // wavefront buffers (mesh scenes), grown on demand
    void* wf_mem = nullptr;       size_t wf_cap = 0;
    bool mesh_megakernel = false; // DRTB_MESH_PIPELINE=megakernel: trace meshes inside render_kernel (A/B aid)
    unsigned long long* d_task_counter = nullptr;
    double* img_peers[kMaxPeers] = {};   // drtb_set_image_peers: full images the render kernel fills directly
    int n_img_peers = 0;
};

namespace {

thread_local std::string g_create_err;

int fail(drtb_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return code;
}

#define CK(ctx, call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, DRTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
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
            d.aa_c[ax][kAxisFast - n + j] = R(c.prims[i].v[ax] * c.prims[i].v[3]);   // s * off, exact (s = +-1)
            d.aa_id[ax][kAxisFast - n + j] = i;
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
}

int shard_rows_impl(int H, int idx, int cnt, int band)
{
    if (H <= 0) return 0;
    if (cnt <= 1) return H;
    if (band < 1) band = 1;
    int rows = 0;
    const int nb = (H + band - 1) / band;
    for (int b = idx; b < nb; b += cnt) rows += std::min(band, H - b * band);
    return rows;
}

struct Plan {
    RenderArgs a;
    int P3;
    bool smallp;
    int grid;
    size_t smem;
    uint64_t paths;
};

// Resident blocks per SM of one render_kernel instantiation at `smem` dynamic bytes.
template <typename K>
int occupancy(drtb_ctx* ctx, K kernel, size_t smem, int& out)
{
    CK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int nb = 0;
    CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kBlock, smem));
    if (nb < 1) return fail(ctx, DRTB_ERR_CUDA, "render kernel does not fit on an SM");
    out = nb;
    return DRTB_OK;
}

// How a render's units of work are cut into the chunks that warps claim from the global counter
// (render_kernel: units = warp tasks; render_regen_kernel: units = pixels).  The first n_big chunks hold
// `big` units each, the rest `small` units each (the last one possibly fewer): big chunks keep the claim and
// the per-chunk gradient row cheap, the small ones of the last round keep the tail of the kernel short.
// Pure host arithmetic, exported as drtb_chunk_plan so that the CPU tests can check that every unit is
// covered exactly once for any size.
struct ChunkPlan { long long big, small, n_big, n_chunks; };

ChunkPlan plan_chunks(long long n_units, int spp, long long resident_warps, bool regen, long long forced_big)
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
constexpr int kReduceDirectRows = 4096, kReduceBlockRows = 1024;
inline size_t reduce_scratch_rows(size_t rows) { return rows > kReduceDirectRows ? (rows + kReduceBlockRows - 1) / kReduceBlockRows : 0; }
int reduce_partials(drtb_ctx* ctx, double* partial, size_t rows, int P3, double* d_grad, cudaStream_t stream)
{
    if (P3 <= 0 || rows == 0) return DRTB_OK;
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

template <typename R, bool SMALLP, int QUEUE, bool MESH, bool GEN>
int launch_variant(drtb_ctx* ctx, const DevScene<R>& sc, RenderArgs& a, size_t smem, long long n_tasks,
                   int P3, bool want_grad, cudaStream_t stream, size_t& rows_out)
{
    int per_sm = 0;
    int rc = occupancy(ctx, render_kernel<R, SMALLP, QUEUE, MESH, GEN>, smem, per_sm);
    if (rc != DRTB_OK) return rc;
    const long long need_blocks = (n_tasks + kWarpsPerBlock - 1) / kWarpsPerBlock;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > need_blocks) grid = need_blocks;
    if (grid < 1) grid = 1;
    long long forced = 0;
    if (const char* e = std::getenv("DRTB_CHUNK_TASKS")) forced = std::max(1, std::atoi(e));     // A/B aid
    const ChunkPlan plan = plan_chunks(n_tasks, a.spp, grid * kWarpsPerBlock, false, forced);
    a.chunk_tasks = int(plan.big);
    a.small_chunk = 1;
    a.n_big_chunks = plan.n_big;
    const long long n_chunks = plan.n_chunks;
    a.n_chunks = n_chunks;
    if (!ctx->d_task_counter) CK(ctx, cudaMalloc(&ctx->d_task_counter, sizeof(unsigned long long)));
    CK(ctx, cudaMemsetAsync(ctx->d_task_counter, 0, sizeof(unsigned long long), stream));
    a.task_counter = ctx->d_task_counter;
    // gradient partials: one row per chunk (SMALLP, summed in chunk order whoever ran the chunk) or
    // per block (shared atomic columns)
    size_t rows = 0;
    if (want_grad && SMALLP) rows = size_t(n_chunks);
    else if (want_grad && !MESH) rows = size_t(grid);
    if (rows) {
        rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (rows + reduce_scratch_rows(rows)) * P3);
        if (rc != DRTB_OK) return rc;
        a.grad_partial = ctx->d_partial;
    }
    if (QUEUE == 2) {
        const size_t per_warp = queue_bytes_per_warp(a.max_depth, sizeof(R), MESH ? sizeof(int32_t) : sizeof(uint8_t));
        rc = ensure(ctx, ctx->d_ring, ctx->ring_cap, size_t(grid) * kWarpsPerBlock * per_warp / sizeof(double));
        if (rc != DRTB_OK) return rc;
        a.ring_scratch = reinterpret_cast<unsigned char*>(ctx->d_ring);
    }
    render_kernel<R, SMALLP, QUEUE, MESH, GEN><<<int(grid), kBlock, smem, stream>>>(sc, a);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    rows_out = rows;
    return DRTB_OK;
}

// Russian-roulette renders of all-diffuse analytic scenes: render_regen_kernel.
template <typename R, bool SMALLP>
int launch_regen(drtb_ctx* ctx, const DevScene<R>& sc, RenderArgs& a, long long npix, int P3, bool want_grad,
                 cudaStream_t stream, size_t& rows_out)
{
    size_t smem = kWarpsPerBlock * regen_smem_per_warp();
    if (want_grad) smem += SMALLP ? size_t(P3) * kBlock * sizeof(double) : size_t(P3) * a.sink_cols * sizeof(double);
    int per_sm = 0;
    CK(ctx, cudaFuncSetAttribute(render_regen_kernel<R, SMALLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_regen_kernel<R, SMALLP>, kBlock, smem));
    if (per_sm < 1) return fail(ctx, DRTB_ERR_CUDA, "regenerating render kernel does not fit on an SM");
    const ChunkPlan plan = plan_chunks(npix, a.spp, (long long)ctx->sm_count * per_sm * kWarpsPerBlock, true, 0);
    a.chunk_tasks = int(plan.big);
    a.small_chunk = int(plan.small);
    a.n_big_chunks = plan.n_big;
    a.n_chunks = plan.n_chunks;
    long long grid = std::min<long long>((long long)ctx->sm_count * per_sm, (a.n_chunks + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (grid < 1) grid = 1;
    if (!ctx->d_task_counter) CK(ctx, cudaMalloc(&ctx->d_task_counter, sizeof(unsigned long long)));
    CK(ctx, cudaMemsetAsync(ctx->d_task_counter, 0, sizeof(unsigned long long), stream));
    a.task_counter = ctx->d_task_counter;
    int rc = ensure(ctx, ctx->d_ring, ctx->ring_cap, size_t(grid) * kWarpsPerBlock * regen_ring_per_warp(sizeof(R)) / sizeof(double));
    if (rc != DRTB_OK) return rc;
    a.ring_scratch = reinterpret_cast<unsigned char*>(ctx->d_ring);
    const size_t rows = !want_grad ? 0 : SMALLP ? size_t(a.n_chunks) : size_t(grid);
    if (rows) {
        rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (rows + reduce_scratch_rows(rows)) * P3);
        if (rc != DRTB_OK) return rc;
        a.grad_partial = ctx->d_partial;
    }
    render_regen_kernel<R, SMALLP><<<int(grid), kBlock, smem, stream>>>(sc, a);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    rows_out = rows;
    return DRTB_OK;
}

template <typename R>
int launch_precision(drtb_ctx* ctx, const DevScene<R>& sc, RenderArgs& a, bool smallp, int queue, bool mesh, bool gen,
                     size_t smem, long long n_tasks, int P3, bool want_grad, cudaStream_t stream, size_t& rows)
{
#define DRTB_LAUNCH(SP, Q, M) launch_variant<R, SP, Q, M, false>(ctx, sc, a, smem, n_tasks, P3, want_grad, stream, rows)
#define DRTB_LAUNCH_GEN(SP, Q) launch_variant<R, SP, Q, false, true>(ctx, sc, a, smem, n_tasks, P3, want_grad, stream, rows)
#define DRTB_BY_QUEUE(L, SP, ...) (queue == 2 ? L(SP, 2, ##__VA_ARGS__) : queue == 1 ? L(SP, 1, ##__VA_ARGS__) : L(SP, 0, ##__VA_ARGS__))
    if (gen) {
        // SpecularBxDF materials and/or a gradient image (analytic scenes; mesh scenes take the wavefront)
        if (mesh) return fail(ctx, DRTB_ERR_UNSUPPORTED, "specular materials / gradient images on a mesh scene need the wavefront pipeline");
        return smallp ? DRTB_BY_QUEUE(DRTB_LAUNCH_GEN, true) : DRTB_BY_QUEUE(DRTB_LAUNCH_GEN, false);
    }
    if (mesh) {
        // mesh scenes: parameters live in global memory; small sets still use the smem gradient columns
        // (the megakernel on a mesh is an A/B aid: shared ring only)
        if (smallp) return queue ? DRTB_LAUNCH(true, 1, true) : DRTB_LAUNCH(true, 0, true);
        return queue ? DRTB_LAUNCH(false, 1, true) : DRTB_LAUNCH(false, 0, true);
    }
    return smallp ? DRTB_BY_QUEUE(DRTB_LAUNCH, true, false) : DRTB_BY_QUEUE(DRTB_LAUNCH, false, false);
#undef DRTB_BY_QUEUE
#undef DRTB_LAUNCH
#undef DRTB_LAUNCH_GEN
}

MeshView mesh_view(const drtb_ctx* ctx)
{
    MeshView m{};
    m.nodes = ctx->d_nodes; m.tri64 = ctx->d_tri64; m.tri32 = ctx->d_tri32;
    m.color = ctx->d_tri_color; m.emis = ctx->d_tri_emis;
    m.n_tris = int32_t(ctx->n_tris); m.n_prims = int32_t(ctx->prims.size());
    return m;
}

void free_mesh(drtb_ctx* ctx)
{
    cudaFree(ctx->d_nodes); cudaFree(ctx->d_tri64); cudaFree(ctx->d_tri32);
    cudaFree(ctx->d_tri_color); cudaFree(ctx->d_tri_emis);
    ctx->d_nodes = nullptr; ctx->d_tri64 = nullptr; ctx->d_tri32 = nullptr;
    ctx->d_tri_color = nullptr; ctx->d_tri_emis = nullptr;
    ctx->n_tris = 0;
}

int validate_opts(drtb_ctx* ctx, const drtb_render_opts* o)
{
    if (!ctx->has_scene) return fail(ctx, DRTB_ERR_INVALID, "no scene uploaded");
    if (!o) return fail(ctx, DRTB_ERR_INVALID, "opts is NULL");
    if (o->spp < 1) return fail(ctx, DRTB_ERR_INVALID, "spp must be >= 1");
    if (o->min_bounces < 0) return fail(ctx, DRTB_ERR_INVALID, "min_bounces must be >= 0");
    if (!(o->absorb >= 0.0 && o->absorb <= 1.0)) return fail(ctx, DRTB_ERR_INVALID, "absorb must be in [0, 1]");
    if (o->precision == DRTB_MIXED) return fail(ctx, DRTB_ERR_UNSUPPORTED, "DRTB_MIXED is not implemented in this build");
    if (o->precision != DRTB_F64 && o->precision != DRTB_F32) return fail(ctx, DRTB_ERR_INVALID, "unknown precision");
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

int effective_max_depth(const drtb_render_opts* o)
{
    if (o->max_depth > 0) return o->max_depth;
    return o->absorb == 1.0 ? std::max(1, o->min_bounces) : kMaxDepth;
}

// Optional per-pixel gradient image of one parameter (drtb_render_grad_image).
struct GradImage {
    int32_t param = -1;
    double* d_out = nullptr;             // shard_rows x W x 3 (device)
};

template <typename R>
int launch_wavefront(drtb_ctx* ctx, const DevScene<R>& sc, const drtb_render_opts* o, const double* d_seed, double* d_img,
                     double* d_grad, drtb_stats* d_stats, const GradImage& gi, cudaStream_t stream);

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
    if (ctx->n_tris > 0 && !ctx->mesh_megakernel)
        return o->precision == DRTB_F32 ? launch_wavefront<float>(ctx, ctx->sc32, o, d_seed, d_img, d_grad, d_stats, gi, stream)
                                        : launch_wavefront<double>(ctx, ctx->sc64, o, d_seed, d_img, d_grad, d_stats, gi, stream);
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
    const bool gen = ctx->has_specular || want_gimg;

    const bool smallp = P <= kSmallP;
    const bool f32 = o->precision == DRTB_F32;
    // lit-path compaction needs whole-pixel warp tasks and records that fit the ring
    const bool queue = o->spp >= 32 && a.max_depth <= kQueueDepth;
    const bool mesh = ctx->n_tris > 0;
    a.mesh = mesh_view(ctx);
    size_t smem = (smallp && want_grad) ? size_t(P3) * kBlock * sizeof(double) : 0;
    const size_t ring_bytes = queue ? kWarpsPerBlock * queue_bytes_per_warp(a.max_depth, f32 ? sizeof(float) : sizeof(double),
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
        const int want_blocks = f32 ? DRTB_MIN_BLOCKS_F32 : DRTB_MIN_BLOCKS;
        const size_t static_smem = (f32 ? sizeof(BlockScene<float>) : sizeof(BlockScene<double>)) + 1024;   // + 1 KB the system reserves per block
        if ((static_smem + smem + ring_bytes) * want_blocks > size_t(228) * 1024) queue_kind = 2;
    }
    // Measured and NOT adopted for records that fit (B <= 8, double): the global ring at every depth is 0.9 % faster
    // (42.51 against 42.88 ms: the freed shared memory goes to the L1 that holds the local-memory records, hit rate
    // 44 -> 81 %) but the L2 writes the constantly rewritten rings back to HBM, 0.9 GB per render against 0.15 GB --
    // a per-launch persisting access-policy window over the rings did not change that -- and keeping the lanes' own
    // records in the freed shared memory instead of local memory was slower (43.71 ms).  DRTB_RING forces either.
    if (queue_kind == 1 && !mesh && ctx->ring_policy == 2) queue_kind = 2;
    if (queue_kind < 2) smem += ring_bytes;
    smem = (smem + 15) & ~size_t(15);
    const long long npix = (long long)rows * W;
    const int ppw = o->spp >= 32 ? 1 : 32 / o->spp;
    const long long n_tasks = (npix + ppw - 1) / ppw;

    if (a.stats) {
        CK(ctx, cudaMemsetAsync(d_stats, 0, sizeof(drtb_stats), stream));
    }
    if (want_grad && !smallp && !shared_atomic) {
        CK(ctx, cudaMemsetAsync(d_grad, 0, sizeof(double) * P3, stream));
        a.grad_atomic = d_grad;
    }
    size_t partial_rows = 0;
    int rc;
    // Russian roulette on an all-diffuse analytic scene: the path-regenerating kernel
    const bool regen = o->absorb < 1.0 && !mesh && !gen && !ctx->no_regen;
    if (regen) {
        rc = f32 ? (smallp ? launch_regen<float, true>(ctx, ctx->sc32, a, npix, P3, want_grad, stream, partial_rows)
                           : launch_regen<float, false>(ctx, ctx->sc32, a, npix, P3, want_grad, stream, partial_rows))
                 : (smallp ? launch_regen<double, true>(ctx, ctx->sc64, a, npix, P3, want_grad, stream, partial_rows)
                           : launch_regen<double, false>(ctx, ctx->sc64, a, npix, P3, want_grad, stream, partial_rows));
    } else {
        rc = f32 ? launch_precision<float>(ctx, ctx->sc32, a, smallp, queue_kind, mesh, gen, smem, n_tasks, P3, want_grad, stream, partial_rows)
                 : launch_precision<double>(ctx, ctx->sc64, a, smallp, queue_kind, mesh, gen, smem, n_tasks, P3, want_grad, stream, partial_rows);
    }
    if (rc != DRTB_OK) return rc;
    if (want_grad && (smallp || shared_atomic)) return reduce_partials(ctx, ctx->d_partial, partial_rows, P3, d_grad, stream);
    return DRTB_OK;
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
    if (a.stats) CK(ctx, cudaMemsetAsync(d_stats, 0, sizeof(drtb_stats), stream));
    if (want_grad && !smallp) {
        CK(ctx, cudaMemsetAsync(d_grad, 0, sizeof(double) * P3, stream));
        a.grad_atomic = d_grad;
    }
    // grids
    int trav_per_sm = 0;
    CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&trav_per_sm, wf_traverse<R>, 128, 0));
    const int trav_grid = ctx->sm_count * std::max(1, trav_per_sm);
    const size_t adj_smem = (smallp && want_grad) ? size_t(P3) * kBlock * sizeof(double) : 0;
    const int adj_grid = ctx->sm_count * 8;
    if (want_grad && smallp) {
        int rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, size_t(adj_grid) * n_batches * P3);
        if (rc != DRTB_OK) return rc;
        a.grad_partial = ctx->d_partial;
    }
    const bool no_bvh = (o->flags & DRTB_FLAG_NO_BVH) != 0;
    const bool deep = D > kQueueDepth;
    for (int bi = 0; bi < n_batches; ++bi) {
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
            if (deep) { CK(ctx, cudaFuncSetAttribute(wf_adjoint<R, true, kMaxDepth>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(adj_smem)));
                        wf_adjoint<R, true, kMaxDepth><<<adj_grid, kBlock, adj_smem, stream>>>(sc, a, b, bi * adj_grid); }
            else      { CK(ctx, cudaFuncSetAttribute(wf_adjoint<R, true, kQueueDepth>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(adj_smem)));
                        wf_adjoint<R, true, kQueueDepth><<<adj_grid, kBlock, adj_smem, stream>>>(sc, a, b, bi * adj_grid); }
        } else {
            if (deep) wf_adjoint<R, false, kMaxDepth><<<adj_grid, kBlock, 0, stream>>>(sc, a, b, 0);
            else      wf_adjoint<R, false, kQueueDepth><<<adj_grid, kBlock, 0, stream>>>(sc, a, b, 0);
        }
        CK(ctx, cudaGetLastError());
        ctx->launches += 2 + 2 * D;
    }
    if (want_grad && smallp) {
        reduce_grad_kernel<<<1, 256, 0, stream>>>(ctx->d_partial, adj_grid * n_batches, adj_grid * n_batches, P3, d_grad);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return DRTB_OK;
}

// Decorrelated adjoint (the reference's integrate_unbiased idea, integrate.hpp:39-52:
// fresh samples for the backward pass): the image comes from stream `seed`, the
// gradients from stream `adjoint_seed`.  With a counter-based RNG that is simply
// a second, gradient-only launch keyed differently.
int launch_render(drtb_ctx* ctx, const drtb_render_opts* o, const double* d_seed, double* d_img,
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
    if (const char* e = std::getenv("DRTB_MESH_PIPELINE")) ctx->mesh_megakernel = std::string(e) == "megakernel";
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
    {   // (sin, cos)(2 pi i 2^23 / M) for Real<double>::sincos_tab, in long double on the host
        static double2 tab[drtb::kSinCosEntries];
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int i = 0; i < drtb::kSinCosEntries; ++i) {
            const long double ang = two_pi * ((long double)i * (long double)(1u << drtb::kSinCosShift)) / 2147483647.0L;
            tab[i] = make_double2(double(sinl(ang)), double(cosl(ang)));
        }
        if (cudaMemcpyToSymbol(drtb::g_sincos_tab, tab, sizeof(tab)) != cudaSuccess) {
            std::string m = cudaGetErrorString(cudaGetLastError());
            drtb_destroy(ctx);
            return fail(nullptr, DRTB_ERR_CUDA, "context setup failed: " + m);
        }
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
    float4 *d_lo = nullptr, *d_hi = nullptr, *d_blo = nullptr, *d_bhi = nullptr, *d_wide = nullptr; uint32_t* d_bounds = nullptr;
    uint64_t *d_keys = nullptr, *d_keys2 = nullptr, *d_flags = nullptr, *d_scan = nullptr;
    uint32_t *d_vals = nullptr, *d_vals2 = nullptr;
    int2 *d_children = nullptr, *d_tasks = nullptr, *d_tasks2 = nullptr;
    int *d_parent = nullptr, *d_arrive = nullptr, *d_clusters = nullptr, *d_clusters2 = nullptr, *d_nearest = nullptr;
    int32_t* d_leaf_order = nullptr; CollapseCounters* d_cnt = nullptr; void *d_tmp = nullptr, *d_tmp2 = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_vert); cudaFree(d_idx); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_blo); cudaFree(d_bhi); cudaFree(d_wide);
        cudaFree(d_bounds); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_flags); cudaFree(d_scan); cudaFree(d_vals);
        cudaFree(d_vals2); cudaFree(d_children); cudaFree(d_tasks); cudaFree(d_tasks2); cudaFree(d_parent); cudaFree(d_arrive);
        cudaFree(d_clusters); cudaFree(d_clusters2); cudaFree(d_nearest); cudaFree(d_leaf_order); cudaFree(d_cnt);
        cudaFree(d_tmp); cudaFree(d_tmp2);
    };
#define CKM(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); free_mesh(ctx); return fail(ctx, e_ == cudaErrorMemoryAllocation ? DRTB_ERR_NOMEM : DRTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    cudaStream_t st = ctx->stream;
    const size_t nn = size_t(n), n_int = nn > 1 ? nn - 1 : 1;
    CKM(cudaMalloc((void**)&ctx->d_tri64, nn * kTri64Stride * sizeof(double)));
    CKM(cudaMalloc((void**)&ctx->d_tri32, nn * kTri32Stride * sizeof(float4)));
    CKM(cudaMalloc((void**)&ctx->d_tri_color, nn * sizeof(int32_t)));
    CKM(cudaMalloc((void**)&ctx->d_tri_emis, nn * sizeof(int32_t)));
    CKM(cudaMalloc((void**)&d_vert, size_t(nv) * 3 * sizeof(double)));
    CKM(cudaMalloc((void**)&d_idx, nn * 3 * sizeof(int32_t)));
    CKM(cudaMalloc((void**)&d_lo, nn * sizeof(float4)));      CKM(cudaMalloc((void**)&d_hi, nn * sizeof(float4)));
    CKM(cudaMalloc((void**)&d_blo, 2 * nn * sizeof(float4))); CKM(cudaMalloc((void**)&d_bhi, 2 * nn * sizeof(float4)));
    CKM(cudaMalloc((void**)&d_wide, nn * kNodeStride * sizeof(float4)));      // a wide node has >= 2 children: < n nodes
    CKM(cudaMalloc((void**)&d_bounds, 6 * sizeof(uint32_t)));
    CKM(cudaMalloc((void**)&d_keys, nn * sizeof(uint64_t)));  CKM(cudaMalloc((void**)&d_keys2, nn * sizeof(uint64_t)));
    CKM(cudaMalloc((void**)&d_vals, nn * sizeof(uint32_t)));  CKM(cudaMalloc((void**)&d_vals2, nn * sizeof(uint32_t)));
    CKM(cudaMalloc((void**)&d_children, n_int * sizeof(int2)));
    CKM(cudaMalloc((void**)&d_tasks, nn * sizeof(int2)));     CKM(cudaMalloc((void**)&d_tasks2, nn * sizeof(int2)));
    CKM(cudaMalloc((void**)&d_leaf_order, nn * sizeof(int32_t)));
    CKM(cudaMalloc((void**)&d_cnt, sizeof(CollapseCounters)));
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
        int m = int(n), made = 0;
        while (m > 1) {
            const int g = (m + T - 1) / T;
            ploc_nearest_kernel<<<g, 256, 0, st>>>(d_clusters, m, bt, d_nearest);
            ploc_flag_kernel<<<g, T, 0, st>>>(d_nearest, m, d_flags);
            CKM(cudaMemsetAsync(d_flags + m, 0, sizeof(uint64_t), st));
            CKM(cub::DeviceScan::ExclusiveSum(d_tmp2, scan_bytes, d_flags, d_scan, m + 1, st));   // scan[m] = totals
            ploc_merge_kernel<<<g, T, 0, st>>>(d_clusters, d_nearest, d_flags, d_scan, m, int(n), made, bt, d_clusters2);
            CKM(cudaGetLastError());
            uint64_t tot = 0;
            CKM(cudaMemcpyAsync(&tot, d_scan + m, sizeof tot, cudaMemcpyDeviceToHost, st));
            CKM(cudaStreamSynchronize(st));
            const int kept = int(tot & 0xffffffffu), merged = int(tot >> 32);
            if (merged < 1 || kept != m - merged) { cleanup(); free_mesh(ctx); return fail(ctx, DRTB_ERR_CUDA, "PLOC iteration made no progress"); }
            made += merged; m = kept;
            std::swap(d_clusters, d_clusters2);
            ctx->launches += 4;
        }
        root = int(n) + made - 1;                            // the last node created
    }
    // 3. collapse to the 4-wide BVH, one level per launch
    const CollapseCounters init_cnt{1, 0, 0, 0};
    const int2 root_task = make_int2(root, 0);
    CKM(cudaMemcpyAsync(d_cnt, &init_cnt, sizeof init_cnt, cudaMemcpyHostToDevice, st));
    CKM(cudaMemcpyAsync(d_tasks, &root_task, sizeof root_task, cudaMemcpyHostToDevice, st));
    CollapseCounters h_cnt = init_cnt;
    for (int n_tasks = 1; n_tasks > 0;) {
        collapse_kernel<<<(n_tasks + T - 1) / T, T, 0, st>>>(d_tasks, n_tasks, int(n), bt, d_vals2, d_wide, d_leaf_order, d_cnt, d_tasks2);
        CKM(cudaGetLastError());
        CKM(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
        CKM(cudaStreamSynchronize(st));
        n_tasks = h_cnt.next;
        CKM(cudaMemsetAsync(&d_cnt->next, 0, sizeof(int), st));
        std::swap(d_tasks, d_tasks2);
        ctx->launches++;
    }
    if (h_cnt.tris != int(n)) { cleanup(); free_mesh(ctx); return fail(ctx, DRTB_ERR_CUDA, "BVH collapse lost triangles"); }
    CKM(cudaMalloc((void**)&ctx->d_nodes, size_t(h_cnt.nodes) * kNodeStride * sizeof(float4)));
    CKM(cudaMemcpyAsync(ctx->d_nodes, d_wide, size_t(h_cnt.nodes) * kNodeStride * sizeof(float4), cudaMemcpyDeviceToDevice, st));
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
    CK(ctx, cudaEventRecord(ctx->ev0, st));
    rc = launch_render(ctx, &oo, seed_img ? ctx->d_seed : nullptr, ctx->d_img, ctx->d_grad, ctx->d_stats, gi, st);
    if (rc != DRTB_OK) return rc;
    CK(ctx, cudaEventRecord(ctx->ev1, st));
    if (want_img && npx3) CK(ctx, cudaMemcpyAsync(img, ctx->d_img, sizeof(double) * npx3, cudaMemcpyDeviceToHost, st));
    if (want_grad && P3) CK(ctx, cudaMemcpyAsync(grad, ctx->d_grad, sizeof(double) * P3, cudaMemcpyDeviceToHost, st));
    if (grad_img && npx3) CK(ctx, cudaMemcpyAsync(grad_img, ctx->d_gimg, sizeof(double) * npx3, cudaMemcpyDeviceToHost, st));
    if (stats) CK(ctx, cudaMemcpyAsync(stats, ctx->d_stats, sizeof(drtb_stats), cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0.f;
        CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        stats->kernel_ms = ms;
        stats->paths = uint64_t(rows) * W * o->spp;
        stats->retraced_paths = 0;
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

int drtb_ipc_alloc(drtb_ctx* ctx, size_t bytes, void** d_ptr, void* handle)
{
    if (!ctx) return DRTB_ERR_INVALID;
    if (!d_ptr || !handle || bytes == 0) return fail(ctx, DRTB_ERR_INVALID, "drtb_ipc_alloc: NULL argument or zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == DRTB_IPC_HANDLE_BYTES, "IPC handle size");
    CK(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    CK(ctx, cudaMalloc(&p, bytes));
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
