// render_kernels.cuh — the analytic-scene render kernels (the pixel loop of src/render.cpp:72-86) and their
// launchers, as templates on the arithmetic type R.  Included by render_f64.cu and render_f32.cu, each of
// which instantiates launch_analytic<R> for its precision.
//
//   render_kernel<R, SMALLP, QUEUE, MESH = false, GEN>   persistent megakernel: one lane = one path
//                              (camera sample -> trace -> radiance -> adjoint), lanes of a warp =
//                              consecutive samples of one pixel; warps claim chunks of pixels from a
//                              global counter
//   render_regen_kernel<R, SMALLP>  the same pixel loop for Russian-roulette renders, with path
//                              regeneration over a chunk of pixels
#pragma once
#include "host.hpp"
#include "sinks.cuh"

namespace drtb {
namespace {

using drtbh::fail;
using drtbh::ensure;
using drtbh::first_use;
using drtbh::ChunkPlan;
using drtbh::plan_chunks;
using drtbh::reduce_scratch_rows;
constexpr int kShortDepth = 8;     // QUEUE == 3: record capacity of the short-record variant

// The pixel loop of src/render.cpp:72-86.
//   SMALLP: <= kSmallP parameters, gradients in per-thread shared columns
//   QUEUE : spp >= 32 and max_depth <= kQueueDepth: lit paths are compacted through a per-warp
//           ring before the sweeps.  0: no ring; 1: the ring lives in shared memory; 2: in a global scratch
//           buffer (L1/L2 resident), chosen when the shared ring of a deep record (max_depth > 8
//           in double) would cost resident blocks -- the ring carries only the ~16-21 % of the
//           paths that are lit, so its latency does not matter, the occupancy does; 3: like 2 for records of at
//           most kShortDepth vertices (the headline's 8 bounces): the per-thread record and the sweeps' L_{v+1}
//           array are half as deep, 288 instead of 552 bytes of local memory per resident thread, which is what
//           the L2 has to hold beside the rings; 4: like 1 (shared ring) with the short records
#ifndef DRTB_MIN_BLOCKS
#define DRTB_MIN_BLOCKS 1
#endif
#ifndef DRTB_MESH_MIN_BLOCKS
#define DRTB_MESH_MIN_BLOCKS DRTB_MIN_BLOCKS
#endif
#ifndef DRTB_MIN_BLOCKS_GEN
#define DRTB_MIN_BLOCKS_GEN (DRTB_MIN_BLOCKS < 5 ? DRTB_MIN_BLOCKS : 5)   // double GEN kernels (lobe code, gradient image): 96 registers
#endif
#ifndef DRTB_MIN_BLOCKS_F32
#define DRTB_MIN_BLOCKS_F32 DRTB_MIN_BLOCKS
#endif
//   MESH  : a triangle mesh + BVH is attached (ids are 32-bit, parameters in global memory)
//   GEN   : the general variant -- SpecularBxDF materials (bxdf.hpp:85-124) and the
//           per-pixel gradient image; the all-diffuse kernels do not carry that code
//   MIXED : DRTB_MIXED's fast pass (float): a path whose trace met a close call (path.cuh, ClosestMargin) is not
//           swept; its (pixel, sample) goes on a list that retrace_kernel re-traces in double
template <typename R, bool SMALLP, int QUEUE, bool MESH, bool GEN, bool MIXED = false>
__global__ void __launch_bounds__(kBlock, MESH ? DRTB_MESH_MIN_BLOCKS : sizeof(R) == 4 ? DRTB_MIN_BLOCKS_F32 : GEN ? DRTB_MIN_BLOCKS_GEN : DRTB_MIN_BLOCKS)
render_kernel(const __grid_constant__ DevScene<R> sc, const __grid_constant__ RenderArgs a)
{
    using Id = typename PrimId<MESH>::type;
    extern __shared__ double s_dyn[];              // [acc: n_params*3*kBlock doubles][rings]
    __shared__ BlockScene<R> bs;
    constexpr bool kPixSmem = sizeof(R) == 8;
    __shared__ double s_pix[kPixSmem ? 3 : 1][kPixSmem ? kBlock : 1];

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

    constexpr int kQD = (QUEUE == 3 || QUEUE == 4) ? kShortDepth : kQueueDepth;   // record capacity of the compacting variants
    constexpr bool kRingGlobal = QUEUE == 2 || QUEUE == 3;
    // records with weights: the general and the mesh kernels; the all-diffuse analytic kernels take the weights
    // from the per-primitive table (PathRecord, path.cuh)
    constexpr bool kHasW = GEN || MESH;
    constexpr size_t kRingReal = kHasW ? sizeof(R) : 0;
    // this warp's ring (QUEUE only)
    const int qdepth = a.max_depth;
    unsigned char* ring = reinterpret_cast<unsigned char*>(s_dyn + acc_doubles) +
                          (QUEUE ? size_t(warp) * queue_bytes_per_warp(qdepth, kRingReal, sizeof(Id)) : 0);
    if constexpr (kRingGlobal)
        ring = a.ring_scratch + (size_t(blockIdx.x) * kWarpsPerBlock + warp) * queue_bytes_per_warp(qdepth, kRingReal, sizeof(Id));
    R* ring_w = reinterpret_cast<R*>(ring);
    Id* ring_prim = reinterpret_cast<Id*>(ring + size_t(qdepth) * kQueueSlots * kRingReal);
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
        if (chunk < 0 || chunk >= a.n_chunks) break;
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
            if (lane_ok) {
                const int r = int(pix / W);
                x = int(pix - (long long)r * W);
                y = a.shard_count > 1 ? ((r / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + r % a.band_rows
                                      : r;
            }
            // This lane's share of the pixel.  Double: a shared-memory column, not six registers held through the
            // trace -- the kernel is latency bound, 7 resident blocks of 72 registers beat 5 of 96 by 10 %, and this
            // (with the adjoint seed fetched in the sweep) is what makes 72 registers free of spills.  The float
            // instantiation has the registers (measured: 2.7 % slower with the shared column).
            double acc_r[3] = {0.0, 0.0, 0.0};
            if constexpr (kPixSmem) { s_pix[0][threadIdx.x] = 0.0; s_pix[1][threadIdx.x] = 0.0; s_pix[2][threadIdx.x] = 0.0; }
            double gacc[3] = {0.0, 0.0, 0.0};          // GEN: this lane's share of the pixel's gradient-image value

            // sweeps over one record; accumulates this lane's share of the pixel and the gradients
            auto sweep = [&](const auto& rec, int n) {
                R L0[3];
                // the adjoint seed of this pixel (src/render.cpp:79-80), fetched here rather than held in six registers
                // through the trace: only the lit paths need it
                R g0[3] = {R(0), R(0), R(0)};
                if (want_grad) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        g0[c] = R(a.seed_scale * (a.seed_img ? a.seed_img[pix * 3 + c] : 1.0));
                }
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
                if constexpr (kPixSmem) {                                                          // render.cpp:78
                    s_pix[0][threadIdx.x] += double(L0[0]); s_pix[1][threadIdx.x] += double(L0[1]);
                    s_pix[2][threadIdx.x] += double(L0[2]);
                } else {
                    acc_r[0] += double(L0[0]); acc_r[1] += double(L0[1]); acc_r[2] += double(L0[2]);
                }
                n_lit += (L0[0] != R(0)) | (L0[1] != R(0)) | (L0[2] != R(0));
            };
            // run the sweeps on the first m queued records, one per lane
            auto drain = [&](int m) {
                __syncwarp();
                if (lane < m) {
                    const int slot = (q_head + lane) & (kQueueSlots - 1);
                    QueueView<R, MESH, kQD, kHasW> qv{ring_w + slot, ring_prim + slot};
                    sweep(qv, ring_n[slot]);
                }
                __syncwarp();
                q_head = (q_head + m) & (kQueueSlots - 1);
                q_count -= m;
            };

            for (int pass = 0; pass < passes; ++pass) {
                const int i = i0 + pass * 32;
                bool lit = false, close_call = false;
                int n = 0;
                PathRecord<R, MESH, QUEUE ? kQD : kMaxDepth, kHasW> rec;
                if (lane_ok && i < spp) {
                    const uint64_t key = a.key0 + ((uint64_t)y * W + x) * (uint64_t)spp + (uint64_t)i;
                    const uint64_t base = key * kKeyMul;
                    V3<R> o = {sc.eye[0], sc.eye[1], sc.eye[2]};
                    V3<R> d = camera_ray(sc, x, y, base);
                    const uint32_t seg0 = cnt.segments;
                    n = trace_path<R, MESH, QUEUE ? kQD : kMaxDepth, GEN, MIXED, kHasW>(sc, bs, mat, no_bvh, base, 2u, o, d, a.min_bounces,
                                                                                         a.absorb, a.max_depth, rec, lit, cnt);
                    if constexpr (MIXED) {
                        if (cnt.close_call) {               // this path belongs to the double re-trace, segments and all
                            close_call = true; lit = false;
                            cnt.close_call = 0; cnt.segments = seg0;
                        }
                    }
                    if (!QUEUE && lit) sweep(rec, n);
                }
                if constexpr (MIXED) {
                    const unsigned cm = __ballot_sync(0xffffffffu, close_call);
                    if (cm) {                               // one counter update per warp, entries in lane order
                        unsigned int first = 0;
                        if (lane == 0) first = atomicAdd(a.retrace_count, (unsigned int)__popc(cm));
                        first = __shfl_sync(0xffffffffu, first, 0);
                        const unsigned int at = first + __popc(cm & ((1u << lane) - 1u));
                        if (close_call && at < a.retrace_cap) a.retrace_list[at] = (unsigned long long)pix * (unsigned long long)spp + (unsigned long long)i;
                    }
                }
                if (QUEUE) {
                    const unsigned m = __ballot_sync(0xffffffffu, lit);
                    if (lit) {
                        const int slot = (q_head + q_count + __popc(m & ((1u << lane) - 1u))) & (kQueueSlots - 1);
                        for (int v = 0; v < n; ++v) {
                            if constexpr (kHasW) ring_w[v * kQueueSlots + slot] = rec.w_[v];
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
            double acc[3] = {acc_r[0], acc_r[1], acc_r[2]};
            if constexpr (kPixSmem) { acc[0] = s_pix[0][threadIdx.x]; acc[1] = s_pix[1][threadIdx.x]; acc[2] = s_pix[2][threadIdx.x]; }
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
using drtbh::kRegenPixels;
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
    unsigned char* ring = a.ring_scratch + (size_t(blockIdx.x) * kWarpsPerBlock + warp) * regen_ring_per_warp(0);
    uint8_t* ring_prim = ring;                     // primitive lists only: the weights are table entries (PathRecord)
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
        if (chunk < 0 || chunk >= a.n_chunks) break;
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
                QueueView<R, false, kQueueDepth, false> qv{nullptr, ring_prim + slot};
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
        PathRecord<R, false, kMaxDepth, false> rec;
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
                for (int v = 0; v < n; ++v) ring_prim[v * kQueueSlots + slot] = rec.prim_[v];
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

// DRTB_MIXED's second pass: the paths the float pass set aside, re-traced in double -- the same camera_ray /
// trace_path / radiance_and_adjoint as the parity kernel, one lane per listed path.  Radiance is ADDED to the
// pixel the float pass wrote (red.global.add.f64; a pixel rarely holds two such paths, so the order of these
// additions matters in the last bit at most), gradients go through the per-thread shared columns into one partial
// row per block behind the float pass's rows.  More close calls than the list holds (retrace_cap = 1/8 of the
// paths; ~0.3 % are expected) cannot be served: the image and the gradients are poisoned with NaN rather than
// returned incomplete (drtb_render re-renders in double when it sees that).
template <typename R>
__global__ void __launch_bounds__(kBlock)
retrace_kernel(const __grid_constant__ DevScene<R> sc, const __grid_constant__ RenderArgs a, int partial_row0)
{
    static_assert(sizeof(R) == 8, "the re-trace is the double instantiation");
    extern __shared__ double s_dyn[];
    __shared__ BlockScene<R> bs;
    __shared__ double s_red[kSmallP * 3][kWarpsPerBlock];
    const bool want_grad = (a.flags & DRTB_FLAG_GRAD) != 0;
    const int P3 = sc.n_params * 3;
    double* s_acc = s_dyn;
    load_block_scene(bs, sc, a.params);
    for (int i = threadIdx.x; i < (want_grad ? P3 * kBlock : 0); i += kBlock) s_acc[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int listed = *a.retrace_count, n = min(listed, a.retrace_cap);
    const int W = sc.width, spp = a.spp;
    const R inv_p = a.absorb < 1.0 ? R(1.0 / (1.0 - a.absorb)) : R(0);
    SmemSink ssink{s_acc + threadIdx.x};
    Materials<R, false> mat;
    mat.bs = &bs;
    TraceCounters cnt;
    uint32_t n_lit = 0;
    for (unsigned int e = blockIdx.x * kBlock + threadIdx.x; e < n; e += gridDim.x * kBlock) {
        const unsigned long long entry = a.retrace_list[e];
        const long long pix = (long long)(entry / (unsigned long long)spp);
        const int i = int(entry - (unsigned long long)pix * (unsigned long long)spp);
        const int r = int(pix / W), x = int(pix - (long long)r * W);
        const int y = a.shard_count > 1 ? ((r / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + r % a.band_rows : r;
        const uint64_t key = a.key0 + ((uint64_t)y * W + x) * (uint64_t)spp + (uint64_t)i;
        const uint64_t base = key * kKeyMul;
        V3<R> o = {sc.eye[0], sc.eye[1], sc.eye[2]};
        V3<R> d = camera_ray(sc, x, y, base);
        PathRecord<R, false, kMaxDepth> rec;
        bool lit;
        const int nv = trace_path<R, false, kMaxDepth, false>(sc, bs, mat, false, base, 2u, o, d, a.min_bounces, a.absorb, a.max_depth, rec, lit, cnt);
        if (!lit) continue;
        R g0[3] = {R(0), R(0), R(0)}, L0[3];
        if (want_grad) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g0[c] = R(a.seed_scale * (a.seed_img ? a.seed_img[pix * 3 + c] : 1.0));
        }
        radiance_and_adjoint(mat, rec, nv, a.min_bounces, inv_p, want_grad, g0, L0, ssink);
        n_lit += (L0[0] != R(0)) | (L0[1] != R(0)) | (L0[2] != R(0));
        if (a.img) {
#pragma unroll
            for (int c = 0; c < 3; ++c) if (L0[c] != R(0)) atomicAdd(a.img + pix * 3 + c, double(L0[c]) / double(spp));
        }
    }
    const bool overflow = listed > a.retrace_cap;
    if (want_grad) {
        __syncthreads();
        for (int j = 0; j < P3; ++j) {
            const double v = warp_sum(s_acc[j * kBlock + threadIdx.x]);
            if (lane == 0) s_red[j][warp] = v;
        }
        __syncthreads();
        if (threadIdx.x < P3) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kWarpsPerBlock; ++w) v += s_red[threadIdx.x][w];
            a.grad_partial[((size_t)partial_row0 + blockIdx.x) * P3 + threadIdx.x] = overflow ? __longlong_as_double(0x7ff8000000000000ll) : v;
        }
    }
    if (overflow && a.img && blockIdx.x == 0 && threadIdx.x < 3) a.img[threadIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
    if (a.stats) {
        unsigned long long seg = cnt.segments, litp = n_lit;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { seg += __shfl_xor_sync(0xffffffffu, seg, o); litp += __shfl_xor_sync(0xffffffffu, litp, o); }
        if (lane == 0) {
            atomicAdd((unsigned long long*)&a.stats->segments, seg);
            atomicAdd((unsigned long long*)&a.stats->lit_paths, litp);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) a.stats->retraced_paths = listed;
    }
}

// Enqueue the re-trace behind the float pass: `row0` partial rows are taken; returns the rows it adds.
template <typename R>
int launch_retrace(drtb_ctx* ctx, const DevScene<R>& sc, RenderArgs& a, int P3, bool want_grad, size_t row0, cudaStream_t stream,
                   size_t& rows_added)
{
    const size_t smem = want_grad ? size_t(P3) * kBlock * sizeof(double) : 0;
    CK(ctx, cudaFuncSetAttribute(retrace_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int grid = ctx->sm_count * 4;
    rows_added = want_grad ? size_t(grid) : 0;
    if (ctx->dry) {
        if (first_use(ctx, (const void*)retrace_kernel<R>)) {
            cudaFuncAttributes fa;
            CK(ctx, cudaFuncGetAttributes(&fa, retrace_kernel<R>));       // loads the kernel (lazy module loading)
        }
        return DRTB_OK;
    }
    retrace_kernel<R><<<grid, kBlock, smem, stream>>>(sc, a, int(row0));
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    return DRTB_OK;
}

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

template <typename R, bool SMALLP, int QUEUE, bool MESH, bool GEN, bool MIXED = false>
int launch_variant(drtb_ctx* ctx, const DevScene<R>& sc, RenderArgs& a, size_t smem, long long n_tasks,
                   int P3, bool want_grad, cudaStream_t stream, size_t& rows_out)
{
    int per_sm = 0;
    int rc = occupancy(ctx, render_kernel<R, SMALLP, QUEUE, MESH, GEN, MIXED>, smem, per_sm);
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
    if (!ctx->d_task_counter) {
        CK(ctx, cudaMalloc(&ctx->d_task_counter, sizeof(unsigned long long)));
        CK(ctx, cudaMemset(ctx->d_task_counter, 0, sizeof(unsigned long long)));     // the first-use launches claim from it too
    }
    a.task_counter = ctx->d_task_counter;
    // gradient partials: one row per chunk (SMALLP, summed in chunk order whoever ran the chunk) or
    // per block (shared atomic columns)
    size_t rows = 0;
    if (want_grad && SMALLP) rows = size_t(n_chunks);
    else if (want_grad && !MESH) rows = size_t(grid);
    if (rows) {
        const size_t total = rows + (MIXED ? ctx->partial_extra_rows : 0);      // the re-trace's rows follow (launch_retrace)
        rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (total + reduce_scratch_rows(total)) * P3);
        if (rc != DRTB_OK) return rc;
        a.grad_partial = ctx->d_partial;
    }
    if (QUEUE == 2 || QUEUE == 3) {
        const size_t per_warp = queue_bytes_per_warp(a.max_depth, (GEN || MESH) ? sizeof(R) : 0, MESH ? sizeof(int32_t) : sizeof(uint8_t));
        rc = ensure(ctx, ctx->d_ring, ctx->ring_cap, size_t(grid) * kWarpsPerBlock * per_warp / sizeof(double));
        if (rc != DRTB_OK) return rc;
        a.ring_scratch = reinterpret_cast<unsigned char*>(ctx->d_ring);
    }
    rows_out = rows;
    if (ctx->dry) {
        // first use of this instantiation: one launch on zero chunks (module load, local-memory reservation)
        if (first_use(ctx, (const void*)render_kernel<R, SMALLP, QUEUE, MESH, GEN, MIXED>)) {
            RenderArgs w = a;
            w.n_chunks = 0; w.stats = nullptr;
            render_kernel<R, SMALLP, QUEUE, MESH, GEN, MIXED><<<1, kBlock, smem, stream>>>(sc, w);
            CK(ctx, cudaGetLastError());
            CK(ctx, cudaStreamSynchronize(stream));
        }
        return DRTB_OK;
    }
    CK(ctx, cudaMemsetAsync(ctx->d_task_counter, 0, sizeof(unsigned long long), stream));
    render_kernel<R, SMALLP, QUEUE, MESH, GEN, MIXED><<<int(grid), kBlock, smem, stream>>>(sc, a);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
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
    if (!ctx->d_task_counter) {
        CK(ctx, cudaMalloc(&ctx->d_task_counter, sizeof(unsigned long long)));
        CK(ctx, cudaMemset(ctx->d_task_counter, 0, sizeof(unsigned long long)));     // the first-use launches claim from it too
    }
    a.task_counter = ctx->d_task_counter;
    int rc = ensure(ctx, ctx->d_ring, ctx->ring_cap, size_t(grid) * kWarpsPerBlock * regen_ring_per_warp(0) / sizeof(double));
    if (rc != DRTB_OK) return rc;
    a.ring_scratch = reinterpret_cast<unsigned char*>(ctx->d_ring);
    const size_t rows = !want_grad ? 0 : SMALLP ? size_t(a.n_chunks) : size_t(grid);
    if (rows) {
        rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (rows + reduce_scratch_rows(rows)) * P3);
        if (rc != DRTB_OK) return rc;
        a.grad_partial = ctx->d_partial;
    }
    rows_out = rows;
    if (ctx->dry) {
        if (first_use(ctx, (const void*)render_regen_kernel<R, SMALLP>)) {
            RenderArgs w = a;
            w.n_chunks = 0; w.stats = nullptr;
            render_regen_kernel<R, SMALLP><<<1, kBlock, smem, stream>>>(sc, w);
            CK(ctx, cudaGetLastError());
            CK(ctx, cudaStreamSynchronize(stream));
        }
        return DRTB_OK;
    }
    CK(ctx, cudaMemsetAsync(ctx->d_task_counter, 0, sizeof(unsigned long long), stream));
    render_regen_kernel<R, SMALLP><<<int(grid), kBlock, smem, stream>>>(sc, a);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    return DRTB_OK;
}

// The one entry point of this translation unit: picks the instantiation.
template <typename R>
int launch_analytic(drtb_ctx* ctx, const DevScene<R>& sc, RenderArgs& a, const drtbh::AnalyticLaunch& l, cudaStream_t stream,
                    size_t& rows)
{
    if (l.regen)
        return l.smallp ? launch_regen<R, true>(ctx, sc, a, l.npix, l.P3, l.want_grad, stream, rows)
                        : launch_regen<R, false>(ctx, sc, a, l.npix, l.P3, l.want_grad, stream, rows);
    if constexpr (sizeof(R) == 4) {
        if (l.mixed)                                  // DRTB_MIXED's fast pass: all-diffuse, <= kSmallP parameters (drtb.cu)
            return l.queue == 2 ? launch_variant<R, true, 2, false, false, true>(ctx, sc, a, l.smem, l.n_tasks, l.P3, l.want_grad, stream, rows)
                 : l.queue == 1 ? launch_variant<R, true, 1, false, false, true>(ctx, sc, a, l.smem, l.n_tasks, l.P3, l.want_grad, stream, rows)
                                : launch_variant<R, true, 0, false, false, true>(ctx, sc, a, l.smem, l.n_tasks, l.P3, l.want_grad, stream, rows);
    }
#define DRTB_LAUNCH(SP, Q, G) launch_variant<R, SP, Q, false, G>(ctx, sc, a, l.smem, l.n_tasks, l.P3, l.want_grad, stream, rows)
#define DRTB_BY_QUEUE(SP, G) (l.queue == 2 ? DRTB_LAUNCH(SP, 2, G) : l.queue == 1 ? DRTB_LAUNCH(SP, 1, G) : DRTB_LAUNCH(SP, 0, G))
    // short records in the global ring: the all-diffuse kernel with per-thread gradient columns only (drtb.cu asks
    // for it only there)
    if (l.queue == 3 && !l.gen && l.smallp) return DRTB_LAUNCH(true, 3, false);
    if (l.queue == 4 && !l.gen && l.smallp) return DRTB_LAUNCH(true, 4, false);
    // GEN: SpecularBxDF materials and/or a gradient image; the all-diffuse kernels do not carry that code
    if (l.gen) return l.smallp ? DRTB_BY_QUEUE(true, true) : DRTB_BY_QUEUE(false, true);
    return l.smallp ? DRTB_BY_QUEUE(true, false) : DRTB_BY_QUEUE(false, false);
#undef DRTB_BY_QUEUE
#undef DRTB_LAUNCH
}

} // namespace
} // namespace drtb
