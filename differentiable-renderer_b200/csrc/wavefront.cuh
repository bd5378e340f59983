// wavefront.cuh — the mesh-scene integrator: a wavefront over SoA ray / hit /
// state buffers in HBM (config 4 of BASELINE.json).
//
// Why a second integrator.  For the analytic scenes the megakernel
// (render_kernel) is the right shape: a segment is ~600 register-resident
// instructions and nothing reaches HBM.  With a BVH the cost of a segment is
// dominated by a traversal whose length varies 4x between the lanes of a warp
// (ncu: 10.7 of 32 lanes active in node steps, 4.0 in leaf steps even with
// warp-synchronous stepping), and inside a megakernel a lane that is done has
// to wait for the slowest ray of its warp.  Here the traversal is its own
// persistent kernel that REFILLS finished lanes from the ray queue, so the warp
// stays full, and it carries no path state, so more warps fit on an SM.
//
// One batch of paths (<= kBatchPaths camera samples, whole pixels) goes through
//   wf_generate                      camera ray + first segment set-up
//   repeat depth = 0 .. max_depth:
//       wf_traverse                  closest triangle, persistent + dynamic fetch
//       wf_shade                     vertex record, BRDF sample, next segment set-up
//   wf_adjoint                       radiance recurrence + adjoint, pixel and gradient sums
// and every path advances exactly one segment per iteration, so the depth is
// the host's loop counter.  Path slot p <-> (pixel, sample) is static: no
// compaction, the image is summed per pixel in sample order (bit-reproducible).
//
// Buffers (SoA, one element per path slot, every access coalesced):
//   ray_a  = (o.x, o.y, o.z, d.x)   ray_b = (d.y, d.z, t_hit, -)    as R4 (float4 / double4)
//   hit    = scene index of the closest primitive so far, -1 = none
//   state  = alive | lit | n (vertices) | next stream slot
//   rec_w[v][p], rec_prim[v][p]      the vertex record the adjoint sweeps read
// The functions of the reference each stage replaces are the ones named in
// path.cuh; this file only re-schedules them.
#pragma once
#include "path.cuh"

namespace drtb {

constexpr int kBatchPaths = 1 << 22;          // camera samples per wavefront batch
#ifndef DRTB_FETCH_BELOW
#define DRTB_FETCH_BELOW 26
#endif
constexpr int kFetchBelow = DRTB_FETCH_BELOW; // traversal: refill the warp when fewer lanes than this hold a ray

template <typename R> struct alignas(4 * sizeof(R)) R4 { R x, y, z, w; };

constexpr uint32_t kStAlive = 1u << 31, kStLit = 1u << 30;      // state = alive | lit | n << 16 | slot

template <typename R>
struct WfBuffers {
    R4<R>*    ray_a;
    R4<R>*    ray_b;
    int32_t*  hit;
    uint32_t* state;
    R*        rec_w;                          // [max_depth][batch]
    int32_t*  rec_prim;                       // [max_depth][batch]
    int32_t*  alive_count;                    // [max_depth + 2]: paths alive entering depth d
    uint32_t* fetch;                          // traversal queue cursor, one per depth
};

struct WfArgs {
    int32_t  spp, min_bounces, max_depth;
    uint32_t flags;
    double   absorb;
    uint64_t key0;
    int32_t  shard_index, shard_count, band_rows;
    long long first_path;                     // of this batch, in the shard's compact (pixel, sample) order
    int32_t  n_paths;                         // in this batch
    int32_t  batch;                           // slot stride of the record arrays
    int32_t  depth;
    double   seed_scale;
    const double* params;
    const double* seed_img;
    double*  img;
    double*  grad_partial;
    double*  grad_atomic;
    drtb_stats* stats;
    double*  gimg;                            // gradient image of parameter gimg_param (or null)
    int32_t  gimg_param;                      // -1 = none
    int32_t  specular;                        // some analytic primitive has a SpecularBxDF
    MeshView mesh;
};

template <typename R>
__device__ __forceinline__ void wf_pixel_of(const DevScene<R>& sc, const WfArgs& a, long long gp, long long& pix, int& i,
                                            int& x, int& y)
{
    pix = gp / a.spp;
    i = int(gp - pix * a.spp);
    const int r = int(pix / sc.width);
    x = int(pix - (long long)r * sc.width);
    y = a.shard_count > 1 ? ((r / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + r % a.band_rows : r;
}

template <typename R>
__device__ __forceinline__ uint64_t wf_base(const DevScene<R>& sc, const WfArgs& a, int x, int y, int i)
{
    return (a.key0 + ((uint64_t)y * sc.width + x) * (uint64_t)a.spp + (uint64_t)i) * kKeyMul;
}

// Start of a segment at `depth` (Pathtracer::trace, pathtracer.hpp:121-136 up to the
// raycast): Russian roulette, the record-capacity cut, then the analytic primitives.
// Returns false when the path ends here.
template <typename R>
__device__ __forceinline__ bool wf_begin_segment(const DevScene<R>& sc, const WfArgs& a, const WfBuffers<R>& b, int p,
                                                 int depth, uint64_t base, uint32_t& slot, int n, V3<R> o, V3<R> d,
                                                 uint32_t& truncated)
{
    if (depth >= a.min_bounces) {
        if (a.absorb >= 1.0) return false;               // every draw is < 1
        const double u = Real<double>::uniform(stream_draw_base(base, slot++));
        if (u < a.absorb) return false;
    }
    if (n >= a.max_depth) { ++truncated; return false; }
    R t = Real<R>::inf();
    int k = -1;
    if (sc.n_prims > 0) { k = closest_hit(sc, o, d, t); if (k < 0) t = Real<R>::inf(); }
    b.ray_a[p] = {o.x, o.y, o.z, d.x};
    b.ray_b[p] = {d.y, d.z, t, R(0)};
    b.hit[p] = k;
    return true;
}

// ---- camera rays -----------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256)
wf_generate(const __grid_constant__ DevScene<R> sc, const __grid_constant__ WfArgs a, const WfBuffers<R> b)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t truncated = 0;
    bool alive = false;
    if (p < a.n_paths) {
        long long pix; int i, x, y;
        wf_pixel_of(sc, a, a.first_path + p, pix, i, x, y);
        const uint64_t base = wf_base(sc, a, x, y, i);
        const V3<R> o = {sc.eye[0], sc.eye[1], sc.eye[2]};
        const V3<R> d = camera_ray(sc, x, y, base);
        uint32_t slot = 2u;
        alive = wf_begin_segment(sc, a, b, p, 0, base, slot, 0, o, d, truncated);
        b.state[p] = (alive ? kStAlive : 0u) | slot;
    }
    const unsigned m = __ballot_sync(0xffffffffu, alive);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(b.alive_count, __popc(m));
    if (truncated && a.stats) atomicAdd((unsigned long long*)&a.stats->truncated_paths, 1ull);
}

// ---- closest triangle: persistent warps, finished lanes refilled from the queue ---
template <typename R> __device__ __forceinline__ void wf_reload_ray(const WfBuffers<R>& b, int p, V3<R>& o, V3<R>& d)
{
    const R4<R> ra = b.ray_a[p], rb = b.ray_b[p];
    o = {ra.x, ra.y, ra.z}; d = {ra.w, rb.x, rb.y};
}

#ifndef DRTB_WF_MIN_BLOCKS
#define DRTB_WF_MIN_BLOCKS 6
#endif

// Per lane: the ray in float (RayF), the closest hit so far, and the traversal state of bvh.cuh -- the node to
// open next (`next`), up to TWO pending triangle groups (tg0 being tested, tg1 waiting; .y = triangles still to cull
// or test), `eg` = triangles of tg0 that survived the float cull and await the exact double test, and a stack of
// child groups.  Each iteration the warp votes and executes ONE kind of step for the lanes that can take it:
//   N  open node `next` (or first pop a child group: groups whose distance bound is beyond the closest hit are
//      dropped unopened): 8 quantised child boxes; the nearest child hit becomes `next`, the other children hit go
//      on the stack as one group, the triangles hit become a pending group.  A lane may open a node while ONE
//      triangle group is still pending (it traverses on with the closest hit it knows, which only costs an
//      occasional node that the pending triangles would have culled): with that slack ~3/4 of the lanes can take
//      an N step at any time, where "no triangle pending" allowed ~1/2 (profiles/r02_wf_traverse_f64_cw3_summary.txt:
//      14 of 32 lanes per instruction)
//   T  one triangle of tg0: float test (float instantiation) or conservative float cull (double)
//   E  one triangle of eg: reload the ray in double, exact Moller-Trumbore                  (double only)
// so that the expensive exact test (80-byte triangle + 64-byte ray + ~50 FP64 operations) also runs on many lanes.
// The cheap kinds go first when they serve enough lanes (a T step is about a quarter of an N step).
// A lane with nothing left and an empty stack has finished its ray and is refilled from the queue.
#ifndef DRTB_PREFETCH
#define DRTB_PREFETCH 0      // measured: prefetching the next node and triangle to L1 costs 3 % (profiles/README.md, round 2)
#endif
#ifndef DRTB_VOTE_T
#define DRTB_VOTE_T 2      // a T step runs when DRTB_VOTE_T * (lanes with a triangle to test) >= lanes that can open a node
#endif
#ifndef DRTB_VOTE_E
#define DRTB_VOTE_E 3      // an E step runs when DRTB_VOTE_E * (lanes with a survivor) >= the lanes of either other kind (1 / 2 / 3: 1 105 / 1 148 / 1 156 Msegments/s)
#endif
// Fatter steps per vote (round 2, gpurun A/B of tools/ab_mesh.sh, 1 M triangles): up to DRTB_TN triangles per T step
// (1 -> 2: +3.1 % double, +7.8 % float; 3: another +0.6 / +0.9 %; 4: -13 % double, registers) and a second node step
// per vote for the lanes that can still take one (+3 % double; float -1.3 % while the kernel spilled, +1.2 % since
// the node step's diet; a third one changes nothing).
#ifndef DRTB_NN
#define DRTB_NN 2          // node steps per N vote, double
#endif
#ifndef DRTB_NN_F32
#define DRTB_NN_F32 2      // the same, float
#endif
#ifndef DRTB_NN_MIN
#define DRTB_NN_MIN 16     // lanes that must be able to take the extra node step (8 / 12 / 16 / 22 measured: flat below 16)
#endif
#ifndef DRTB_TN
#define DRTB_TN 3          // triangles of the pending group per T step
#endif
template <typename R>
__global__ void __launch_bounds__(128, DRTB_WF_MIN_BLOCKS)
wf_traverse(const __grid_constant__ WfArgs a, const WfBuffers<R> b)
{
    if (b.alive_count[a.depth] == 0) return;
    const MeshView& m = a.mesh;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kNone = 0xffffffffu;
    uint32_t n_nodes = 0, n_tests = 0;
#ifdef DRTB_TRAV_DEBUG
    uint32_t dbg_stale = 0, dbg_children = 0, dbg_leaf_hits = 0;
#endif
    // lane state
    int p = -1;                                   // path slot whose ray this lane traverses, -1 = none
    RayF r{};
    R tmin = R(0);
    float tmax = 0.f;
    int best = -1, sp = 0;
    uint32_t next = kNone;
    uint2 tg0 = make_uint2(0u, 0u), tg1 = make_uint2(0u, 0u);
    uint32_t eg = 0u;
    bool overflow = false;
    __shared__ uint2 s_stack[kSmemStack + 1][128];
    SmemStack<128> stack;
    stack.col = &s_stack[0][threadIdx.x];
    bool exhausted = false;                       // warp-uniform: the queue has no more rays
    for (;;) {
        const unsigned busy = __ballot_sync(0xffffffffu, p >= 0);
        if (!exhausted && __popc(busy) < kFetchBelow) {
            // refill every idle lane with the next rays of the queue (consecutive slots: coalesced)
            const unsigned idle = ~busy;
            uint32_t first = 0;
            if (lane == 0) first = atomicAdd(b.fetch + a.depth, (uint32_t)__popc(idle));
            first = __shfl_sync(0xffffffffu, first, 0);
            if (p < 0) {
                const uint32_t q = first + __popc(idle & ((1u << lane) - 1u));
                if (q < (uint32_t)a.n_paths && (b.state[q] & kStAlive)) {
                    p = int(q);
                    const R4<R> ra = b.ray_a[p], rb = b.ray_b[p];
                    r = make_rayf<R>(V3<R>{ra.x, ra.y, ra.z}, V3<R>{ra.w, rb.x, rb.y});
                    tmin = rb.z; tmax = upper_float<R>(tmin);
                    best = -1; sp = 0; overflow = false;
                    next = 0u; tg0 = make_uint2(0u, 0u); tg1 = make_uint2(0u, 0u); eg = 0u;          // the root
                }
            }
            exhausted = first + (uint32_t)__popc(idle) >= (uint32_t)a.n_paths;
            if (__ballot_sync(0xffffffffu, p >= 0) == 0u) { if (exhausted) break; else continue; }
        } else if (busy == 0u) break;
        const bool on = p >= 0;
        const bool can_e = sizeof(R) == 8 && on && eg != 0u;
        const bool can_t = on && tg0.y != 0u;
        const bool can_n = on && tg1.y == 0u && (next != kNone || sp > 0);      // a node to open (or a group to pop) and room for its triangles
        const int nn = __popc(__ballot_sync(0xffffffffu, can_n)), nt = __popc(__ballot_sync(0xffffffffu, can_t)),
                  ne = sizeof(R) == 8 ? __popc(__ballot_sync(0xffffffffu, can_e)) : 0;
        const int kind = (ne > 0 && DRTB_VOTE_E * ne >= nt && DRTB_VOTE_E * ne >= nn) ? 2 : (nt > 0 && DRTB_VOTE_T * nt >= nn) ? 1 : 0;
        if (kind == 0) {
            // up to DRTB_NN node steps per vote: the second one runs for the lanes that can still open a node (those
            // whose first node gave them no second pending triangle group) when at least DRTB_NN_MIN of them can
#pragma unroll 1
            for (int rep = 0; rep < (sizeof(R) == 8 ? DRTB_NN : DRTB_NN_F32); ++rep) {
                const bool cn = rep == 0 ? can_n : (on && tg1.y == 0u && (next != kNone || sp > 0));
                if (rep > 0 && __popc(__ballot_sync(0xffffffffu, cn)) < DRTB_NN_MIN) break;
                if (cn) {
                    if (next == kNone) {                  // nothing pending: pop one child group
                        uint2 g = stack.get(sp - 1);
                        const bool ok = take_from_popped(g, tmax, r.octinv4, next);
                        const bool keep = ok && (g.y & kHitBits) != 0u;       // its other children stay on the stack
                        stack.put(sp - 1, g, keep);
                        sp -= keep ? 0 : 1;
                    }
                    if (next != kNone) {
                        uint2 ng, tg, rest;
                        int m1, m2;
                        ++n_nodes;
                        node8_step(m, r, tmax, next, ng, tg, m1, m2);
#ifdef DRTB_TRAV_DEBUG
                        dbg_stale += (ng.y & kHitBits) == 0u && tg.y == 0u;
                        dbg_children += __popc(ng.y >> 24);
                        dbg_leaf_hits += __popc(tg.y);
#endif
                        const bool free0 = tg0.y == 0u && eg == 0u;           // tg0's base is still needed while eg is pending
                        tg0 = free0 ? tg : tg0;
                        tg1 = free0 ? tg1 : tg;
                        next = kNone;
                        if (ng.y & kHitBits) {
                            next = take_nearest(ng, m1, m2, rest);
                            const bool more = (rest.y & kHitBits) != 0u;    // the other children hit wait on the stack
                            stack.put(sp, rest, more && sp < kBvhStack);
                            overflow |= more && sp >= kBvhStack;
                            sp += (more && sp < kBvhStack) ? 1 : 0;
#if DRTB_PREFETCH
                            // the child is opened a few iterations from now (its siblings' triangles come first): start its
                            // 128-byte line on the way to L1 now -- the top stall of this kernel is the wait for node data
                            asm volatile("prefetch.global.L1 [%0];" :: "l"(m.nodes + (size_t)next * kNodeStride));
#endif
                        }
#if DRTB_PREFETCH
                        if (tg.y) asm volatile("prefetch.global.L1 [%0];" :: "l"(m.tri32 + (size_t)(tg.x + __ffs(tg.y) - 1) * kTri32Stride));
#endif
                    }
                }
            }
        } else if (kind == 1) {
            if (can_t) {
                const int bit = __ffs(tg0.y) - 1;
                tg0.y &= tg0.y - 1u;
                ++n_tests;
                // up to DRTB_TN triangles of the group in one step: their loads are in flight together and the group
                // needs that many fewer (vote + step) iterations.  A lane with fewer left repeats its first one,
                // which changes nothing (the cull is a pure function, the exact test is idempotent).
                int bits[DRTB_TN];
                TriF T[DRTB_TN];
                bits[0] = bit;
#pragma unroll
                for (int j = 1; j < DRTB_TN; ++j) {
                    const bool more = tg0.y != 0u;
                    bits[j] = more ? __ffs(tg0.y) - 1 : bit;
                    tg0.y &= tg0.y - 1u;                  // 0 & 0xffffffff = 0 when none is left
                    n_tests += more ? 1u : 0u;
                }
#pragma unroll
                for (int j = 0; j < DRTB_TN; ++j) T[j] = load_trif(m, int(tg0.x) + bits[j]);
#pragma unroll
                for (int j = 0; j < DRTB_TN; ++j) {
                    if constexpr (sizeof(R) == 8) {
                        eg |= tri_cull_f(T[j], r, tmax) ? 0u : (1u << bits[j]);
                    } else {
                        const TriData<R> D = {{T[j].v0x, T[j].v0y, T[j].v0z}, {T[j].e1x, T[j].e1y, T[j].e1z}, {T[j].e2x, T[j].e2y, T[j].e2z}};
                        tri_test_exact<R>(D, T[j].id, V3<R>{r.ox, r.oy, r.oz}, V3<R>{r.dx, r.dy, r.dz}, tmin, best);
                    }
                }
                if constexpr (sizeof(R) == 4) tmax = upper_float<R>(tmin);
            }
        } else if (can_e) {
            const int bit = __ffs(eg) - 1;
            eg &= eg - 1u;
            const int tri = __float_as_int(__ldg(&m.tri32[(size_t)(int(tg0.x) + bit) * kTri32Stride + 2].w));   // original index
            V3<R> o, d;
            wf_reload_ray(b, p, o, d);                // o, d stay out of the registers between the rare exact tests
            tri_test_exact<R>(load_tri<R>(m, tri), tri, o, d, tmin, best);
            tmax = upper_float<R>(tmin);
        }
        if (on && tg0.y == 0u && eg == 0u) {
            // tg0 is done: the waiting group moves up; with nothing left anywhere the ray is finished
            tg0 = tg1; tg1.y = 0u;
            if (tg0.y == 0u && next == kNone && sp == 0) {
                if (overflow) {
                    V3<R> o, d;
                    wf_reload_ray(b, p, o, d);
                    brute_closest<R>(m, o, d, tmin, best, n_tests);
                }
                if (best >= 0) { b.ray_b[p].z = tmin; b.hit[p] = m.n_prims + best; }
                p = -1;
            }
        }
    }
    if (a.stats) {
        unsigned long long nn = n_nodes, nt = n_tests;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { nn += __shfl_xor_sync(0xffffffffu, nn, o); nt += __shfl_xor_sync(0xffffffffu, nt, o); }
        if (lane == 0) {
            atomicAdd((unsigned long long*)&a.stats->bvh_nodes, nn);
            atomicAdd((unsigned long long*)&a.stats->tri_tests, nt);
        }
#ifdef DRTB_TRAV_DEBUG
        atomicAdd((unsigned long long*)&a.stats->truncated_paths, (unsigned long long)dbg_stale);
        atomicAdd((unsigned long long*)&a.stats->retraced_paths, (unsigned long long)dbg_children);
        atomicAdd((unsigned long long*)&a.stats->paths, (unsigned long long)dbg_leaf_hits);
#endif
    }
}

// test aid (DRTB_FLAG_NO_BVH): every triangle against every ray
template <typename R>
__global__ void __launch_bounds__(128)
wf_traverse_brute(const __grid_constant__ WfArgs a, const WfBuffers<R> b)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_paths || !(b.state[p] & kStAlive)) return;
    V3<R> o, d;
    wf_reload_ray(b, p, o, d);
    R tmin = b.ray_b[p].z;
    int best = -1;
    uint32_t n_tests = 0;
    brute_closest<R>(a.mesh, o, d, tmin, best, n_tests);
    if (best >= 0) { b.ray_b[p].z = tmin; b.hit[p] = a.mesh.n_prims + best; }
    if (a.stats) atomicAdd((unsigned long long*)&a.stats->tri_tests, (unsigned long long)n_tests);
}

// Out of line: the lobe code (three pow calls) must not raise the register count of the
// all-diffuse shade stage, which runs at 4 blocks of 256 threads per SM.
template <typename R>
__device__ __noinline__ void wf_specular_sample(const V3<R>* in /* n, tg, bt, d */, R expo, R u_theta, R sp, R cp, R* out /* dout.xyz, w */)
{
    R w;
    const V3<R> dout = specular_sample(in[0], in[1], in[2], in[3], expo, u_theta, sp, cp, w);
    out[0] = dout.x; out[1] = dout.y; out[2] = dout.z; out[3] = w;
}

// ---- the vertex: record, sample, next segment (Pathtracer::scatter, pathtracer.hpp:91-115)
template <typename R>
__global__ void __launch_bounds__(256, 4)
wf_shade(const __grid_constant__ DevScene<R> sc, const __grid_constant__ WfArgs a, const WfBuffers<R> b)
{
    if (b.alive_count[a.depth] == 0) return;
    __shared__ BlockScene<R> bs;
    load_block_scene(bs, sc, a.params);
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t truncated = 0;
    bool alive = false, was_alive = false;
    if (p < a.n_paths) {
        uint32_t st = b.state[p];
        was_alive = (st & kStAlive) != 0;
        if (was_alive) {
            Materials<R, true> mat;
            mat.bs = &bs; mat.mesh = a.mesh; mat.params = a.params;
            int n = int((st >> 16) & 0xffu);
            uint32_t slot = st & 0xffffu;
            bool lit = (st & kStLit) != 0;
            const int k = b.hit[p];
            if (k >= 0) {                                         // else: miss, pathtracer.hpp:134-135
                const R4<R> ra = b.ray_a[p], rb = b.ray_b[p];
                V3<R> o = {ra.x, ra.y, ra.z}, d = {ra.w, rb.x, rb.y};
                const R t = rb.z;
                const V3<R> pt = {o.x + t * d.x, o.y + t * d.y, o.z + t * d.z};
                const int2 ec = mat.em_col(k);
                const int em = ec.x, col = ec.y;
                lit |= em >= 0;
                b.rec_prim[(size_t)n * a.batch + p] = k;
                if (col < 0) {                                    // null BxDF, :25-26, 38-39
                    b.rec_w[(size_t)n * a.batch + p] = R(0);
                    ++n;
                } else {
                    V3<R> nrm, tg, bt;
                    if (k >= a.mesh.n_prims) {                    // unit geometric normal, drtb.h
                        const TriData<R> T = load_tri<R>(a.mesh, k - a.mesh.n_prims);
                        nrm = normalize(cross(T.e1, T.e2));
                        unit_frame(nrm, tg, bt);
                    } else {
                        analytic_frame(bs, k, pt, nrm, tg, bt);
                    }
                    long long pix; int i, x, y;
                    wf_pixel_of(sc, a, a.first_path + p, pix, i, x, y);
                    const uint64_t base = wf_base(sc, a, x, y, i);
                    const R u_theta = Real<R>::uniform_fast(stream_draw_base(base, slot));
                    R sp, cp;                                     // phi = 2 * pi * uniform(), bxdf.hpp:74
                    Real<R>::sincos_tab(bs.tab, stream_draw_base(base, slot + 1), &sp, &cp);
                    slot += 2;
                    R w;
                    V3<R> dout;
                    // SpecularBxDF (bxdf.hpp:85-124) on analytic primitives; triangles are diffuse (drtb.h)
                    if (a.specular && k < a.mesh.n_prims && bs.mtype[k] == DRTB_SPECULAR) {
                        const V3<R> in[4] = {nrm, tg, bt, d};
                        R out[4];
                        wf_specular_sample<R>(in, bs.expo[k], u_theta, sp, cp, out);
                        dout = {out[0], out[1], out[2]}; w = out[3];
                        lit |= !(Real<R>::abs(w) < Real<R>::inf());   // NaN * 0 = NaN upstream: see trace_path
                    } else
                        dout = diffuse_sample(nrm, tg, bt, u_theta, sp, cp, w);
                    b.rec_w[(size_t)n * a.batch + p] = w;
                    ++n;
                    const R eps = Real<R>::origin_eps();          // 1e-3, pathtracer.hpp:99
                    o = {Real<R>::fma(eps, dout.x, pt.x), Real<R>::fma(eps, dout.y, pt.y), Real<R>::fma(eps, dout.z, pt.z)};
                    alive = wf_begin_segment(sc, a, b, p, a.depth + 1, base, slot, n, o, dout, truncated);
                }
            }
            b.state[p] = (alive ? kStAlive : 0u) | (lit ? kStLit : 0u) | (uint32_t(n) << 16) | slot;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, alive), s = __ballot_sync(0xffffffffu, was_alive);
    const unsigned tr = __ballot_sync(0xffffffffu, truncated != 0);
    if ((threadIdx.x & 31) == 0) {
        if (m) atomicAdd(b.alive_count + a.depth + 1, __popc(m));
        if (a.stats && s) atomicAdd((unsigned long long*)&a.stats->segments, (unsigned long long)__popc(s));
        if (a.stats && tr) atomicAdd((unsigned long long*)&a.stats->truncated_paths, (unsigned long long)__popc(tr));
    }
}

// A path's record in the wavefront buffers, as radiance_and_adjoint reads it.
template <typename R, int CAP>
struct WfRecordView {
    static constexpr int kCap = CAP;
    static constexpr bool kHasW = true;   // the wavefront keeps the weights in HBM beside the primitives
    const R* w_;
    const int32_t* prim_;
    int stride;
    __device__ __forceinline__ R w(int v) const { return w_[(size_t)v * stride]; }
    __device__ __forceinline__ int prim(int v) const { return prim_[(size_t)v * stride]; }
};

} // namespace drtb
