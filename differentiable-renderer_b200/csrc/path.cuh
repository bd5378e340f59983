// path.cuh — the hot path: camera ray, closest hit, diffuse sampling, vertex
// record, radiance recurrence and the tape-free adjoint.  One device function
// per reference function; see each for the file:line it replaces.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/drtb.h"
#include "real.cuh"
#include "rng.cuh"

namespace drtb {

constexpr int kMaxPrims   = 32;    // analytic scenes ride in the kernel-parameter constant bank
constexpr int kMaxParams  = 64;    // RGB parameters staged in shared memory
constexpr int kMaxDepth   = 64;    // vertex-record capacity per path
constexpr int kSmallP     = 8;     // <= this many parameters: per-thread smem gradient columns
constexpr int kBlock      = 128;
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr int kFast       = 8;     // planes / spheres scanned by straight-line code (compile-time slots)
constexpr int kAxisFast   = 2;     // axis-aligned unit planes per axis scanned by straight-line code (a slab)
constexpr int kSlots      = 2 * kFast + kMaxPrims;
constexpr int kMaxPeers   = 8;     // GPUs of one NVSwitch box whose full images a render can fill
#ifndef DRTB_PACK_HIT
#define DRTB_PACK_HIT 1
#endif
constexpr bool kPackHit   = DRTB_PACK_HIT;
#ifndef DRTB_SKIP_LAST
#define DRTB_SKIP_LAST 1           // 0: sample the BxDF at a vertex whose continuation is certainly absorbed (A/B aid)
#endif   // double analytic kernels: closest hit on integer keys (Closest<double, true>)

// Scene as the kernels see it.  Passed BY VALUE as a __grid_constant__ kernel
// parameter: the closest-hit scan indexes prim[] with a warp-uniform i, so
// every operand comes straight out of the constant bank with no load
// instruction.  (Flattened Scene<T>/Shape<T>/Camera<T>, src/render.cpp:26-65.)
template <typename R>
struct alignas(16) DevScene {
    // Scan slots.  [0, kFast): the first planes of the scene, RIGHT-aligned, and
    // [kFast, 2 kFast): the first spheres, right-aligned -- the closest-hit scan
    // enters straight-line code at slot kFast - n, so every operand of those
    // tests is a compile-time constant-bank address (no LDC, no loop).  From
    // 2 kFast on: the planes, then the spheres, that did not fit (rolled loops).
    // id[] is the position in Scene<T>, which decides exact ties (pathtracer.hpp:80).
    // Planes whose normal is exactly +-e_x, +-e_y or +-e_z (the walls of a box
    // scene; 5 of the Cornell box's 6) are NOT scanned there: for n = s e_a,
    // t = (o.n - off) / (d.(-n)) = (s off - o_a) / d_a with every product exact, so
    // they sit in aa[][] (up to kAxisFast per axis, right-aligned) and cost one
    // subtraction and one multiplication by the per-segment 1 / d_a.  Their n, off
    // stay in prim[] behind the scanned slots for the per-lane lookups.
    R       prim[kSlots][4];             // plane: n.xyz (RAW), offset ; sphere: c.xyz, r
    R       sph[kSlots][4];              // sphere slots as the scan reads them: c.xyz, r * r (host double) -- two 128-bit loads
    int32_t id[kSlots];                  // scan slot -> scene index
    struct alignas(16) AxisSlot { R c; int32_t id; int32_t pad; };   // s * offset (the plane is p_a = c) and the scene index: one load
    AxisSlot aa[3][kAxisFast];
    int32_t n_aa[3];
    int32_t n_fast_planes, n_fast_spheres, n_over_planes, n_over_spheres;
    // Indexed by SCENE index:
    R       frame[kMaxPrims][6];         // planes: make_frame(n) tangent, bitangent (bxdf.hpp:29-41), host double
    int8_t  type[kMaxPrims];             // DRTB_SPHERE | DRTB_PLANE
    R       wtab[kMaxPrims];             // diffuse weight cos / pdf = pi |n|^2 of the primitive (diffuse_sample), host double
    int32_t color[kMaxPrims];            // param index of the albedo, -1 = null BxDF
    int32_t emis[kMaxPrims];             // param index of the emission, -1 = no emitter
    int8_t  slot[kMaxPrims];             // scene index -> scan slot
    int8_t  mtype[kMaxPrims];            // DRTB_DIFFUSE | DRTB_SPECULAR (only read by the SPEC kernels)
    R       expo[kMaxPrims];             // SpecularBxDF::m_exponent (bxdf.hpp:123)
    int32_t n_prims;
    int32_t n_params;
    // Camera (camera.hpp:51-60), constants folded on the host in double with the
    // host libm (the same tan() the reference calls):
    R eye[3], fwd[3], right[3], nup[3];  // nup = -1 * up
    R aspect, tan_half;                  // W/H, tan(vfov/2)
    R inv_w, inv_h;                      // only used by the float instantiation
    int32_t width, height;
    // DRTB_MIXED: 1 if every primitive either passes EXACTLY through the eye (its t is an exact 0 for every camera ray
    // in float as in double, and is rejected by t > 0 in both) or stays clear of it by more than the near-zero margin
    int32_t eye_clear;
};


template <typename R> struct V3 { R x, y, z; };

template <typename R> __device__ __forceinline__ R dot(V3<R> a, V3<R> b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z;        // vector.hpp:573-578
}
template <typename R> __device__ __forceinline__ V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename R> __device__ __forceinline__ V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename R> __device__ __forceinline__ V3<R> operator*(V3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename R> __device__ __forceinline__ V3<R> cross(V3<R> a, V3<R> b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};   // vector.hpp:592-600
}
template <typename R> __device__ __forceinline__ V3<R> normalize(V3<R> a)
{
    return a * Real<R>::rsqrt(dot(a, a));             // vector.hpp:580-590
}

} // namespace drtb
#include "bvh.cuh"
namespace drtb {

struct RenderArgs {
    int32_t  spp, min_bounces, max_depth;
    uint32_t flags;
    double   absorb;
    uint64_t key0;                       // seed * kSeedMul
    int32_t  shard_index, shard_count, band_rows, shard_rows;
    double   seed_scale;
    const double* params;                // n_params x 3 (device)
    const double* seed_img;              // shard_rows x W x 3 or null
    double*  img;                        // shard_rows x W x 3 or null
    double*  grad_partial;               // gridDim.x x (n_params*3)  (small-P path)
    double*  grad_atomic;                // n_params*3, pre-zeroed     (large-P path)
    drtb_stats* stats;                   // or null
    double*  gimg;                       // shard_rows x W x 3 gradient image of parameter gimg_param, or null
    int32_t  gimg_param;                 // -1 = none
    int32_t  sink_cols;                  // shared atomic gradient columns (9 .. 64 parameters), a power of two
    MeshView mesh;                       // n_tris == 0: analytic scene only
    double*  peer_img[kMaxPeers];        // FULL images (H x W x 3, row = image row) on every GPU of the job,
    int32_t  n_peer_img;                 // written pixel by pixel over NVLink (drtb_set_image_peers); 0 = off
    unsigned char* ring_scratch;         // QUEUE == 2: per-warp lit-path rings in global memory
    unsigned long long* task_counter;    // zeroed before the launch: next unclaimed chunk of warp tasks
    int32_t  chunk_tasks;                // consecutive warp tasks per big chunk
    long long n_big_chunks, n_chunks;    // chunks [n_big_chunks, n_chunks) are single tasks (render_kernel)
    int32_t  small_chunk;                // ... or small_chunk pixels each (render_regen_kernel, whose tasks are pixels)
    // DRTB_MIXED: paths whose float trace met a close call, as (pixel-in-shard * spp + sample), for the double re-trace
    unsigned long long* retrace_list;
    unsigned int* retrace_count;         // entries appended so far (may exceed retrace_cap: the excess is dropped AND reported)
    unsigned int retrace_cap;
};

// Per-block shared copy of what is looked up with a PER-LANE index (the prim a
// lane actually hit, the parameters of its material).
template <typename R>
struct BlockScene {
    R       prim[kMaxPrims][4];
    R       frame[kMaxPrims][6];         // plane tangent + bitangent
    R       param[kMaxParams * 3];       // staged only when n_params <= kMaxParams
    int2    em_col[kMaxPrims];           // (emission, albedo) parameter indices: one 64-bit load per vertex
    R       wtab[kMaxPrims];             // diffuse weight of the primitive (records without weights, PathRecord)
    int8_t  type[kMaxPrims];
    int8_t  mtype[kMaxPrims];
    R       expo[kMaxPrims];
    SinCosTab<R> tab;                    // double: (sin, cos) table of Real<double>::sincos_tab
};

template <typename R>
__device__ __forceinline__ void load_block_scene(BlockScene<R>& bs, const DevScene<R>& sc,
                                                 const double* __restrict__ params)
{
    for (int i = threadIdx.x; i < sc.n_prims * 4; i += blockDim.x)      // bs.prim is by SCENE index
        bs.prim[i >> 2][i & 3] = sc.prim[sc.slot[i >> 2]][i & 3];
    for (int i = threadIdx.x; i < sc.n_prims * 6; i += blockDim.x)
        bs.frame[i / 6][i % 6] = sc.frame[i / 6][i % 6];
    for (int i = threadIdx.x; i < sc.n_prims; i += blockDim.x) {
        bs.type[i] = sc.type[i]; bs.em_col[i] = make_int2(sc.emis[i], sc.color[i]); bs.wtab[i] = sc.wtab[i];
        bs.mtype[i] = sc.mtype[i]; bs.expo[i] = sc.expo[i];
    }
    if (sc.n_params <= kMaxParams)
        for (int i = threadIdx.x; i < sc.n_params * 3; i += blockDim.x) bs.param[i] = R(params[i]);
    if constexpr (Real<R>::kTable)
        for (int i = threadIdx.x; i < kSinCosEntries; i += blockDim.x) bs.tab.t[i] = g_sincos_tab[i];
}

// ---------------------------------------------------------------------------
// Camera<T>::sample, camera.hpp:51-60: two draws (slots 0, 1), pdf = 1.
// ---------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ V3<R> camera_ray(const DevScene<R>& sc, int x, int y, uint64_t base)
{
    R u0 = Real<R>::uniform_fast(stream_draw_base(base, 0));
    R u1 = Real<R>::uniform_fast(stream_draw_base(base, 1));
    R s = Real<R>::div(R(x) + u0, R(sc.width));
    R t = Real<R>::div(R(y) + u1, R(sc.height));
    R cx = (R(2) * s - R(1)) * sc.aspect * sc.tan_half;
    R cy = (R(2) * t - R(1)) * sc.tan_half;
    V3<R> d = {sc.fwd[0] + cx * sc.right[0] + cy * sc.nup[0],
               sc.fwd[1] + cx * sc.right[1] + cy * sc.nup[1],
               sc.fwd[2] + cx * sc.right[2] + cy * sc.nup[2]};
    return normalize(d);
}

// The running closest hit of one scan.  offer(t, id) is the reference's
// acceptance (pathtracer.hpp:80, shape.hpp:55, 91-99): t > 0, strictly closer
// than the best so far, the lower scene index winning exact ties (the scan
// order here is not the scene order, so the index takes part in the compare).
// t = +-inf / NaN (ray parallel to a plane) fails every compare, as upstream
// where inf >= tmin skips it.
template <typename R, bool PACK>
struct Closest {
    R bt; int best;
    __device__ __forceinline__ Closest() : bt(Real<R>::inf()), best(-1) {}
    __device__ __forceinline__ void offer(R t, int id)
    {
        const bool closer = (t < bt) | ((t == bt) & (id < best));           // bitwise: no branch
        if (Real<R>::is_pos(t) & closer) { bt = t; best = id; }
    }
    __device__ __forceinline__ void offer_nz(R t, int id) { offer(t, id); }
    __device__ __forceinline__ int finish(R& tmin) const { tmin = bt; return best; }
};
// Double, analytic scenes: the three FP64 compares per primitive (six issue
// slots of the half-rate pipe) become integer ones.  Positive doubles order
// like their bit patterns, negative ones and NaNs compare above +inf as
// unsigned integers, so "0 < t < bt" is one unsigned 64-bit compare plus the
// sign-word test that rejects +0.  The scene index rides in the five low
// mantissa bits (kMaxPrims = 32), which makes the tie rule part of the same
// compare: equal t -> lower index.  Cost: the t that comes back differs from
// the computed one by < 32 ulp (7e-15 relative; the Newton rcp/sqrt feeding it
// are good to 2 ulp), and two primitives whose t agree to 32 ulp are ordered by
// index instead of by t.
template <>
struct Closest<double, true> {
    static constexpr uint32_t keep = ~uint32_t(kMaxPrims - 1);
    unsigned long long key;
    __device__ __forceinline__ Closest() : key(0x7ff0000000000000ull) {}
    __device__ __forceinline__ void offer(double t, int id)
    {
        const int hi = __double2hiint(t);
        const uint32_t lo = (uint32_t(__double2loint(t)) & keep) | uint32_t(id);
        const unsigned long long k = (unsigned long long)(uint32_t)hi << 32 | lo;
        if ((hi > 0) & (k < key)) key = k;
    }
    // The same for a t that is never +0 (the plane tests: Real<double>::mul_nz turns an exact zero into the
    // number -2^-1000): as UNSIGNED integers negative doubles, -0 and NaNs compare above the +inf the key
    // starts from, so the one compare rejects them -- five instructions instead of six.
    __device__ __forceinline__ void offer_nz(double t, int id)
    {
        const uint32_t lo = (uint32_t(__double2loint(t)) & keep) | uint32_t(id);
        const unsigned long long k = (unsigned long long)(uint32_t)__double2hiint(t) << 32 | lo;
        if (k < key) key = k;
    }
    __device__ __forceinline__ int finish(double& tmin) const
    {
        tmin = __longlong_as_double((long long)key);
        return key == 0x7ff0000000000000ull ? -1 : int(uint32_t(key) & ~keep);
    }
};

// DRTB_MIXED: the closest hit in FLOAT that also says whether the decision was a close call.  A float path takes
// the same sequence of primitives as the double path as long as no closest-hit decision along it could go the
// other way; its radiance then differs by the float rounding of the weights only (~1e-6 relative after 8 vertices).
// A decision is a close call if
//   * the runner-up is within kGapAbs + kGapRel t of the winner (hit point within ~1e-3 of an edge or silhouette),
//   * some primitive's t is within kNearZero of 0 (acceptance t > 0; a ray leaving a surface sits at |t| = 1e-3,
//     the origin offset of pathtracer.hpp:99, a decade away),
//   * a sphere's discriminant is within kDiscRel r^2 of 0 (hit / miss of a grazing ray), or
//   * a plane is nearly parallel to the ray (|d.n| <= kParallel: t is huge or infinite either way).
// The float evaluation error of t is ~1e-6 at scene scale 6 and the float path drifts from the double one by
// ~1e-6 per plane bounce, one to two orders below these margins.  Two more things can part the two paths, and
// trace_path watches both:
//   * make_frame (bxdf.hpp:29-41) picks its helper axis by |n.x| < |n.y|: on a sphere whose normal has the two within
//     kFrameGap of each other the float and the double frame may differ by a rotation -- a close call like the others;
//   * DRIFT: the float path's hit points move away from the double path's, and what a margin has to cover is that
//     distance, not just one segment's rounding.  trace_path carries a bound on it: at a hit at distance t under
//     incidence cosine c the position bound becomes e' = (e_pos + t e_dir + rounding) / c (a displaced ray meets the
//     surface up to 1 / c further along it); the direction sampled off a PLANE carries only fresh rounding (its
//     frame is a constant), off a SPHERE of radius r it inherits e' / r (the normal is (hit point - centre) / r).
//     Every margin of the next scan is widened by kDriftK times the bound carried to it (`slack`), so the close-call
//     test scales with how far the two paths may already be apart; a bound beyond kDriftMax gives the path up.
// These are engineering margins, not a proof: what is guaranteed is the re-trace of every path that trips one, and
// what is TESTED is the outcome -- segment and lit-path counts equal to the double reference's and the parity bar at
// BASELINE's sizes (tests/test_gpu_mixed.py: every pixel of full config 2 through its tile sums).
// Such a path is not used: its key goes on a list and a double kernel re-traces it (render_kernels.cuh, retrace_kernel).
constexpr float kGapAbs = 1e-3f, kGapRel = 2e-4f, kNearZero = 1e-4f, kDiscRel = 1e-3f, kParallel = 1e-5f;
constexpr float kFrameGap = 1e-3f, kDriftK = 2.0f, kDriftMax = 2e-2f, kRoundPos = 3e-7f, kRoundDir = 3e-7f;
struct ClosestMargin {
    static constexpr bool kMargin = true;
    float bt, bt2, ta, slack; int best; bool flag;   // ta = the smallest |t| any primitive returned, accepted or not; slack: see DRIFT
    __device__ __forceinline__ ClosestMargin()
        : bt(Real<float>::inf()), bt2(Real<float>::inf()), ta(Real<float>::inf()), slack(0.0f), best(-1), flag(false) {}
    // Eight instructions per primitive (the plain float scan takes six): the two smallest accepted t by a min / max
    // pair, the winner's id, and min |t|.  Exact ties need no index rule here: a tie is a close call by definition.
    __device__ __forceinline__ void offer(float t, int id)
    {
        ta = fminf(ta, fabsf(t));                                // NaN (ray parallel to a plane through its origin) is dropped
        const float tp = t > 0.0f ? t : Real<float>::inf();      // acceptance t > 0, shape.hpp:55, 91-99
        bt2 = fminf(bt2, fmaxf(bt, tp));
        best = tp < bt ? id : best;
        bt = fminf(bt, tp);
    }
    __device__ __forceinline__ void offer_nz(float t, int id) { offer(t, id); }
    __device__ __forceinline__ void rejected(float t) { ta = fminf(ta, fabsf(t)); }   // a sphere's first root when it is <= 0
    // skip_zero: the caller knows that no primitive can be within kNearZero of t = 0 except exactly AT 0 (camera rays
    // of a scene whose eye lies exactly on a plane, as the Cornell box's does: DevScene::eye_clear)
    __device__ __forceinline__ int finish(float& tmin, bool skip_zero)
    {
        flag |= (bt2 - bt) < kGapAbs + slack + kGapRel * bt;  // a miss (bt = inf) gives NaN: no flag
        // (capped below 1e-3: a ray leaving a surface sees that surface at |t| = 1e-3 whatever the drift, pathtracer.hpp:99)
        if (!skip_zero) flag |= ta < fminf(kNearZero + slack, 5e-4f);
        tmin = bt;
        return best;
    }
};
template <typename C> struct HasMargin { static constexpr bool value = false; };
template <> struct HasMargin<ClosestMargin> { static constexpr bool value = true; };

// Axis-aligned unit plane p_a = c: t = (c - o_a) / d_a, inv_a = 1 / d_a once per segment.
template <typename R, typename C>
__device__ __forceinline__ void axis_plane_test(const DevScene<R>& sc, int axis, int slot, R o_a, R inv_a, C& cl)
{
    cl.offer_nz(Real<R>::mul_nz(sc.aa[axis][slot].c - o_a, inv_a), sc.aa[axis][slot].id);
}

template <typename R, typename C>
__device__ __forceinline__ void plane_test(const DevScene<R>& sc, int slot, V3<R> o, V3<R> d, C& cl)
{
    const R a0 = sc.prim[slot][0], a1 = sc.prim[slot][1], a2 = sc.prim[slot][2], a3 = sc.prim[slot][3];
    const R h = Real<R>::fma(o.x, a0, Real<R>::fma(o.y, a1, Real<R>::fma(o.z, a2, -a3)));   // o.n - offset
    const R g = Real<R>::fma(d.x, a0, Real<R>::fma(d.y, a1, d.z * a2));                     // d.n ; t = h / -g
    if constexpr (HasMargin<C>::value) cl.flag |= Real<R>::abs(g) <= kParallel;
    cl.offer_nz(Real<R>::mul_nz(-h, Real<R>::rcp(g)), sc.id[slot]);
}

// Sphere::intersect, shape.hpp:78-103 (a == 1): with hb = b / 2, c = |oc|^2 - r^2, disc = hb^2 - c (an exact
// rescaling of b^2 - 4 c) the roots are t1 = -hb - sqrt(disc) <= t2 = -hb + sqrt(disc) and the reference returns t1
// if t1 > 0, else t2.  t1 > 0 <=> the origin is outside (c > 0) and the centre lies ahead (hb < 0), so the root is
// chosen from the two sign bits and ONE subtraction is made: sign logic on the high words instead of a second FP64
// add, a compare and a 64-bit select.  (Not the same only when the origin lies on the sphere to the last bit --
// c == 0 or a c below the rounding of hb^2 -- which a ray that left a surface by the 1e-3 offset of
// pathtracer.hpp:99 cannot do.)  r^2 comes from the host; the chain starts from -r^2, one FMA fewer.
template <typename R, typename C>
__device__ __forceinline__ void sphere_test(const DevScene<R>& sc, int slot, V3<R> o, V3<R> d, C& cl)
{
    const R a0 = sc.sph[slot][0], a1 = sc.sph[slot][1], a2 = sc.sph[slot][2];
    const V3<R> oc = {o.x - a0, o.y - a1, o.z - a2};
    const R hb = dot(oc, d);                       // b/2
    const R c = Real<R>::fma(oc.z, oc.z, Real<R>::fma(oc.y, oc.y, Real<R>::fma(oc.x, oc.x, -sc.sph[slot][3])));
    const R disc = Real<R>::fma(hb, hb, -c);       // (b^2 - 4c)/4
    if constexpr (HasMargin<C>::value) {
        const R a3 = sc.prim[slot][3];
        cl.flag |= Real<R>::abs(disc) < kDiscRel * a3 * a3 + cl.slack * 4.0f * (Real<R>::abs(hb) + a3);
    }
    const R sq = Real<R>::sqrt(disc);              // NaN when disc < 0: every compare below fails
    if constexpr (HasMargin<C>::value) cl.rejected(-hb - sq);   // the root that is not offered must not be a near-zero either
    cl.offer(Real<R>::neg_if_outside_ahead(sq, c, hb) - hb, sc.id[slot]);
}

// ---------------------------------------------------------------------------
// Pathtracer::raycast, pathtracer.hpp:72-89 with Plane::intersect
// (shape.hpp:49-56) and Sphere::intersect (shape.hpp:78-103, a == 1).
//
// The reference divides once per plane (t = h / dot(dir, -n)).  Here a plane
// costs a Newton reciprocal (no IEEE division), and the axis-aligned unit
// planes share three reciprocals 1 / d_a per segment.  The first kAxisFast
// axis planes per axis, kFast other planes and kFast spheres are tested by
// straight-line code entered at the first live slot (their operands are
// immediate constant-bank addresses); larger scenes continue in rolled loops
// with a warp-uniform slot.
// ---------------------------------------------------------------------------
template <typename R, bool MARGIN> struct ClosestOf { template <bool PACK> using type = Closest<R, PACK>; };
template <> struct ClosestOf<float, true> { template <bool PACK> using type = ClosestMargin; };

// The scan itself, for any running-closest type C (offer / flag interface above).
template <typename R, typename C>
__device__ __forceinline__ void scan_prims(const DevScene<R>& sc, V3<R> o, V3<R> d, C& cl)
{
    // Entry into the straight-line tests by a compare tree on the (warp-uniform)
    // first live slot: ~6 instructions, where the compiler's jump table for the
    // equivalent switch cost ~20 per entry.
    static_assert(kAxisFast == 2, "the axis windows below are written out for two slots");
#define DRTB_AXIS(AX, OA, DA)                                                                  \
    if (sc.n_aa[AX] > 0) {                                                                      \
        const R inv = Real<R>::rcp(DA);                                                         \
        if constexpr (HasMargin<C>::value) cl.flag |= Real<R>::abs(DA) <= kParallel;             \
        if (sc.n_aa[AX] > 1) axis_plane_test(sc, AX, 0, OA, inv, cl);                     \
        axis_plane_test(sc, AX, 1, OA, inv, cl);                                          \
    }
    DRTB_AXIS(0, o.x, d.x)
    DRTB_AXIS(1, o.y, d.y)
    DRTB_AXIS(2, o.z, d.z)
#undef DRTB_AXIS
#define DRTB_ENTER(first, L)                                                                   \
    if (first >= 4) { if (first >= 6) { if (first >= 7) { if (first == 7) goto L##7; goto L##8; } goto L##6; } \
                      if (first == 5) goto L##5; goto L##4; }                                   \
    if (first >= 2) { if (first == 3) goto L##3; goto L##2; }                                   \
    if (first == 1) goto L##1;
    {
        const int first = kFast - sc.n_fast_planes;
        DRTB_ENTER(first, P)
        plane_test(sc, 0, o, d, cl);
    P1: plane_test(sc, 1, o, d, cl);
    P2: plane_test(sc, 2, o, d, cl);
    P3: plane_test(sc, 3, o, d, cl);
    P4: plane_test(sc, 4, o, d, cl);
    P5: plane_test(sc, 5, o, d, cl);
    P6: plane_test(sc, 6, o, d, cl);
    P7: plane_test(sc, 7, o, d, cl);
    P8:;
    }
    for (int i = 0; i < sc.n_over_planes; ++i) plane_test(sc, 2 * kFast + i, o, d, cl);
    {
        const int first = kFast - sc.n_fast_spheres;
        DRTB_ENTER(first, S)
        sphere_test(sc, kFast + 0, o, d, cl);
    S1: sphere_test(sc, kFast + 1, o, d, cl);
    S2: sphere_test(sc, kFast + 2, o, d, cl);
    S3: sphere_test(sc, kFast + 3, o, d, cl);
    S4: sphere_test(sc, kFast + 4, o, d, cl);
    S5: sphere_test(sc, kFast + 5, o, d, cl);
    S6: sphere_test(sc, kFast + 6, o, d, cl);
    S7: sphere_test(sc, kFast + 7, o, d, cl);
    S8:;
    }
#undef DRTB_ENTER
    for (int i = 0; i < sc.n_over_spheres; ++i)
        sphere_test(sc, 2 * kFast + sc.n_over_planes + i, o, d, cl);
}

template <typename R, bool PACK = false, bool MARGIN = false>
__device__ __forceinline__ int closest_hit(const DevScene<R>& sc, V3<R> o, V3<R> d, R& tmin, bool* close_call = nullptr,
                                           bool skip_zero = false, float slack = 0.0f)
{
    if constexpr (MARGIN) {
        typename ClosestOf<R, MARGIN>::template type<PACK> cl;
        cl.slack = slack;
        scan_prims(sc, o, d, cl);
        const int k = cl.finish(tmin, skip_zero);
        *close_call = cl.flag;
        return k;
    } else {
        typename ClosestOf<R, MARGIN>::template type<PACK> cl;
        scan_prims(sc, o, d, cl);
        return cl.finish(tmin);
    }
}

// make_frame (bxdf.hpp:29-41) for a UNIT normal (spheres, triangles): with
// e the axis the reference picks, tangent = (e - n (e.n)) / sqrt(1 - (e.n)^2)
// because |e - n (e.n)|^2 = 1 - (e.n)^2 when |n| = 1, and n x tangent is then
// already a unit vector, so the reference's second normalize() is the
// identity up to rounding.  (e.n)^2 <= 1/2 by the axis choice: well conditioned.
// Plane normals may be non-unit (src/render.cpp:42); their frames come from
// the host, computed with the reference's own formula.
template <typename R>
__device__ __forceinline__ void unit_frame(V3<R> n, V3<R>& tg, V3<R>& bt)
{
    const bool ex = Real<R>::abs(n.x) < Real<R>::abs(n.y);      // e = (1,0,0) : (0,1,0)
    const R a = Real<R>::select(ex, n.x, n.y);                  // e.n
    const R s = Real<R>::fma(-a, a, R(1));                      // |e - n a|^2 = pivot component of e - n a
    const R r = Real<R>::rsqrt(s);
    const R m = -a * r;
    const R piv = s * r;
    tg = {Real<R>::select(ex, piv, n.x * m), Real<R>::select(ex, n.y * m, piv), n.z * m};
    bt = cross(n, tg);
}

// ---------------------------------------------------------------------------
// DiffuseBxDF::sample (bxdf.hpp:69-79) + angle_to_dir (:43-52) on the frame
// (tg, bt, n), then cos = dot(n, dir_out) (pathtracer.hpp:103).  (sp, cp) =
// (sin, cos)(phi), phi = 2 pi u_phi, come from the caller.  Returns
// w = cos / pdf.  n is used RAW (the non-unit green-wall normal stays non-unit).
// sin(asin(sqrt u)) = sqrt u, cos(asin(sqrt u)) = sqrt(1 - u); u < 1 always, so
// one rsqrt(1 - u) yields both cos(theta) and the 1/cos(theta) that pdf needs.
// With dir_out = x tg + y bt + cos(theta) n and bt = n x tg orthogonal to n,
//     w = pi |n|^2 + pi (n.tg) x / cos(theta),
// and n.tg = 0 whenever n is a unit vector or orthogonal to make_frame's helper
// axis (every primitive of the Cornell box, raw green-wall normal included; all
// spheres and triangles): w is then a CONSTANT of the primitive.  The all-diffuse
// kernels exploit that: their records carry no weights at all (PathRecord), the
// sweeps take w from DevScene::wtab, and a scene with a tilted non-unit plane
// normal renders with the general kernels, which evaluate the reference's dot
// product and division as below.
// ---------------------------------------------------------------------------
// The direction alone (records without weights, see PathRecord): w is then the primitive's table entry.
template <typename R>
__device__ __forceinline__ V3<R> diffuse_direction(V3<R> n, V3<R> tg, V3<R> bt, R u_theta, R sp, R cp)
{
    const R st = Real<R>::sqrt(u_theta);
    const R ct = Real<R>::sqrt(R(1) - u_theta);    // u < 1 always
    const R x = cp * st, y = sp * st;
    return {x * tg.x + y * bt.x + ct * n.x,
            x * tg.y + y * bt.y + ct * n.y,
            x * tg.z + y * bt.z + ct * n.z};
}

template <typename R>
__device__ __forceinline__ V3<R> diffuse_sample(V3<R> n, V3<R> tg, V3<R> bt, R u_theta, R sp, R cp, R& w)
{
    const R st = Real<R>::sqrt(u_theta);
    const R om = R(1) - u_theta;
    const R inv_ct = Real<R>::rsqrt(om);
    const R ct = om * inv_ct;
    const R x = cp * st, y = sp * st;
    const V3<R> dout = {x * tg.x + y * bt.x + ct * n.x,
                        x * tg.y + y * bt.y + ct * n.y,
                        x * tg.z + y * bt.z + ct * n.z};
    w = dot(n, dout) * (Real<R>::pi() * inv_ct);       // dot(n, dout) / (cos(theta) / pi)
    return dout;
}

// Normal and tangent frame of analytic primitive k at the hit point pt.  Planes: constants of the primitive (the
// frame is make_frame(n), bxdf.hpp:29-41, evaluated on the host).  Spheres: normal(point) = normalize(point -
// centre), shape.hpp:105-106, and make_frame of that unit normal.  (The division by the norm cannot be replaced
// by 1 / r: Sphere::intersect takes a == 1 even for the non-unit directions that leave the raw green-wall normal,
// shape.hpp:83, so such a "hit point" is not on the sphere -- measured: 4 of 12 288 paths of a 48 x 32 Cornell
// box take another course.)
template <typename R>
__device__ __forceinline__ bool analytic_frame(const BlockScene<R>& bs, int k, V3<R> pt, V3<R>& nrm, V3<R>& tg, V3<R>& bt)
{
    nrm = {bs.prim[k][0], bs.prim[k][1], bs.prim[k][2]};
    const bool sphere = bs.type[k] == DRTB_SPHERE;
    if (sphere) {
        nrm = normalize(V3<R>{pt.x - nrm.x, pt.y - nrm.y, pt.z - nrm.z});
        unit_frame(nrm, tg, bt);
    } else {
        tg = {bs.frame[k][0], bs.frame[k][1], bs.frame[k][2]};
        bt = {bs.frame[k][3], bs.frame[k][4], bs.frame[k][5]};
    }
    return sphere;
}

// ---------------------------------------------------------------------------
// SpecularBxDF (bxdf.hpp:85-124): sample() draws a half vector around the
// normal from the two uniforms (theta = acos(sqrt(u^(2/(e+2)))), phi = 2 pi u),
// flips it to the side of dir_in = -d with reflect(h, n) (vector.hpp:602-606),
// and returns reflect(dir_in, h) with pdf = (e+2)/(2 pi) cos^(e+1)(theta)
// sin(theta); operator() re-derives the half vector from (dir_in, dir_out) and
// returns color * (e+2)/(2 pi) (n.h)^e sqrt(1 - (n.h)^2).  The scatter term of
// pathtracer.hpp:100-104 is brdf * L * dot(n, dir_out) / pdf, so the record
// keeps w = pi * lobe * dot(n, dir_out) / pdf and the sweeps, which multiply by
// rho / pi, need not know which BxDF made the vertex.  (e+2)/(2 pi) is common to
// lobe and pdf and is cancelled.  cos(acos x) = x, sin(acos x) = sqrt(1 - x^2).
// Nothing is clamped: a direction below the surface keeps its negative cosine
// and a lobe argument above 1 (non-unit plane normal) yields NaN, as upstream.
// ---------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ V3<R> specular_sample(V3<R> n, V3<R> tg, V3<R> bt, V3<R> d, R expo, R u_theta, R sp, R cp, R& w)
{
    const R ct2 = Real<R>::pow(u_theta, Real<R>::div(R(2), expo + R(2)));
    const R ct = Real<R>::sqrt(ct2), st = Real<R>::sqrt(R(1) - ct2);
    const R x = cp * st, y = sp * st;
    V3<R> h = {x * tg.x + y * bt.x + ct * n.x, x * tg.y + y * bt.y + ct * n.y, x * tg.z + y * bt.z + ct * n.z};
    const V3<R> din = {-d.x, -d.y, -d.z};                              // pathtracer.hpp:101, 109
    if (dot(h, din) < R(0)) {                                          // bxdf.hpp:114-115
        const R k2 = R(2) * dot(n, h);
        h = {k2 * n.x - h.x, k2 * n.y - h.y, k2 * n.z - h.z};
    }
    const R k2 = R(2) * dot(h, din);
    const V3<R> dout = {k2 * h.x - din.x, k2 * h.y - din.y, k2 * h.z - din.z};   // reflect(dir_in, h), :116
    const R pdf = Real<R>::pow(ct, expo + R(1)) * st;                  // x (e+2)/(2 pi), :117-118
    const V3<R> hw = normalize(V3<R>{din.x + dout.x, din.y + dout.y, din.z + dout.z});   // :98
    const R c = dot(n, hw);
    const R lobe = Real<R>::pow(c, expo) * Real<R>::sqrt(Real<R>::fma(-c, c, R(1)));      // x (e+2)/(2 pi), :99-103
    w = Real<R>::div(Real<R>::pi() * lobe * dot(n, dout), pdf);
    return dout;
}

// Per-path vertex record: what the reference keeps as ~17 heap-allocated tape
// nodes per segment (vector.hpp:194-213) shrinks to (prim id, w) per vertex;
// p_v is a function of the depth alone (pathtracer.hpp:130).  Scene indices fit
// a byte for analytic scenes, 32 bits once a mesh is attached.
template <bool MESH> struct PrimId { using type = uint8_t; };
template <> struct PrimId<true> { using type = int32_t; };

// HASW = false (the all-diffuse analytic kernels): the weights are constants of the primitives (diffuse_sample), so
// the record is the list of primitives alone -- no weight is computed, stored, queued or re-read, the lit-path ring
// shrinks from 9 to 1 byte per vertex (and fits shared memory at 7 resident blocks), the local frame by a third.
template <typename R, bool MESH, int CAP, bool HASW = true>
struct PathRecord {
    using Id = typename PrimId<MESH>::type;
    static constexpr int kCap = CAP;     // kQueueDepth for the compacting kernels, kMaxDepth otherwise: the
    static constexpr bool kHasW = HASW;  // per-thread local-memory footprint has to stay inside L2
    R  w_[HASW ? CAP : 1];
    Id prim_[CAP];
    __device__ __forceinline__ R w(int v) const { return w_[v]; }
    __device__ __forceinline__ int prim(int v) const { return int(prim_[v]); }
};

// Lit-path compaction.  Only ~16 % of the Cornell box's paths reach the light,
// so running the sweeps right after each trace keeps ~5 of 32 lanes busy.
// Instead every lit lane appends its record to a per-warp ring in shared
// memory (SoA, slot-contiguous => conflict-free) and the warp runs the sweeps
// on 32 queued records at a time.  Ballot-ordered, hence deterministic.
constexpr int kQueueSlots = 64;    // ring capacity per warp (power of two)
constexpr int kQueueDepth = 16;    // deepest record the ring stores; deeper runs use the direct path

template <typename R, bool MESH, int CAP = kQueueDepth, bool HASW = true>
struct QueueView {                 // one record of one warp's ring
    using Id = typename PrimId<MESH>::type;
    static constexpr int kCap = CAP;   // bounds the L_{v+1} array of the sweeps (radiance_and_adjoint)
    static constexpr bool kHasW = HASW;
    const R* w_;                   // &ring_w[slot], stride kQueueSlots
    const Id* prim_;
    __device__ __forceinline__ R w(int v) const { return w_[v * kQueueSlots]; }
    __device__ __forceinline__ int prim(int v) const { return int(prim_[v * kQueueSlots]); }
};

__host__ __device__ constexpr size_t queue_bytes_per_warp(int depth, size_t real_size, size_t id_size)
{
    // w[depth][slots] | prim[depth][slots] | n[slots], each part 8-byte aligned
    return size_t(depth) * kQueueSlots * real_size + size_t(depth) * kQueueSlots * id_size + kQueueSlots;
}

// Who emits, who scatters, and the parameter values, by scene index.  Analytic
// primitives answer from the block's shared copy; triangles (scene index >=
// n_prims) and large parameter sets answer from global memory through the
// read-only path.
template <typename R, bool MESH> struct Materials;
template <typename R> struct Materials<R, false> {
    const BlockScene<R>* bs;
    __device__ __forceinline__ int2 em_col(int k) const { return bs->em_col[k]; }
    __device__ __forceinline__ R w(int k) const { return bs->wtab[k]; }
    __device__ __forceinline__ R param(int i) const { return bs->param[i]; }
};
template <typename R> struct Materials<R, true> {
    const BlockScene<R>* bs;
    MeshView mesh;
    const double* params;
    __device__ __forceinline__ int2 em_col(int k) const
    {
        return k < mesh.n_prims ? bs->em_col[k] : make_int2(__ldg(mesh.emis + (k - mesh.n_prims)), __ldg(mesh.color + (k - mesh.n_prims)));
    }
    __device__ __forceinline__ R w(int k) const { return k < mesh.n_prims ? bs->wtab[k] : Real<R>::pi(); }
    __device__ __forceinline__ R param(int i) const { return R(__ldg(params + i)); }
};

struct TraceCounters { uint32_t segments = 0, truncated = 0, bvh_nodes = 0, tri_tests = 0, close_call = 0; };

// ---------------------------------------------------------------------------
// Pathtracer::trace + scatter (pathtracer.hpp:91-136), recursion unrolled into
// a loop that only records vertices.  Returns the vertex count; `lit` tells
// whether any vertex carries an emitter (otherwise radiance and every
// gradient of the path are exactly zero and the sweeps are skipped).
// `slot` is the next stream slot (2 after the camera draws).
// ---------------------------------------------------------------------------
// SPEC: the scene has SpecularBxDF materials (per-lane material lookup and the
// lobe code are compiled in; the all-diffuse kernels do not carry them).
// MIXED (float, analytic, all-diffuse): every closest hit is checked for a close call (ClosestMargin); the first one
// ends the trace with cnt.close_call = 1 and the caller hands the path to the double re-trace.
template <typename R, bool MESH, int CAP, bool SPEC = false, bool MIXED = false, bool HASW = true>
__device__ __forceinline__ int trace_path(const DevScene<R>& sc, const BlockScene<R>& bs,
                                          const Materials<R, MESH>& mat, bool no_bvh,
                                          uint64_t base, uint32_t slot, V3<R> o, V3<R> d,
                                          int min_bounces, double absorb, int max_depth,
                                          PathRecord<R, MESH, CAP, HASW>& rec, bool& lit, TraceCounters& cnt)
{
    using Id = typename PrimId<MESH>::type;
    int n = 0;
    lit = false;
#ifndef DRTB_SYNC_DEPTH
#define DRTB_SYNC_DEPTH 0                                   // 1: re-converge the analytic kernels per segment as well
#endif
    constexpr bool kSync = MESH || DRTB_SYNC_DEPTH;         // mesh kernels: the lanes must enter the BVH traversal together
    unsigned live = kSync ? __activemask() : 0u;
    bool alive = true;
    float e_pos = 0.f, e_dir = kRoundDir;                   // MIXED: bound on the float path's distance from the double one
    uint64_t ctr = base + kGolden + slot;                  // splitmix64's increment folded in (rng.cuh)
    for (int depth = 0;; ++depth) {
        if constexpr (kSync) live = __ballot_sync(live, alive);
        if (!alive) break;
        if (depth >= min_bounces) {                         // Russian roulette, :128-130
            // every draw is < 1 (k <= M - 1), so absorb >= 1 ends the path whatever the draw says
            if (absorb >= 1.0) { alive = false; continue; }
            double u = Real<double>::uniform(stream_draw_ctr(ctr++));
            if (u < absorb) { alive = false; continue; }
        }
        if (n >= max_depth) { ++cnt.truncated; alive = false; continue; }
        R t;
        int k;
        if constexpr (MIXED) {
            bool close_call;
            k = closest_hit<R, false, true>(sc, o, d, t, &close_call, depth == 0 && sc.eye_clear != 0, kDriftK * (e_pos + 4.0f * e_dir));
            if (close_call) { cnt.close_call = 1; alive = false; continue; }
        } else {
            k = closest_hit<R, kPackHit && !MESH>(sc, o, d, t);   // analytic primitives
        }
        int tri = -1;
        if constexpr (MESH) {                               // then the mesh; analytic wins exact ties
            if (k < 0) t = Real<R>::inf();
            if (no_bvh) brute_closest(mat.mesh, o, d, t, tri, cnt.tri_tests);
            else        bvh_closest(mat.mesh, o, d, t, tri, cnt.bvh_nodes, cnt.tri_tests);
            if (tri >= 0) k = mat.mesh.n_prims + tri;
        }
        ++cnt.segments;
        if (k < 0) { alive = false; continue; }             // miss, :134-135
        V3<R> pt = {o.x + t * d.x, o.y + t * d.y, o.z + t * d.z};
        const int2 ec = mat.em_col(k);
        const int em = ec.x, col = ec.y;
        lit |= em >= 0;
        rec.prim_[n] = Id(k);
        // absorb >= 1: the next trace() call returns 0 before it looks at the ray (:128-129), so this
        // vertex's scattered term is brdf * 0 * cos / pdf and its w meets only L_{v+1} = 0 in both
        // sweeps.  For the diffuse BxDF w is always finite, hence unobservable: the two draws, the
        // frame and the direction are skipped (warp-uniform test).  SpecularBxDF keeps them, its w can
        // be NaN and NaN * 0 poisons the path upstream.
        const bool last_vertex = DRTB_SKIP_LAST && !SPEC && absorb >= 1.0 && depth + 1 >= min_bounces;
        static_assert(HASW || (!SPEC && !MESH), "records without weights: all-diffuse analytic scenes only");
        if (col < 0 || last_vertex) {                       // null BxDF, :25-26, 38-39
            if constexpr (HASW) rec.w_[n] = R(0);           // (without weights: this vertex's w only ever meets L_{v+1} = 0)
            ++n;
            alive = false;
            continue;
        }
        V3<R> nrm, tg, bt;
        bool on_mesh = false;
        if constexpr (MESH) {
            if (tri >= 0) {                                 // unit geometric normal, drtb.h
                const TriData<R> T = load_tri<R>(mat.mesh, tri);
                nrm = normalize(cross(T.e1, T.e2));
                unit_frame(nrm, tg, bt);
                on_mesh = true;
            }
        }
        if (!on_mesh) {
            const bool sphere = analytic_frame(bs, k, pt, nrm, tg, bt);
            if constexpr (MIXED) {                          // the frame's axis switch and the drift bound (see ClosestMargin)
                const float cosi = fmaxf(fabsf(float(dot(d, nrm))), 0.05f);
                const float e_hit = (e_pos + float(t) * e_dir + kRoundPos * (1.0f + float(t))) * Real<float>::rcp(cosi);
                const bool axis_call = sphere && fabsf(fabsf(float(nrm.x)) - fabsf(float(nrm.y))) < kFrameGap + kDriftK * e_hit;
                if (axis_call || e_hit > kDriftMax) { cnt.close_call = 1; alive = false; continue; }
                e_pos = e_hit;
                e_dir = sphere ? e_hit * Real<float>::rcp(float(bs.prim[k][3])) + kRoundDir : kRoundDir;
            }
        }
        R u_theta = Real<R>::uniform_fast(stream_draw_ctr(ctr));
        R sp, cp;                                           // phi = 2 * pi * uniform(), bxdf.hpp:74
        Real<R>::sincos_tab(bs.tab, stream_draw_ctr(ctr + 1), &sp, &cp);
        ctr += 2;
        R w;
        V3<R> dout;
        bool spec = false;
        if constexpr (SPEC) spec = !on_mesh && bs.mtype[k] == DRTB_SPECULAR;
        if (spec) {
            dout = specular_sample(nrm, tg, bt, d, bs.expo[k], u_theta, sp, cp, w);
            // upstream a NaN/inf weight poisons the path even if it never meets the light (NaN * 0):
            // such a path must run the sweeps, which then produce the reference's NaN
            lit |= !(Real<R>::abs(w) < Real<R>::inf());
        } else if constexpr (HASW) {
            dout = diffuse_sample(nrm, tg, bt, u_theta, sp, cp, w);
        } else {
            dout = diffuse_direction(nrm, tg, bt, u_theta, sp, cp);
        }
        if constexpr (HASW) rec.w_[n] = w;
        ++n;
        const R eps = Real<R>::origin_eps();                // 1e-3, pathtracer.hpp:99
        o = {Real<R>::fma(eps, dout.x, pt.x), Real<R>::fma(eps, dout.y, pt.y), Real<R>::fma(eps, dout.z, pt.z)};
        d = dout;
    }
    return n;
}

// ---------------------------------------------------------------------------
// ONE segment of trace_path for one lane whose path state lives in the caller
// (analytic scenes, diffuse BxDFs): the regenerating kernel
// (render_regen_kernel) steps every lane once per iteration and hands a fresh camera
// sample to lanes whose path has ended.  Same arithmetic, same draws, same
// record as trace_path; the roulette of the NEXT trace() call is decided here
// as well.  Returns true when the path ends with this call.
// ---------------------------------------------------------------------------
// Russian roulette of the trace() call at `depth` (pathtracer.hpp:128-130).  The regenerating
// kernel asks right after the segment that leads there (and once for the camera ray), so that a
// lane never spends an iteration only to learn that its path was absorbed; the draw keeps its
// place in the path's stream (after the two sampling draws of the previous vertex).
__device__ __forceinline__ bool roulette_absorbs(uint64_t& ctr, int depth, int min_bounces, double absorb)
{
    if (depth < min_bounces) return false;
    if (absorb >= 1.0) return true;
    return Real<double>::uniform(stream_draw_ctr(ctr++)) < absorb;
}

template <typename R, int CAP, bool HASW>
__device__ __forceinline__ bool trace_segment(const DevScene<R>& sc, const BlockScene<R>& bs,
                                              const Materials<R, false>& mat, uint64_t& ctr, V3<R>& o, V3<R>& d,
                                              int& depth, int& n, bool& lit, int min_bounces, double absorb,
                                              int max_depth, PathRecord<R, false, CAP, HASW>& rec, TraceCounters& cnt)
{
    if (n >= max_depth) { ++cnt.truncated; return true; }
    R t;
    const int k = closest_hit<R, kPackHit>(sc, o, d, t);
    ++cnt.segments;
    if (k < 0) return true;                                 // miss, :134-135
    const V3<R> pt = {o.x + t * d.x, o.y + t * d.y, o.z + t * d.z};
    const int2 ec = mat.em_col(k);
        const int em = ec.x, col = ec.y;
    lit |= em >= 0;
    rec.prim_[n] = uint8_t(k);
    if (col < 0) {                                          // null BxDF, :25-26, 38-39
        if constexpr (HASW) rec.w_[n] = R(0);
        ++n;
        return true;
    }
    V3<R> nrm, tg, bt;
    R w;
    analytic_frame(bs, k, pt, nrm, tg, bt);
    const R u_theta = Real<R>::uniform_fast(stream_draw_ctr(ctr));
    R sp, cp;                                               // phi = 2 * pi * uniform(), bxdf.hpp:74
    Real<R>::sincos_tab(bs.tab, stream_draw_ctr(ctr + 1), &sp, &cp);
    ctr += 2;
    V3<R> dout;
    if constexpr (HASW) { dout = diffuse_sample(nrm, tg, bt, u_theta, sp, cp, w); rec.w_[n] = w; }
    else                dout = diffuse_direction(nrm, tg, bt, u_theta, sp, cp);
    ++n;
    const R eps = Real<R>::origin_eps();                    // 1e-3, pathtracer.hpp:99
    o = {Real<R>::fma(eps, dout.x, pt.x), Real<R>::fma(eps, dout.y, pt.y), Real<R>::fma(eps, dout.z, pt.z)};
    d = dout;
    ++depth;
    return roulette_absorbs(ctr, depth, min_bounces, absorb);
}

// ---------------------------------------------------------------------------
// Radiance recurrence + adjoint: replaces the reverse tape (vector.hpp:120-318,
// 418-557).  Backward sweep  L_v = (E_v + (rho_v/pi) L_{v+1} w_v) / p_v  gives
// the path radiance L_0; forward sweep with g_0 = seed:
//   gp = g_v / p_v ; grad[E_v] += gp ; grad[rho_v] += gp w_v L_{v+1} / pi ;
//   g_{v+1} = gp w_v rho_v / pi.           (per channel; channels never mix)
// Sink::add(param_index, channel, value) receives the contributions.
// ---------------------------------------------------------------------------
// The same for a record deeper than kQueueDepth, without the L_{v+1} array: the forward sweep
// re-runs the backward recurrence from the end of the path down to v + 1 for every vertex
// (O(n^2), n <= 64, rare) -- the same operations in the same order, hence the same bits.
template <typename R, typename Mat, typename Rec, typename Sink>
__device__ __noinline__ void radiance_and_adjoint_deep(const Mat& mat, const Rec& rec, int n, int min_bounces, R inv_p,
                                                       bool want_grad, const R g0[3], R L0[3], Sink& sink)
{
    auto radiance_from = [&](int first, R L[3]) {           // L_first, sweeping v = n - 1 .. first
        L[0] = L[1] = L[2] = R(0);
        for (int v = n - 1; v >= first; --v) {
            const int k = rec.prim(v);
            const int2 ec = mat.em_col(k);
        const int em = ec.x, col = ec.y;
            const R ip = v >= min_bounces ? inv_p : R(1);
            R wv;
            if constexpr (Rec::kHasW) wv = rec.w(v); else wv = mat.w(k);
            const R f = wv * Real<R>::inv_pi();
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                R E = em >= 0 ? mat.param(3 * em + c) : R(0);
                R rho = col >= 0 ? mat.param(3 * col + c) : R(0);
                L[c] = (E + rho * f * L[c]) * ip;
            }
        }
    };
    radiance_from(0, L0);
    if (!want_grad) return;
    R g[3] = {g0[0], g0[1], g0[2]};
    for (int v = 0; v < n; ++v) {
        const int k = rec.prim(v);
        const int2 ec = mat.em_col(k);
        const int em = ec.x, col = ec.y;
        const R ip = v >= min_bounces ? inv_p : R(1);
        R wv;
            if constexpr (Rec::kHasW) wv = rec.w(v); else wv = mat.w(k);
            const R f = wv * Real<R>::inv_pi();
        R Ln[3];
        radiance_from(v + 1, Ln);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            R gp = g[c] * ip;
            if (em >= 0) sink.add(em, c, gp);
            if (col >= 0) {
                sink.add(col, c, gp * f * Ln[c]);
                g[c] = gp * f * mat.param(3 * col + c);
            } else {
                g[c] = R(0);
            }
        }
    }
}

template <typename R, typename Mat, typename Rec, typename Sink>
__device__ __forceinline__ void radiance_and_adjoint(const Mat& mat, const Rec& rec,
                                                     int n, int min_bounces, R inv_p,
                                                     bool want_grad, const R g0[3], R L0[3], Sink& sink)
{
    // L_{v+1} of every vertex is kept for the forward sweep -- for up to kQueueDepth vertices.  Deeper records
    // (capacity kMaxDepth; p ~ 1e-5 at the reference's defaults) recompute it instead, see below: a 65 x 3 array
    // per thread would triple the local memory the driver has to reserve for every resident thread.
    constexpr int kKeep = Rec::kCap < kQueueDepth ? Rec::kCap : kQueueDepth;
    if constexpr (Rec::kCap > kQueueDepth) {
        if (n > kQueueDepth) {
            radiance_and_adjoint_deep(mat, rec, n, min_bounces, inv_p, want_grad, g0, L0, sink);
            return;
        }
    }
    R Ls[kKeep + 1][3];
    R L[3] = {R(0), R(0), R(0)};
    for (int v = n - 1; v >= 0; --v) {
        const int k = rec.prim(v);
        const int2 ec = mat.em_col(k);
        const int em = ec.x, col = ec.y;
        const R ip = v >= min_bounces ? inv_p : R(1);
        R wv;
            if constexpr (Rec::kHasW) wv = rec.w(v); else wv = mat.w(k);
            const R f = wv * Real<R>::inv_pi();
        if (want_grad) { Ls[v + 1][0] = L[0]; Ls[v + 1][1] = L[1]; Ls[v + 1][2] = L[2]; }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            R E = em >= 0 ? mat.param(3 * em + c) : R(0);
            R rho = col >= 0 ? mat.param(3 * col + c) : R(0);
            L[c] = (E + rho * f * L[c]) * ip;
        }
    }
    L0[0] = L[0]; L0[1] = L[1]; L0[2] = L[2];
    if (!want_grad) return;
    R g[3] = {g0[0], g0[1], g0[2]};
    for (int v = 0; v < n; ++v) {
        const int k = rec.prim(v);
        const int2 ec = mat.em_col(k);
        const int em = ec.x, col = ec.y;
        const R ip = v >= min_bounces ? inv_p : R(1);
        R wv;
            if constexpr (Rec::kHasW) wv = rec.w(v); else wv = mat.w(k);
            const R f = wv * Real<R>::inv_pi();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            R gp = g[c] * ip;
            if (em >= 0) sink.add(em, c, gp);
            if (col >= 0) {
                sink.add(col, c, gp * f * Ls[v + 1][c]);
                g[c] = gp * f * mat.param(3 * col + c);
            } else {
                g[c] = R(0);
            }
        }
    }
}

} // namespace drtb
