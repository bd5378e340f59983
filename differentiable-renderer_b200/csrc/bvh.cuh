// bvh.cuh — triangle meshes: GPU-built 4-wide BVH and its traversal.
//
// NEW functionality (the reference has planes and spheres only, scanned
// linearly: pathtracer.hpp:72-89).  Semantics are fixed in include/drtb.h so
// that a BVH traversal returns exactly what the reference's linear scan would
// return if it had a Triangle shape: closest t > 0, exact ties to the lower
// scene index.
//
// Build, all on the device:
//   1. per-triangle bounds + centroid -> 63-bit Morton code -> cub radix sort
//   2. binary tree over the sorted leaves, either
//        PLOC  (parallel locally-ordered clustering, Meister & Bittner 2018):
//              every cluster looks kPlocRadius neighbours left and right along
//              the Morton order for the partner with the smallest merged
//              surface area, mutual nearest neighbours merge, repeat -- an
//              agglomerative build whose SAH cost is far below LBVH's; or
//        LBVH  (Karras 2012): binary radix tree over the codes + bottom-up refit
//              (kept as the fast-build option, DRTB_BVH=lbvh)
//   3. collapse to an 8-wide COMPRESSED BVH (after Ylitie, Karras, Laine 2017, "Efficient
//      incoherent ray traversal on GPUs through compressed wide BVHs"), top-down and
//      level-synchronous: a wide node adopts the descendants with the largest surface
//      area first until it has 8 children; subtrees of <= kLeafMax triangles become
//      leaves.  A node's internal children are stored contiguously (child base + rank in
//      the internal mask) and so are the triangles of its leaf children, in leaf order.
// Wide node = 128 B (one cache line): the node's box origin (3 floats), a power-of-two scale
// per axis (3 exponent bytes), the internal-child mask, two base indices, 8 meta bytes and
// the 8 child boxes quantised to 15 bits per plane relative to the origin (the paper's 8-bit
// planes make an 80-byte node; a 16-bit field drops into a float's mantissa with ONE byte
// permute, so a plane costs PRMT + FFMA instead of PRMT + FADD + FFMA, and an aligned
// 128-byte node touches no more 32-byte sectors than an unaligned 80-byte one).
// Quantisation rounds outward, the leaf boxes are padded and the slab test is widened, so
// float culling can never reject a triangle the exact test would accept.  Against the
// 4-wide float nodes of round 1: 8 children per 128 B instead of 4, no sorting network, and
// at most ONE 8-byte stack entry per visited node (a child GROUP: base, hit mask, and a
// lower bound on the entry distance of what is left in it) instead of up to three.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "real.cuh"

namespace drtb {

constexpr int kBvhStack   = 64;           // traversal stack entries per lane (overflow falls back to a linear scan)
#ifndef DRTB_LEAF_MAX
#define DRTB_LEAF_MAX 3
#endif
constexpr int kLeafMax    = DRTB_LEAF_MAX; // triangles per leaf, <= 3 (3 unary bits of the child's meta byte)
#ifndef DRTB_PLOC_RADIUS
#define DRTB_PLOC_RADIUS 16
#endif
constexpr int kPlocRadius = DRTB_PLOC_RADIUS;           // PLOC neighbour search radius along the Morton order
constexpr int kTri64Stride = 10;          // doubles per triangle: v0, e1, e2, pad (16-byte aligned rows)
constexpr int kTri32Stride = 4;           // float4 per triangle (64 B: two 256-bit loads)
constexpr int kNodeStride  = 8;           // 16-byte words per wide node (128 B)

struct MeshView {
    const float4*  nodes;                 // kNodeStride 16-byte words per wide node (raw bits); node 0 is the root
    const double*  tri64;                 // ORIGINAL order: v0.xyz e1.xyz e2.xyz pad
    const float4*  tri32;                 // LEAF order: (v0.xyz,e1.x) (e1.yz,e2.xy) (e2.z, max|e1|, max|e2|, original index)
    const int32_t* color;                 // per triangle (original order): param index of the albedo, -1 = null BxDF
    const int32_t* emis;                  // per triangle (original order): param index of the emission, -1 = none
    int32_t n_tris;
    int32_t n_prims;                      // scene index of triangle 0
};

// ---------------------------------------------------------------------------
// build kernels
// ---------------------------------------------------------------------------
#ifdef DRTB_BVH_BUILD_KERNELS        // mesh.cu only: the other translation units need the traversal below, not the build

// order-preserving float <-> uint map for atomicMin/atomicMax on floats
__device__ __forceinline__ uint32_t float_to_ordered(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// per triangle: edge form in double, outward-rounded float AABB; scene bounds
static __global__ void mesh_prepare_kernel(const double* __restrict__ vertices, const int32_t* __restrict__ indices, int n,
                                    double* __restrict__ tri64, float4* __restrict__ leaf_lo, float4* __restrict__ leaf_hi,
                                    uint32_t* __restrict__ bounds /* [6] ordered: lo xyz, hi xyz */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        double v[3][3];
        for (int c = 0; c < 3; ++c) {
            const double* p = vertices + 3ll * indices[3ll * i + c];
            v[c][0] = p[0]; v[c][1] = p[1]; v[c][2] = p[2];
        }
        double* t = tri64 + (size_t)i * kTri64Stride;
        for (int a = 0; a < 3; ++a) {
            t[a] = v[0][a]; t[3 + a] = v[1][a] - v[0][a]; t[6 + a] = v[2][a] - v[0][a];
            const double mn = fmin(v[0][a], fmin(v[1][a], v[2][a])), mx = fmax(v[0][a], fmax(v[1][a], v[2][a]));
            lo[a] = __double2float_rd(mn); hi[a] = __double2float_ru(mx);
        }
        t[9] = 0.0;
        leaf_lo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        leaf_hi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    for (int a = 0; a < 3; ++a) {
        float l = lo[a], h = hi[a];
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bounds + a, float_to_ordered(l));
            atomicMax(bounds + 3 + a, float_to_ordered(h));
        }
    }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v)      // 21 bits -> every third bit of 63
{
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

static __global__ void mesh_morton_kernel(const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                                   const uint32_t* __restrict__ bounds, int n, uint64_t* __restrict__ keys,
                                   uint32_t* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 l = leaf_lo[i], h = leaf_hi[i];
    const float cl[3] = {l.x, l.y, l.z}, ch[3] = {h.x, h.y, h.z};
    uint64_t code = 0;
    for (int a = 0; a < 3; ++a) {
        const float lo = ordered_to_float(bounds[a]), hi = ordered_to_float(bounds[3 + a]);
        const float c = 0.5f * (cl[a] + ch[a]);
        const float ext = hi - lo;
        float u = ext > 0.f ? (c - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        const uint32_t q = min(uint32_t(u * 2097152.0f), 2097151u);
        code |= spread21(q) << (2 - a);
    }
    keys[i] = code;
    vals[i] = uint32_t(i);
}

// Binary-tree node ids shared by both builders: id < n is the leaf at SORTED
// position id, id >= n is internal node id - n.  Per node: box (lo.xyz | hi.xyz),
// lo.w = triangle count (as int bits).
struct BinTree {
    float4* lo;                           // [2n - 1]
    float4* hi;                           // [2n - 1]
    int2*   children;                     // [n - 1], indexed by id - n
};

__device__ __forceinline__ float box_area(float4 lo, float4 hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// leaves of the binary tree, in Morton order, boxes padded for the float slab test
static __global__ void bin_leaves_kernel(const uint32_t* __restrict__ sorted_tri, const float4* __restrict__ leaf_lo,
                                  const float4* __restrict__ leaf_hi, const uint32_t* __restrict__ bounds, int n,
                                  BinTree t)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    // the scale of the mesh's coordinates: its extent or its distance from the origin, whichever is larger
    float ext = 0.f;
    for (int a = 0; a < 3; ++a) {
        const float lo = ordered_to_float(bounds[a]), hi = ordered_to_float(bounds[3 + a]);
        ext = fmaxf(ext, fmaxf(hi - lo, fmaxf(fabsf(lo), fabsf(hi))));
    }
    // covers the float rounding of the ray and of the slab test for ray origins within a few `ext` of the mesh (the
    // slab distance of axis a is off by <= 2^-21 (|o_a| + ext) / |d_a|); node8_step's relative widening covers the rest
    const float pad = 4e-6f * ext + FLT_MIN;
    const uint32_t tri = sorted_tri[p];
    const float4 l = leaf_lo[tri], h = leaf_hi[tri];
    t.lo[p] = make_float4(l.x - pad, l.y - pad, l.z - pad, __int_as_float(1));
    t.hi[p] = make_float4(h.x + pad, h.y + pad, h.z + pad, 0.f);
}

// ---- LBVH (Karras 2012) ----------------------------------------------------
// longest common prefix of sorted keys i and j; equal keys fall back to the index
__device__ __forceinline__ int lbvh_delta(const uint64_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(uint32_t(i) ^ uint32_t(j)) : __clzll((long long)(a ^ b));
}

static __global__ void lbvh_hierarchy_kernel(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children,
                                      int* __restrict__ parent /* [2n-1] by node id */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = lo == gamma ? gamma : n + gamma;              // leaf id : internal id
    const int right = hi == gamma + 1 ? gamma + 1 : n + gamma + 1;
    children[i] = make_int2(left, right);
    parent[left] = n + i;
    parent[right] = n + i;
    if (i == 0) parent[n] = -1;
}

// one thread per leaf climbs; the second arrival at a node owns it
static __global__ void lbvh_refit_kernel(int n, const int* __restrict__ parent, int* __restrict__ arrivals, BinTree t)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int node = parent[p];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(arrivals + (node - n), 1) == 0) return;
        const int2 ch = t.children[node - n];
        const volatile float4* vlo = t.lo; const volatile float4* vhi = t.hi;
        const float4 al = {vlo[ch.x].x, vlo[ch.x].y, vlo[ch.x].z, vlo[ch.x].w};
        const float4 bl = {vlo[ch.y].x, vlo[ch.y].y, vlo[ch.y].z, vlo[ch.y].w};
        const float4 ah = {vhi[ch.x].x, vhi[ch.x].y, vhi[ch.x].z, 0.f};
        const float4 bh = {vhi[ch.y].x, vhi[ch.y].y, vhi[ch.y].z, 0.f};
        t.lo[node] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z),
                                 __int_as_float(__float_as_int(al.w) + __float_as_int(bl.w)));
        t.hi[node] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
        node = parent[node];
    }
}

// ---- PLOC -------------------------------------------------------------------
// clusters[i] = node id of the i-th live cluster, in Morton order.
// nearest[i] = the j in [i - R, i + R] \ {i} minimising area(box_i U box_j); ties -> smaller j.
// The build loops are DEVICE-paced: the live cluster count m, the nodes made so far and the collapse's task count
// live in a BuildState in device memory, every kernel of an iteration reads them and exits at once when there is
// nothing (left) for its block to do, and a one-thread kernel advances the state between iterations.  The host
// launches a fixed schedule of iterations (grids sized for an upper bound of the work) and reads the state back
// once per batch, not once per iteration.
struct BuildState { int m, made, n_tasks, error; };

static __global__ void __launch_bounds__(256)
ploc_nearest_kernel(const int* __restrict__ clusters, const BuildState* __restrict__ state, BinTree t, int* __restrict__ nearest)
{
    constexpr int R = kPlocRadius, T = 256;
    __shared__ float4 s_lo[T + 2 * R], s_hi[T + 2 * R];
    const int m = state->m;
    if (m <= 1 || blockIdx.x * T >= m) return;
    const int base = blockIdx.x * T - R;
    for (int k = threadIdx.x; k < T + 2 * R; k += T) {
        const int g = base + k;
        if (g >= 0 && g < m) { const int id = clusters[g]; s_lo[k] = t.lo[id]; s_hi[k] = t.hi[id]; }
    }
    __syncthreads();
    const int i = blockIdx.x * T + threadIdx.x;
    if (i >= m) return;
    const float4 al = s_lo[threadIdx.x + R], ah = s_hi[threadIdx.x + R];
    float best = FLT_MAX; int bj = -1;
    for (int k = 0; k <= 2 * R; ++k) {
        const int j = base + threadIdx.x + k;                     // i - R + k
        if (k == R || j < 0 || j >= m) continue;
        const float4 bl = s_lo[threadIdx.x + k], bh = s_hi[threadIdx.x + k];
        const float4 ul = {fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f};
        const float4 uh = {fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f};
        const float a = box_area(ul, uh);
        if (a < best) { best = a; bj = j; }
    }
    nearest[i] = bj;
}

// flags[i]: low 32 bits = 1 if cluster i survives (alone or as the merged pair's
// left member), high 32 bits = 1 if i leads a merge (allocates a node)
// (the grid covers n + 1 entries: everything from m on is zeroed, so that one scan of fixed length serves every iteration)
static __global__ void ploc_flag_kernel(const int* __restrict__ nearest, const BuildState* __restrict__ state, int n,
                                        uint64_t* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = state->m;
    if (m <= 1 || i > n) return;
    if (i >= m) { flags[i] = 0; return; }
    const int j = nearest[i];
    const bool mutual = j >= 0 && nearest[j] == i;
    const uint64_t keep = !(mutual && j < i);
    const uint64_t lead = mutual && i < j;
    flags[i] = keep | (lead << 32);
}

// scan[i] = exclusive prefix sums of flags; merged leaders create node n + first_node + (#leaders before i)
static __global__ void ploc_merge_kernel(const int* __restrict__ clusters, const int* __restrict__ nearest,
                                  const uint64_t* __restrict__ flags, const uint64_t* __restrict__ scan,
                                  const BuildState* __restrict__ state, int n, BinTree t, int* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = state->m, first_node = state->made;
    if (m <= 1 || i >= m) return;
    const uint64_t f = flags[i];
    if (!(f & 1u)) return;
    const int pos = int(scan[i] & 0xffffffffu);
    int id = clusters[i];
    if (f >> 32) {
        const int other = clusters[nearest[i]];
        const int k = first_node + int(scan[i] >> 32);            // internal index
        t.children[k] = make_int2(id, other);
        const float4 al = t.lo[id], ah = t.hi[id], bl = t.lo[other], bh = t.hi[other];
        t.lo[n + k] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z),
                                  __int_as_float(__float_as_int(al.w) + __float_as_int(bl.w)));
        t.hi[n + k] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
        id = n + k;
    }
    out[pos] = id;
}

// after the merge: m <- survivors, made += merged (scan[m] holds both totals); an iteration without a merge cannot
// happen (the globally closest pair is mutual) and is reported instead of looping for ever
static __global__ void ploc_advance_kernel(const uint64_t* __restrict__ scan, BuildState* __restrict__ state)
{
    const int m = state->m;
    if (m <= 1) return;
    const uint64_t tot = scan[m];
    const int kept = int(tot & 0xffffffffu), merged = int(tot >> 32);
    if (merged < 1 || kept != m - merged) { state->error = 1; state->m = 0; return; }
    state->m = kept;
    state->made += merged;
}

// ---- collapse to the 8-wide compressed BVH ------------------------------------
struct CollapseCounters { int nodes, tris, next; int pad; };

__device__ __forceinline__ int bin_count(const BinTree& t, int id) { return __float_as_int(t.lo[id].w); }

// Node layout (8 x 16 bytes):
//   w0 = origin.x, origin.y, origin.z (float bits), ex | ey << 8 | ez << 16 | imask << 24
//   w1 = child base, triangle base, meta[0..3], meta[4..7]
//   w2 = qlo.x[0..7]   w3 = qlo.y[0..7]   w4 = qlo.z[0..7]      (8 x 16 bits each, child s in half-word s)
//   w5 = qhi.x[0..7]   w6 = qhi.y[0..7]   w7 = qhi.z[0..7]
// child box plane = origin + q * 2^(e - 127 - 8), q in [0, 32767]; the exponent byte is stored with the + 8
// the traversal's mantissa trick needs (see node8_step).  imask bit s = slot s holds an internal child, stored at
// node index child base + popc(imask & ((1 << s) - 1)); meta[s] = 0 (empty), 1 << 5 | (24 + s) (internal) or
// unary(n_tris) << 5 | offset (leaf whose triangles sit at triangle base + offset, leaf order).
// Children are assigned to slots so that slot bit a says on which side of the node's centre (axis a) the child
// lies; a ray then visits the slots of a popped group in the order of (slot ^ its direction octant), near side
// first, without sorting anything.
// One thread per (binary node -> wide node index) task of this level.
static __global__ void collapse8_kernel(const int2* __restrict__ tasks, const BuildState* __restrict__ state, int n, BinTree t,
                                        const uint32_t* __restrict__ sorted_tri, uint4* __restrict__ nodes,
                                        int32_t* __restrict__ leaf_order, CollapseCounters* __restrict__ cnt,
                                        int2* __restrict__ next_tasks)
{
    const int ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= state->n_tasks) return;
    const int b = tasks[ti].x, self = tasks[ti].y;
    int cand[8]; int nc;
    if (b < n || bin_count(t, b) <= kLeafMax) { cand[0] = b; nc = 1; }        // tiny mesh: the root is a leaf
    else {
        const int2 ch = t.children[b - n];
        cand[0] = ch.x; cand[1] = ch.y; nc = 2;
        while (nc < 8) {                                  // open the largest child that is not a leaf yet
            int pick = -1; float pa = -1.f;
            for (int k = 0; k < nc; ++k) {
                const int id = cand[k];
                if (id < n || bin_count(t, id) <= kLeafMax) continue;
                const float a = box_area(t.lo[id], t.hi[id]);
                if (a > pa) { pa = a; pick = k; }
            }
            if (pick < 0) break;
            const int2 c2 = t.children[cand[pick] - n];
            cand[pick] = c2.x; cand[nc++] = c2.y;
        }
    }
    // quantisation frame of this node
    const float4 nlo = t.lo[b], nhi = t.hi[b];
    const float org[3] = {nlo.x, nlo.y, nlo.z}, top[3] = {nhi.x, nhi.y, nhi.z};
    int ebits[3]; double inv_scale[3];
    for (int a = 0; a < 3; ++a) {
        const float ext = top[a] - org[a];
        int e = ext > 0.f ? ilogbf(ext * (1.0f / 32767.0f)) + 1 : -126;     // 2^e > ext / 32767
        e = max(-126, min(110, e));
        ebits[a] = e + 127 + 8;                           // stored with the + 8 of node8_step's a' = 256 * 2^e / d
        inv_scale[a] = exp2(double(-e));
    }
    // slot assignment: greedy on cost(child, slot) = sum_a (slot bit a ? + : -) (child centre - node centre)_a
    float cv[8][3];
    for (int k = 0; k < nc; ++k) {
        const float4 l = t.lo[cand[k]], h = t.hi[cand[k]];
        cv[k][0] = (l.x + h.x) - (nlo.x + nhi.x); cv[k][1] = (l.y + h.y) - (nlo.y + nhi.y); cv[k][2] = (l.z + h.z) - (nlo.z + nhi.z);
    }
    int slot_of[8], child_in[8];
    for (int k = 0; k < 8; ++k) { slot_of[k] = -1; child_in[k] = -1; }
    for (int round = 0; round < nc; ++round) {
        float bestc = -FLT_MAX; int bk = -1, bs = -1;
        for (int k = 0; k < nc; ++k) {
            if (slot_of[k] >= 0) continue;
            for (int sl = 0; sl < 8; ++sl) {
                if (child_in[sl] >= 0) continue;
                const float c = ((sl & 1) ? cv[k][0] : -cv[k][0]) + ((sl & 2) ? cv[k][1] : -cv[k][1]) + ((sl & 4) ? cv[k][2] : -cv[k][2]);
                if (c > bestc) { bestc = c; bk = k; bs = sl; }
            }
        }
        slot_of[bk] = bs; child_in[bs] = bk;
    }
    // allocation: internal children contiguous, triangles of the leaf children contiguous
    int n_int = 0, n_tri = 0;
    for (int sl = 0; sl < 8; ++sl) {
        if (child_in[sl] < 0) continue;
        const int c = bin_count(t, cand[child_in[sl]]);
        if (c <= kLeafMax) n_tri += c; else ++n_int;
    }
    const int child_base = n_int ? atomicAdd(&cnt->nodes, n_int) : 0;
    const int task_base = n_int ? atomicAdd(&cnt->next, n_int) : 0;
    const int tri_base = n_tri ? atomicAdd(&cnt->tris, n_tri) : 0;
    uint32_t imask = 0, meta[2] = {0u, 0u}, q[6][4] = {};
    int ri = 0, off = 0;
    for (int sl = 0; sl < 8; ++sl) {
        if (child_in[sl] < 0) continue;
        const int id = cand[child_in[sl]];
        const float4 l = t.lo[id], h = t.hi[id];
        const float cl[3] = {l.x, l.y, l.z}, chh[3] = {h.x, h.y, h.z};
        for (int a = 0; a < 3; ++a) {
            const int lo_q = max(0, min(32767, int(floor((double(cl[a]) - double(org[a])) * inv_scale[a]))));
            const int hi_q = max(0, min(32767, int(ceil((double(chh[a]) - double(org[a])) * inv_scale[a]))));
            q[a][sl >> 1] |= uint32_t(lo_q) << (16 * (sl & 1));
            q[3 + a][sl >> 1] |= uint32_t(hi_q) << (16 * (sl & 1));
        }
        const int c = __float_as_int(l.w);
        uint32_t mb;
        if (c <= kLeafMax) {                              // leaf: its triangles at tri_base + off, in leaf order
            int st[kLeafMax], sp = 0, w = 0;
            st[sp++] = id;
            while (sp > 0) {
                const int x = st[--sp];
                if (x < n) leaf_order[tri_base + off + w++] = int32_t(sorted_tri[x]);
                else { const int2 c2 = t.children[x - n]; st[sp++] = c2.y; st[sp++] = c2.x; }
            }
            mb = (((1u << c) - 1u) << 5) | uint32_t(off);
            off += c;
        } else {
            imask |= 1u << sl;
            next_tasks[task_base + ri] = make_int2(id, child_base + ri);
            ++ri;
            mb = (1u << 5) | uint32_t(24 + sl);
        }
        meta[sl >> 2] |= mb << (8 * (sl & 3));
    }
    uint4* out = nodes + (size_t)self * kNodeStride;
    out[0] = make_uint4(__float_as_uint(org[0]), __float_as_uint(org[1]), __float_as_uint(org[2]),
                        uint32_t(ebits[0]) | uint32_t(ebits[1]) << 8 | uint32_t(ebits[2]) << 16 | imask << 24);
    out[1] = make_uint4(uint32_t(child_base), uint32_t(tri_base), meta[0], meta[1]);
#pragma unroll
    for (int k = 0; k < 6; ++k) out[2 + k] = make_uint4(q[k][0], q[k][1], q[k][2], q[k][3]);
}

// between two levels of the collapse: the tasks the level queued become the next level's
static __global__ void collapse_advance_kernel(CollapseCounters* __restrict__ cnt, BuildState* __restrict__ state)
{
    state->n_tasks = cnt->next;
    cnt->next = 0;
}

// leaf-ordered float triangles: (v0.xyz, e1.x) (e1.yz, e2.xy) (e2.z, max|e1|, max|e2|, original index)
static __global__ void leaf_triangles_kernel(const int32_t* __restrict__ leaf_order, const double* __restrict__ tri64, int n,
                                      float4* __restrict__ tri32)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int tri = leaf_order[s];
    const double* t = tri64 + (size_t)tri * kTri64Stride;
    float f[9];
    for (int k = 0; k < 9; ++k) f[k] = float(t[k]);
    // max-norms rounded up: they scale the float cull's error bound
    const float c1 = fmaxf(fabsf(f[3]), fmaxf(fabsf(f[4]), fabsf(f[5]))) * 1.0000002f;
    const float c2 = fmaxf(fabsf(f[6]), fmaxf(fabsf(f[7]), fabsf(f[8]))) * 1.0000002f;
    tri32[(size_t)s * kTri32Stride + 0] = make_float4(f[0], f[1], f[2], f[3]);
    tri32[(size_t)s * kTri32Stride + 1] = make_float4(f[4], f[5], f[6], f[7]);
    tri32[(size_t)s * kTri32Stride + 2] = make_float4(f[8], c1, c2, __int_as_float(tri));
    tri32[(size_t)s * kTri32Stride + 3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

#endif  // DRTB_BVH_BUILD_KERNELS

// ---------------------------------------------------------------------------
// ray-triangle and traversal
// ---------------------------------------------------------------------------
template <typename R> struct TriData { V3<R> v0, e1, e2; };

// geometry of triangle `tri` (ORIGINAL index) in the precision of the instantiation
template <typename R> __device__ __forceinline__ TriData<R> load_tri(const MeshView& m, int tri)
{
    const double2* p = reinterpret_cast<const double2*>(m.tri64 + (size_t)tri * kTri64Stride);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
    return {{R(a.x), R(a.y), R(b.x)}, {R(b.y), R(c.x), R(c.y)}, {R(d.x), R(d.y), R(e.x)}};
}

// 256-bit read-only load (LDG.E.ENL2.256 on sm_100): one L1 wavefront per lane where two
// LDG.128 cost two -- the traversal kernel is bound by L1 data-pipe wavefronts
// (profiles/r01_wf_traverse_f64_v3_smemstack_summary.txt: 73 % of peak).
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b)
{
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

struct TriF { float v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z, c1, c2; int id; };
__device__ __forceinline__ TriF load_trif(const MeshView& m, int slot)
{
    float4 a, b;
    ldg256(m.tri32 + (size_t)slot * kTri32Stride, a, b);
    const float4 c = __ldg(m.tri32 + (size_t)slot * kTri32Stride + 2);
    return {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, __float_as_int(c.w)};
}

// Moller-Trumbore with the acceptance rules of drtb.h.  `tmin`/`best` are the
// closest hit so far (best < 0: an analytic primitive or nothing, which wins ties).
template <typename R>
__device__ __forceinline__ void tri_test_exact(const TriData<R>& T, int tri, V3<R> o, V3<R> d, R& tmin, int& best)
{
    const V3<R> p = cross(d, T.e2);
    const R inv = Real<R>::rcp(dot(T.e1, p));          // det == 0 -> inf -> NaN below -> miss
    const V3<R> tv = {o.x - T.v0.x, o.y - T.v0.y, o.z - T.v0.z};
    const R u = dot(tv, p) * inv;
    const V3<R> q = cross(tv, T.e1);
    const R v = dot(d, q) * inv;
    const R t = dot(T.e2, q) * inv;
    const bool inside = u >= R(0) && u <= R(1) && v >= R(0) && u + v <= R(1) && t > R(0);
    if (inside && (t < tmin || (t == tmin && best >= 0 && tri < best))) { tmin = t; best = tri; }
}

// The ray as the float stages see it.
struct RayF {
    float ox, oy, oz, dx, dy, dz;
    float ix, iy, iz, oix, oiy, oiz;      // 1/d and o/d for the slab test  t = plane * (1/d) - o/d
    float om, dm;                         // max-norms of o and d (error bound of the float cull)
    uint32_t octinv4;                     // (7 - direction octant) in every byte: slot visiting order of the wide nodes
};

template <typename R>
__device__ __forceinline__ RayF make_rayf(V3<R> o, V3<R> d)
{
    RayF r;
    r.ox = float(o.x); r.oy = float(o.y); r.oz = float(o.z);
    r.dx = float(d.x); r.dy = float(d.y); r.dz = float(d.z);
    // a zero component would give inf * 0 = NaN in the slab test; 1e-30 keeps it finite and conservative
    const float sx = fabsf(r.dx) < 1e-30f ? copysignf(1e-30f, r.dx) : r.dx;
    const float sy = fabsf(r.dy) < 1e-30f ? copysignf(1e-30f, r.dy) : r.dy;
    const float sz = fabsf(r.dz) < 1e-30f ? copysignf(1e-30f, r.dz) : r.dz;
    r.ix = 1.0f / sx; r.iy = 1.0f / sy; r.iz = 1.0f / sz;
    r.oix = r.ox * r.ix; r.oiy = r.oy * r.iy; r.oiz = r.oz * r.iz;
    r.om = fmaxf(fabsf(r.ox), fmaxf(fabsf(r.oy), fabsf(r.oz)));
    r.dm = fmaxf(fabsf(r.dx), fmaxf(fabsf(r.dy), fabsf(r.dz)));
    // near slot = the child octant the ray enters first: bit a set where d_a < 0 (it comes from the high side)
    const uint32_t near_slot = (r.ix < 0.f ? 1u : 0u) | (r.iy < 0.f ? 2u : 0u) | (r.iz < 0.f ? 4u : 0u);
    r.octinv4 = (near_slot ^ 7u) * 0x01010101u;
    return r;
}

// Conservative float cull for the double instantiation: true = the exact test
// CANNOT accept this triangle.  Every Moller-Trumbore numerator is evaluated in
// float; its distance from the exact (double) value is bounded by
//   K (|o|+|tv|) |d| |e2|  (u),  K (|o|+|tv|) |d| |e1|  (v),
//   K (|o|+|tv|) |e1| |e2| (t),  K |d| |e1| |e2|        (det)      (max-norms)
// with K = 128 * 2^-24, which covers the rounding of o, d, v0, e1, e2 to float
// (cancellation in o - v0 included) and every float operation (worst case 72
// units).  A triangle is culled only if some acceptance condition fails by
// more than its bound; otherwise the exact double test decides.
__device__ __forceinline__ bool tri_cull_f(const TriF& T, const RayF& r, float tmaxf)
{
    const float px = r.dy * T.e2z - r.dz * T.e2y, py = r.dz * T.e2x - r.dx * T.e2z, pz = r.dx * T.e2y - r.dy * T.e2x;
    const float det = T.e1x * px + T.e1y * py + T.e1z * pz;
    const float tx = r.ox - T.v0x, ty = r.oy - T.v0y, tz = r.oz - T.v0z;
    const float up = tx * px + ty * py + tz * pz;
    const float qx = ty * T.e1z - tz * T.e1y, qy = tz * T.e1x - tx * T.e1z, qz = tx * T.e1y - ty * T.e1x;
    const float vp = r.dx * qx + r.dy * qy + r.dz * qz;
    const float tp = T.e2x * qx + T.e2y * qy + T.e2z * qz;
    const float K = 128.0f / 16777216.0f;
    const float a = K * (r.om + fmaxf(fabsf(tx), fmaxf(fabsf(ty), fabsf(tz))));
    const float ab = a * r.dm, c12 = T.c1 * T.c2;
    const float Eu = ab * T.c2, Ev = ab * T.c1, Et = a * c12, Ed = K * r.dm * c12;
    const float ad = fabsf(det);
    const uint32_t s = __float_as_uint(det) & 0x80000000u;
    const float us = __uint_as_float(__float_as_uint(up) ^ s), vs = __uint_as_float(__float_as_uint(vp) ^ s),
                ts = __uint_as_float(__float_as_uint(tp) ^ s);
    // bitwise, not short-circuit: the compiler turned the || chain into branches that ran at 5-8 of 32 lanes
    const bool out = (us < -Eu) | (vs < -Ev) | (us + vs > ad + (Eu + Ev + Ed)) | (ts < -Et) |
                     (ts > fmaf(tmaxf, ad + Ed, Et) * 1.000001f);
    return (ad > Ed) & out;
}

template <typename R> __device__ __forceinline__ float upper_float(R t);
template <> __device__ __forceinline__ float upper_float<double>(double t) { return __double2float_ru(t); }
template <> __device__ __forceinline__ float upper_float<float>(float t) { return t; }

// one triangle of a leaf (LEAF-order slot) against the running closest hit
template <typename R>
__device__ __forceinline__ void leaf_tri_test(const MeshView& m, int slot, const RayF& rf, V3<R> o, V3<R> d, R& tmin,
                                              int& best);
template <>
__device__ __forceinline__ void leaf_tri_test<float>(const MeshView& m, int slot, const RayF&, V3<float> o, V3<float> d,
                                                     float& tmin, int& best)
{
    const TriF T = load_trif(m, slot);
    const TriData<float> D = {{T.v0x, T.v0y, T.v0z}, {T.e1x, T.e1y, T.e1z}, {T.e2x, T.e2y, T.e2z}};
    tri_test_exact<float>(D, T.id, o, d, tmin, best);
}
template <>
__device__ __forceinline__ void leaf_tri_test<double>(const MeshView& m, int slot, const RayF& rf, V3<double> o,
                                                      V3<double> d, double& tmin, int& best)
{
    const TriF T = load_trif(m, slot);
    if (tri_cull_f(T, rf, upper_float<double>(tmin))) return;
    tri_test_exact<double>(load_tri<double>(m, T.id), T.id, o, d, tmin, best);
}

// test aid (DRTB_FLAG_NO_BVH) and stack-overflow fallback: the linear scan the BVH must agree with
template <typename R>
__device__ __forceinline__ void brute_closest(const MeshView& m, V3<R> o, V3<R> d, R& tmin, int& best, uint32_t& n_tests)
{
    for (int i = 0; i < m.n_tris; ++i) { ++n_tests; tri_test_exact<R>(load_tri<R>(m, i), i, o, d, tmin, best); }
}

// Traversal stacks.  An entry is a child GROUP of one visited node: (child base, hit mask << 24 | internal mask),
// the siblings still to be visited.  LocalStack lives in local memory: lanes sit at different depths, so one
// push touches up to 32 different cache lines.  SmemStack keeps the first kSmemStack entries in shared memory
// as [entry][thread] (every lane owns a bank column: conflict-free whatever its depth) and spills deeper
// entries to local memory.
struct LocalStack {
    uint2 e[kBvhStack];
    __device__ __forceinline__ void put(int i, uint2 v, bool pred) { if (pred) e[i] = v; }
    __device__ __forceinline__ uint2 get(int i) const { return e[i]; }
};
#ifndef DRTB_SMEM_STACK
#define DRTB_SMEM_STACK 12
#endif
constexpr int kSmemStack = DRTB_SMEM_STACK;
template <int THREADS>
struct SmemStack {
    uint2* col;                                  // &s_stack[0][threadIdx.x]; row kSmemStack is a write-only dummy
    uint2 spill[kBvhStack - kSmemStack];
    // predicated push without a branch on the common path: a lane that does not push writes the dummy row
    __device__ __forceinline__ void put(int i, uint2 v, bool pred)
    {
        if (pred && i >= kSmemStack) spill[i - kSmemStack] = v;
        else col[(pred ? i : kSmemStack) * THREADS] = v;
    }
    __device__ __forceinline__ uint2 get(int i) const { return i < kSmemStack ? col[i * THREADS] : spill[i - kSmemStack]; }
};

// Half-word H (0 / 1) of w dropped into a float's mantissa: bits = 0x43 << 24 | q << 8, i.e. 128 + q / 256 for the
// 15-bit q -- one PRMT, no conversion and no subtraction: the 128 is folded into the slab's offset (node8_step).
// `k43` is 0x43000000 held in a REGISTER (node8_step makes it opaque to the compiler): PRMT takes one immediate, and
// with the constant as the immediate the compiler re-materialised the selector before almost every one of the 48
// permutes of a node (profiles/r02_wf_traverse_f64_cw3_summary.txt: 79 instructions per node on this line).
template <int H> __device__ __forceinline__ float plane_to_float(uint32_t w, uint32_t k43)
{
    uint32_t f;
    if (H) asm("prmt.b32 %0, %1, %2, 0x7324;" : "=r"(f) : "r"(w), "r"(k43));
    else   asm("prmt.b32 %0, %1, %2, 0x7104;" : "=r"(f) : "r"(w), "r"(k43));
    return __uint_as_float(f);
}

// Slab test of child S of a wide node; ORs its hit bits into `hits` and keeps the two smallest entry
// distances of the INTERNAL children hit as integer keys (distance bits, low 3 bits = slot).  Branch-free.
//   wn*, wf* : the words holding child S's near / far plane per axis (swizzled by the ray's direction signs)
//   a*, b*   : t = f a + b for a plane f of that axis               (f = plane_to_float, see node8_step)
//   lb_lo/hi : 0x7f in byte s where child s is NOT an internal child (leaf_bytes); sl_lo/hi: 0x03020100 / 0x07060504
// Key of a child: one PRMT builds (leaf byte << 24 | slot), one LOP3 merges it with the distance bits, one SEL
// replaces it by the sentinel on a miss.  A leaf or empty slot so gets a key >= 0x7f000000 (1.7e38 as a float: beyond
// any entry distance), which never becomes the nearest INTERNAL child as long as one was hit, and node8_step's
// callers look at the keys only then.
constexpr float kSlabWiden = 1.0f + 1.0f / 524288.0f;
struct SlabCoef { float ax, ay, az, bx, by, bz; uint32_t k43, lb_lo, lb_hi, sl_lo, sl_hi; };
template <int S>
__device__ __forceinline__ void child_slab(uint32_t wnx, uint32_t wny, uint32_t wnz, uint32_t wfx, uint32_t wfy, uint32_t wfz,
                                           const SlabCoef& c, float tmax, uint32_t meta4, uint32_t& hits, int& m1, int& m2)
{
    constexpr int H = S & 1, J = S & 3;
    const float tnx = fmaf(plane_to_float<H>(wnx, c.k43), c.ax, c.bx), tny = fmaf(plane_to_float<H>(wny, c.k43), c.ay, c.by),
                tnz = fmaf(plane_to_float<H>(wnz, c.k43), c.az, c.bz);
    const float tfx = fmaf(plane_to_float<H>(wfx, c.k43), c.ax, c.bx), tfy = fmaf(plane_to_float<H>(wfy, c.k43), c.ay, c.by),
                tfz = fmaf(plane_to_float<H>(wfz, c.k43), c.az, c.bz);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
    // meta byte: internal child 1 << 5 | (24 + slot), leaf unary(n) << 5 | offset: its hit bits are (meta >> 5) << (meta & 31)
    const uint32_t mb = (meta4 >> (8 * J)) & 0xffu;
    const bool hit = tn <= tf * kSlabWiden;
    hits |= hit ? (mb >> 5) << (mb & 31u) : 0u;
    // result bytes 3..0 = (leaf byte J, 0, 0, slot): selector nibble 8 | x replicates the sign bit of byte x, and every
    // byte of the slot words is below 0x80, so "sign of byte 4" is the zero byte
    uint32_t x;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(x) : "r"(S < 4 ? c.lb_lo : c.lb_hi), "r"(S < 4 ? c.sl_lo : c.sl_hi),
        "n"((J << 12) | (0xc << 8) | (0xc << 4) | (4 + J)));
    const uint32_t kbits = (__float_as_uint(tn) & ~7u) | x;
    const int key = hit ? int(kbits) : 0x7fffffff;
    m2 = min(m2, max(m1, key));
    m1 = min(m1, key);
}

// 0x7f in byte s of (lo: s = 0..3, hi: s = 4..7) for every bit s of `mask`: the multiply spreads four bits to the low
// bits of four bytes without carries (bit i lands at 7 i + i), the second one widens 1 to 0x7f.
__device__ __forceinline__ void leaf_bytes(uint32_t mask, uint32_t& lo, uint32_t& hi)
{
    lo = (((mask & 0xfu) * 0x00204081u) & 0x01010101u) * 0x7fu;
    hi = ((((mask >> 4) & 0xfu) * 0x00204081u) & 0x01010101u) * 0x7fu;
}

// One wide-node step of a lane: fetch node `node`, slab-test its 8 children.  Returns the child group
// ng = (child base, internal hits BY SLOT << 24 | imask), the triangle group tg = (triangle base, hit bits of the
// triangles of the leaf children) and the keys (m1, m2) of the nearest and second-nearest internal child hit
// (meaningful when at least one / two internal children were hit; see child_slab).
//
// Plane a of child s is origin_a + q 2^E; with f = 128 + q / 256 (plane_to_float) the slab distance is
//   t = (origin_a + q 2^E - o_a) / d_a = f A + B,   A = 256 2^E / d_a,   B = (origin_a - o_a) / d_a - 128 A.
// The float evaluation of t on axis a is off by at most 2^-21 (|o_a| + mesh scale) / |d_a|.  For a ray that starts
// within a few mesh scales of the mesh that is inside the padding the leaf boxes get at build time (4e-6 mesh scales
// on every side, bin_leaves_kernel; every inner box contains its leaves' padded boxes).  For a ray that starts far
// away the error is relative to t instead (|o_a| ~ t |d_a|), and the far bound is widened by kSlabWiden = 1 + 2^-19
// in the one compare -- one multiply per child where scaling the twelve slab coefficients by (1 -/+ 4e-7) cost twelve
// per node and six registers.
__device__ __forceinline__ void node8_step(const MeshView& m, const RayF& r, float tmax, uint32_t node, uint2& ng, uint2& tg,
                                           int& m1, int& m2)
{
    const float4* nd = m.nodes + (size_t)node * kNodeStride;
    float4 f0, f1, f2, f3, f4, f5, f6, f7;
    ldg256(nd, f0, f1); ldg256(nd + 2, f2, f3); ldg256(nd + 4, f4, f5); ldg256(nd + 6, f6, f7);
    const uint32_t ew = __float_as_uint(f0.w), imask = ew >> 24;
    const float ax = __uint_as_float((ew & 0xffu) << 23) * r.ix, ay = __uint_as_float(((ew >> 8) & 0xffu) << 23) * r.iy,
                az = __uint_as_float(((ew >> 16) & 0xffu) << 23) * r.iz;
    const float bx = fmaf(-128.0f, ax, (f0.x - r.ox) * r.ix), by = fmaf(-128.0f, ay, (f0.y - r.oy) * r.iy),
                bz = fmaf(-128.0f, az, (f0.z - r.oz) * r.iz);
    // constants as run-time values (the triangle base is < 2^28, which the compiler cannot know), so that they are
    // not folded back into the permutes as their one immediate
    const uint32_t opaque = __float_as_uint(f1.y) >> 31;
    SlabCoef c = {ax, ay, az, bx, by, bz, 0x43000000u | opaque, 0u, 0u, 0x03020100u | opaque, 0x07060504u | opaque};
    leaf_bytes(~imask, c.lb_lo, c.lb_hi);
    const bool sx = r.ix < 0.f, sy = r.iy < 0.f, sz = r.iz < 0.f;
    // near / far plane words per axis: lo planes in w2..w4, hi planes in w5..w7
    const float4 nX = sx ? f5 : f2, fX = sx ? f2 : f5, nY = sy ? f6 : f3, fY = sy ? f3 : f6, nZ = sz ? f7 : f4, fZ = sz ? f4 : f7;
    uint32_t hits = 0;
    m1 = 0x7fffffff; m2 = 0x7fffffff;
    const uint32_t meta_lo = __float_as_uint(f1.z), meta_hi = __float_as_uint(f1.w);
#define DRTB_W(V, C) __float_as_uint(V.C)
    child_slab<0>(DRTB_W(nX, x), DRTB_W(nY, x), DRTB_W(nZ, x), DRTB_W(fX, x), DRTB_W(fY, x), DRTB_W(fZ, x), c, tmax, meta_lo, hits, m1, m2);
    child_slab<1>(DRTB_W(nX, x), DRTB_W(nY, x), DRTB_W(nZ, x), DRTB_W(fX, x), DRTB_W(fY, x), DRTB_W(fZ, x), c, tmax, meta_lo, hits, m1, m2);
    child_slab<2>(DRTB_W(nX, y), DRTB_W(nY, y), DRTB_W(nZ, y), DRTB_W(fX, y), DRTB_W(fY, y), DRTB_W(fZ, y), c, tmax, meta_lo, hits, m1, m2);
    child_slab<3>(DRTB_W(nX, y), DRTB_W(nY, y), DRTB_W(nZ, y), DRTB_W(fX, y), DRTB_W(fY, y), DRTB_W(fZ, y), c, tmax, meta_lo, hits, m1, m2);
    child_slab<4>(DRTB_W(nX, z), DRTB_W(nY, z), DRTB_W(nZ, z), DRTB_W(fX, z), DRTB_W(fY, z), DRTB_W(fZ, z), c, tmax, meta_hi, hits, m1, m2);
    child_slab<5>(DRTB_W(nX, z), DRTB_W(nY, z), DRTB_W(nZ, z), DRTB_W(fX, z), DRTB_W(fY, z), DRTB_W(fZ, z), c, tmax, meta_hi, hits, m1, m2);
    child_slab<6>(DRTB_W(nX, w), DRTB_W(nY, w), DRTB_W(nZ, w), DRTB_W(fX, w), DRTB_W(fY, w), DRTB_W(fZ, w), c, tmax, meta_hi, hits, m1, m2);
    child_slab<7>(DRTB_W(nX, w), DRTB_W(nY, w), DRTB_W(nZ, w), DRTB_W(fX, w), DRTB_W(fY, w), DRTB_W(fZ, w), c, tmax, meta_hi, hits, m1, m2);
#undef DRTB_W
    ng = make_uint2(__float_as_uint(f1.x), (hits & 0xff000000u) | imask);
    tg = make_uint2(__float_as_uint(f1.y), hits & 0x00ffffffu);
}

// Child groups on the stack: x = child base, y = hits << 24 | bound << 8 | imask, hit bit 24 + s = the internal child
// in slot s, and `bound` = the upper 16 bits of a float that is <= the entry distance of every child left in the
// group (a conservative cull at pop time).
constexpr uint32_t kHitBits = 0xff000000u;
// Index of the child node in slot `slot` of group g.
__device__ __forceinline__ uint32_t child_index(uint2 g, uint32_t slot)
{
    return g.x + __popc(g.y & 0xffu & ((1u << slot) - 1u));       // rank among the internal children
}
// After node8_step: take the NEAREST internal child hit (key m1) out of ng; what is left goes to `rest` with the
// second-nearest distance as its bound (rest.y & kHitBits == 0: nothing left).  Returns the child's node index.
__device__ __forceinline__ uint32_t take_nearest(uint2 ng, int m1, int m2, uint2& rest)
{
    const uint32_t slot = uint32_t(m1) & 7u;
    rest = make_uint2(ng.x, (ng.y & ~(0x01000000u << slot)) | ((uint32_t(m2) >> 8) & 0x00ffff00u));
    return child_index(ng, slot);
}
// A popped group: false if its bound is beyond tmax (every child left is culled); else takes the child on the
// ray's near side -- the slot s with the smallest s ^ near_slot, where slot bit a says on which side of the node's
// centre the child lies -- and returns its node index, leaving the others in g.
__device__ __forceinline__ bool take_from_popped(uint2& g, float tmax, uint32_t octinv4, uint32_t& node)
{
    if (__uint_as_float((g.y << 8) & 0xffff0000u) > tmax) return false;
    const uint32_t oi = octinv4 & 7u, h = g.y >> 24;
    // permute the hit bits by s -> s ^ octinv (three conditional swaps); the highest bit is then the nearest slot
    uint32_t v = h;
    v = (oi & 1u) ? ((v & 0x55u) << 1) | ((v >> 1) & 0x55u) : v;
    v = (oi & 2u) ? ((v & 0x33u) << 2) | ((v >> 2) & 0x33u) : v;
    v = (oi & 4u) ? ((v & 0x0fu) << 4) | ((v >> 4) & 0x0fu) : v;
    const uint32_t slot = (31u - __clz(v)) ^ oi;
    g.y &= ~(0x01000000u << slot);
    node = child_index(g, slot);
    return true;
}

// Closest triangle along (o, d) that beats `tmin`: the per-lane form of the traversal (explicit rays,
// drtb_trace_rays); the wavefront's wf_traverse runs the same steps warp-synchronously.
template <typename R>
__device__ __forceinline__ void bvh_closest(const MeshView& m, V3<R> o, V3<R> d, R& tmin, int& best,
                                            uint32_t& n_nodes, uint32_t& n_tests)
{
    const RayF r = make_rayf(o, d);
    float tmax = upper_float<R>(tmin);
    LocalStack stack;
    int sp = 0;
    bool overflow = false;
    uint32_t node = 0;                                           // the root
    for (;;) {
        uint2 ng, tg, rest;
        int m1, m2;
        ++n_nodes;
        node8_step(m, r, tmax, node, ng, tg, m1, m2);
        bool have = false;
        if (ng.y & kHitBits) {
            node = take_nearest(ng, m1, m2, rest);
            have = true;
            if (rest.y & kHitBits) { if (sp < kBvhStack) stack.e[sp++] = rest; else overflow = true; }
        }
        while (tg.y) {
            const int bit = __ffs(tg.y) - 1;
            tg.y &= tg.y - 1u;
            ++n_tests;
            leaf_tri_test<R>(m, int(tg.x) + bit, r, o, d, tmin, best);
        }
        tmax = upper_float<R>(tmin);
        while (!have && sp > 0) {
            uint2 g = stack.e[--sp];
            if (take_from_popped(g, tmax, r.octinv4, node)) {
                have = true;
                if (g.y & kHitBits) stack.e[sp++] = g;
            }
        }
        if (!have) break;
    }
    if (overflow) brute_closest<R>(m, o, d, tmin, best, n_tests);   // never seen on a PLOC tree; correctness first
}

} // namespace drtb
