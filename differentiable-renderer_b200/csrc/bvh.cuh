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
//   3. collapse to a 4-wide BVH, top-down and level-synchronous: a wide node
//      adopts the grandchildren with the largest surface area first; subtrees
//      of <= kLeafMax triangles become leaves and their triangles are stored
//      contiguously in leaf order.
// Wide node = 128 B (one cache line): the 4 child boxes as SoA float4 rows
// (lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4]) + 4 child links.  Boxes
// are float, rounded outward and padded, and the slab test is conservative, so
// float culling can never reject a triangle the exact test would accept.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "real.cuh"

namespace drtb {

constexpr int kBvhStack   = 64;           // traversal stack entries per lane (overflow falls back to a linear scan)
#ifndef DRTB_LEAF_MAX
#define DRTB_LEAF_MAX 4
#endif
constexpr int kLeafMax    = DRTB_LEAF_MAX; // triangles per leaf, <= 4 (2 bits of the leaf link)
constexpr int kPlocRadius = 16;           // PLOC neighbour search radius along the Morton order
constexpr int kTri64Stride = 10;          // doubles per triangle: v0, e1, e2, pad (16-byte aligned rows)
constexpr int kTri32Stride = 4;           // float4 per triangle (64 B: two 256-bit loads)
constexpr int kNodeStride  = 8;           // float4 per wide node
constexpr int kEmptyLink   = 0x7fffffff;

struct MeshView {
    const float4*  nodes;                 // kNodeStride float4 per wide node; node 0 is the root
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
    float ext = 0.f;
    for (int a = 0; a < 3; ++a) ext = fmaxf(ext, ordered_to_float(bounds[3 + a]) - ordered_to_float(bounds[a]));
    const float pad = 4e-6f * ext + FLT_MIN;          // covers float rounding of the ray and of the slab test
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
static __global__ void __launch_bounds__(256)
ploc_nearest_kernel(const int* __restrict__ clusters, int m, BinTree t, int* __restrict__ nearest)
{
    constexpr int R = kPlocRadius, T = 256;
    __shared__ float4 s_lo[T + 2 * R], s_hi[T + 2 * R];
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
static __global__ void ploc_flag_kernel(const int* __restrict__ nearest, int m, uint64_t* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nearest[i];
    const bool mutual = j >= 0 && nearest[j] == i;
    const uint64_t keep = !(mutual && j < i);
    const uint64_t lead = mutual && i < j;
    flags[i] = keep | (lead << 32);
}

// scan[i] = exclusive prefix sums of flags; merged leaders create node n + first_node + (#leaders before i)
static __global__ void ploc_merge_kernel(const int* __restrict__ clusters, const int* __restrict__ nearest,
                                  const uint64_t* __restrict__ flags, const uint64_t* __restrict__ scan, int m, int n,
                                  int first_node, BinTree t, int* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
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

// ---- collapse to the 4-wide BVH ----------------------------------------------
struct CollapseCounters { int nodes, tris, next; int pad; };

__device__ __forceinline__ int bin_count(const BinTree& t, int id) { return __float_as_int(t.lo[id].w); }

// One thread per (binary node -> wide node slot) task of this level.
static __global__ void collapse_kernel(const int2* __restrict__ tasks, int n_tasks, int n, BinTree t,
                                const uint32_t* __restrict__ sorted_tri, float4* __restrict__ nodes,
                                int32_t* __restrict__ leaf_order, CollapseCounters* __restrict__ cnt,
                                int2* __restrict__ next_tasks)
{
    const int ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= n_tasks) return;
    const int b = tasks[ti].x, slot = tasks[ti].y;
    int cand[4]; int nc;
    if (b < n || bin_count(t, b) <= kLeafMax) { cand[0] = b; nc = 1; }        // tiny mesh: the root is a leaf
    else {
        const int2 ch = t.children[b - n];
        cand[0] = ch.x; cand[1] = ch.y; nc = 2;
        while (nc < 4) {                                  // open the largest child that is not a leaf yet
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
    float lo[3][4], hi[3][4]; int link[4];
    for (int k = 0; k < 4; ++k) {
        if (k >= nc) {
            // NaN planes: every slab compare fails, the empty slot can never be entered
            for (int a = 0; a < 3; ++a) { lo[a][k] = __int_as_float(0x7fc00000); hi[a][k] = __int_as_float(0x7fc00000); }
            link[k] = kEmptyLink;
            continue;
        }
        const int id = cand[k];
        const float4 l = t.lo[id], h = t.hi[id];
        lo[0][k] = l.x; lo[1][k] = l.y; lo[2][k] = l.z; hi[0][k] = h.x; hi[1][k] = h.y; hi[2][k] = h.z;
        const int c = __float_as_int(l.w);
        if (c <= kLeafMax) {                              // leaf: triangles stored contiguously in leaf order
            const int first = atomicAdd(&cnt->tris, c);
            int st[kLeafMax], sp = 0, w = 0;
            st[sp++] = id;
            while (sp > 0) {
                const int x = st[--sp];
                if (x < n) leaf_order[first + w++] = int32_t(sorted_tri[x]);
                else { const int2 c2 = t.children[x - n]; st[sp++] = c2.y; st[sp++] = c2.x; }
            }
            link[k] = ~((first << 2) | (c - 1));
        } else {
            const int s = atomicAdd(&cnt->nodes, 1);
            next_tasks[atomicAdd(&cnt->next, 1)] = make_int2(id, s);
            link[k] = s;
        }
    }
    float4* out = nodes + (size_t)slot * kNodeStride;
    for (int a = 0; a < 3; ++a) {
        out[a] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
        out[3 + a] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
    }
    out[6] = make_float4(__int_as_float(link[0]), __int_as_float(link[1]), __int_as_float(link[2]), __int_as_float(link[3]));
    out[7] = make_float4(0.f, 0.f, 0.f, 0.f);
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
    const bool out = us < -Eu || vs < -Ev || us + vs > ad + (Eu + Ev + Ed) || ts < -Et ||
                     ts > fmaf(tmaxf, ad + Ed, Et) * 1.000001f;
    return ad > Ed && out;
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

// sort key of a slab hit: entry distance (non-negative float, so its bits order
// like an int) with the child slot in the two low mantissa bits; clearing them
// only lowers the distance, which keeps the pop-time cull conservative
__device__ __forceinline__ int hit_key(float tn, float tf, float tmax, int j)
{
    return (tn <= tf && tn <= tmax) ? ((__float_as_int(tn) & ~3) | j) : 0x7fffffff;
}
__device__ __forceinline__ void cswap(int& a, int& b) { const int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
__device__ __forceinline__ int pick_link(float4 lk, int j)
{
    const float a = (j & 1) ? lk.y : lk.x, b = (j & 1) ? lk.w : lk.z;
    return __float_as_int((j & 2) ? b : a);
}

// Traversal stacks.  Entries are (sort key, link).  LocalStack lives in local memory:
// lanes sit at different depths, so one push touches up to 32 different cache lines.
// SmemStack keeps the first kSmemStack entries in shared memory as [entry][thread]
// (every lane owns a bank column: conflict-free whatever its depth) and spills deeper
// entries to local memory.
struct LocalStack {
    int2 e[kBvhStack];
    __device__ __forceinline__ void put(int i, int2 v, bool pred) { if (pred) e[i] = v; }
    __device__ __forceinline__ int2 get(int i) const { return e[i]; }
};
#ifndef DRTB_SMEM_STACK
#define DRTB_SMEM_STACK 16
#endif
constexpr int kSmemStack = DRTB_SMEM_STACK;
template <int THREADS>
struct SmemStack {
    int2* col;                                   // &s_stack[0][threadIdx.x]; row kSmemStack is a write-only dummy
    int2 spill[kBvhStack - kSmemStack];
    // predicated push without a branch on the common path: a lane that does not push writes the dummy row
    __device__ __forceinline__ void put(int i, int2 v, bool pred)
    {
        if (pred && i >= kSmemStack) spill[i - kSmemStack] = v;
        else col[(pred ? i : kSmemStack) * THREADS] = v;
    }
    __device__ __forceinline__ int2 get(int i) const { return i < kSmemStack ? col[i * THREADS] : spill[i - kSmemStack]; }
};

// One wide-node step of a lane: slab-test the 4 children, sort the hits by entry
// distance, push the far ones (far first, so the nearest is popped first), descend
// into the nearest.  Returns true when nothing was hit (the caller pops).
template <typename Stack>
__device__ __forceinline__ bool bvh_node_step(const MeshView& m, const RayF& r, float tmax, int& cur, Stack& stack, int& sp,
                                              bool& overflow)
{
    const float4* nd = m.nodes + (size_t)cur * kNodeStride;
    float4 lx, ly, lz, hx, hy, hz;
    ldg256(nd, lx, ly); ldg256(nd + 2, lz, hx); ldg256(nd + 4, hy, hz);
    const float4 lk = __ldg(nd + 6);
    int key[4];
#define DRTB_SLAB(J, C)                                                                           \
    {                                                                                             \
        const float ax = fmaf(lx.C, r.ix, -r.oix), bx = fmaf(hx.C, r.ix, -r.oix);                 \
        const float ay = fmaf(ly.C, r.iy, -r.oiy), by = fmaf(hy.C, r.iy, -r.oiy);                 \
        const float az = fmaf(lz.C, r.iz, -r.oiz), bz = fmaf(hz.C, r.iz, -r.oiz);                 \
        float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));                     \
        float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));                     \
        tn = fmaxf(tn - fabsf(tn) * 4e-7f, 0.0f);           /* conservative: widen by a few ulp */ \
        tf += fabsf(tf) * 4e-7f;                                                                  \
        key[J] = hit_key(tn, tf, tmax, J);                                                        \
    }
    DRTB_SLAB(0, x) DRTB_SLAB(1, y) DRTB_SLAB(2, z) DRTB_SLAB(3, w)
#undef DRTB_SLAB
    cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
    // the stores are predicated, not branched
#pragma unroll
    for (int k = 3; k >= 1; --k) {
        const bool hit = key[k] != 0x7fffffff;
        stack.put(sp, make_int2(key[k], pick_link(lk, key[k] & 3)), hit && sp < kBvhStack);
        overflow |= hit && sp >= kBvhStack;
        sp += (hit && sp < kBvhStack) ? 1 : 0;
    }
    cur = pick_link(lk, key[0] & 3);
    return key[0] == 0x7fffffff;
}

// Closest triangle along (o, d) that beats `tmin`; ordered traversal, nearest child first.
template <typename R>
__device__ __forceinline__ void bvh_closest(const MeshView& m, V3<R> o, V3<R> d, R& tmin, int& best,
                                            uint32_t& n_nodes, uint32_t& n_tests)
{
    const RayF r = make_rayf(o, d);
    float tmax = upper_float<R>(tmin);
    LocalStack stack;                                    // (key, link)
    int sp = 0, cur = 0;
    bool overflow = false, alive = true;
    // Warp-synchronous stepping with a majority vote.  Left to the compiler's own
    // reconvergence this loop ran with 4.3 of 32 lanes active; converged but executing
    // the node code and the leaf code in every iteration, the leaf code still ran at
    // 2.9 of 32 (profiles/r01_mesh_f64_v2_bvh4_summary.txt, ..._v3_sync_summary.txt).
    // So each iteration issues ONE kind of step -- the one more lanes are waiting for --
    // and the minority keeps its state: lanes that reached a leaf wait until the leaf
    // lanes outnumber the lanes still descending, then all leaves are tested together.
    const unsigned live = __activemask();
#ifndef DRTB_LEAF_WEIGHT
#define DRTB_LEAF_WEIGHT 1
#endif
    for (;;) {
        const bool at_node = alive && cur >= 0, at_leaf = alive && cur < 0;
        const unsigned node_m = __ballot_sync(live, at_node), leaf_m = __ballot_sync(live, at_leaf);
        if ((node_m | leaf_m) == 0u) break;
        const bool leaf_step = node_m == 0u || DRTB_LEAF_WEIGHT * __popc(leaf_m) >= __popc(node_m);
        bool pop = false;
        if (!leaf_step) {
            if (at_node) {                               // wide node
                ++n_nodes;
                pop = bvh_node_step(m, r, tmax, cur, stack, sp, overflow);
            }
        } else if (at_leaf) {                            // leaf: ~((first << 2) | (count - 1))
            const int code = ~cur, first = code >> 2, count = (code & 3) + 1;
            for (int k = 0; k < count; ++k) { ++n_tests; leaf_tri_test<R>(m, first + k, r, o, d, tmin, best); }
            tmax = upper_float<R>(tmin);
            pop = true;
        }
        if (pop) {
            alive = false;
            while (sp > 0) {
                const int2 e = stack.get(--sp);
                if (__int_as_float(e.x & ~3) <= tmax) { cur = e.y; alive = true; break; }
            }
        }
    }
    if (overflow) brute_closest<R>(m, o, d, tmin, best, n_tests);   // never seen on a PLOC tree; correctness first
}

} // namespace drtb
