// bvh.cuh — triangle meshes: GPU-built LBVH and its traversal.
//
// NEW functionality (the reference has planes and spheres only, scanned
// linearly: pathtracer.hpp:72-89).  Semantics are fixed in include/drtb.h so
// that a BVH traversal returns exactly what the reference's linear scan would
// return if it had a Triangle shape: closest t > 0, exact ties to the lower
// scene index.
//
// Build (all on the device, Karras 2012): per-triangle bounds + centroid ->
// 63-bit Morton code -> cub radix sort -> binary radix tree over the sorted
// codes -> bottom-up refit with one atomic counter per internal node.
// Node = 64 B: the AABBs of BOTH children + two child links, so one 64-byte
// read (4 x float4 through the read-only path) decides both subtrees.
// Boxes are float, rounded outward and padded, and the slab test is
// conservative, so float culling can never reject a triangle the double
// intersection test would accept.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "real.cuh"

namespace drtb {

constexpr int kBvhStack = 96;            // >= 63 Morton bits + 32 index bits of LBVH depth
constexpr int kTri64Stride = 10;          // doubles per triangle: v0, e1, e2, pad (16-byte aligned rows)
constexpr int kTri32Stride = 3;           // float4 per triangle

struct MeshView {
    const float4*  nodes;                 // 4 float4 per node (n_tris - 1 nodes)
    const double*  tri64;                 // v0.xyz e1.xyz e2.xyz pad
    const float4*  tri32;                 // (v0.xyz,e1.x) (e1.yz,e2.xy) (e2.z,0,0,0)
    const int32_t* color;                 // per triangle: param index of the albedo, -1 = null BxDF
    const int32_t* emis;                  // per triangle: param index of the emission, -1 = none
    int32_t n_tris;
    int32_t n_prims;                      // scene index of triangle 0
};

// ---------------------------------------------------------------------------
// build kernels
// ---------------------------------------------------------------------------

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

// per triangle: edge form in double and float, outward-rounded float AABB, centroid; scene bounds
__global__ void mesh_prepare_kernel(const double* __restrict__ vertices, const int32_t* __restrict__ indices, int n,
                                    double* __restrict__ tri64, float4* __restrict__ tri32,
                                    float* __restrict__ leaf_lo, float* __restrict__ leaf_hi,
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
            leaf_lo[3ll * i + a] = lo[a]; leaf_hi[3ll * i + a] = hi[a];
        }
        t[9] = 0.0;
        tri32[(size_t)i * 3 + 0] = make_float4(float(t[0]), float(t[1]), float(t[2]), float(t[3]));
        tri32[(size_t)i * 3 + 1] = make_float4(float(t[4]), float(t[5]), float(t[6]), float(t[7]));
        tri32[(size_t)i * 3 + 2] = make_float4(float(t[8]), 0.f, 0.f, 0.f);
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

__global__ void mesh_morton_kernel(const float* __restrict__ leaf_lo, const float* __restrict__ leaf_hi,
                                   const uint32_t* __restrict__ bounds, int n, uint64_t* __restrict__ keys,
                                   uint32_t* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t code = 0;
    for (int a = 0; a < 3; ++a) {
        const float lo = ordered_to_float(bounds[a]), hi = ordered_to_float(bounds[3 + a]);
        const float c = 0.5f * (leaf_lo[3ll * i + a] + leaf_hi[3ll * i + a]);
        const float ext = hi - lo;
        float u = ext > 0.f ? (c - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        const uint32_t q = min(uint32_t(u * 2097152.0f), 2097151u);
        code |= spread21(q) << (2 - a);
    }
    keys[i] = code;
    vals[i] = uint32_t(i);
}

// longest common prefix of sorted keys i and j; equal keys fall back to the index (Karras 2012, §4)
__device__ __forceinline__ int lbvh_delta(const uint64_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(uint32_t(i) ^ uint32_t(j)) : __clzll((long long)(a ^ b));
}

// child link encoding while building: >= 0 internal node, < 0 leaf at SORTED position ~c
__global__ void lbvh_hierarchy_kernel(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children,
                                      int* __restrict__ parent_of_node, int* __restrict__ parent_of_leaf)
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
    const int left = lo == gamma ? ~gamma : gamma;
    const int right = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    if (left < 0) parent_of_leaf[~left] = i; else parent_of_node[left] = i;
    if (right < 0) parent_of_leaf[~right] = i; else parent_of_node[right] = i;
    if (i == 0) parent_of_node[0] = -1;
}

struct Box { float lo[3], hi[3]; };

__device__ __forceinline__ Box node_union(const volatile float* nd)   // union of a finished node's two child boxes
{
    Box b;
    for (int a = 0; a < 3; ++a) {
        b.lo[a] = fminf(nd[a], nd[6 + a]);
        b.hi[a] = fmaxf(nd[3 + a], nd[9 + a]);
    }
    return b;
}

// one thread per leaf climbs; the second arrival at a node owns it
__global__ void lbvh_refit_kernel(const uint32_t* __restrict__ sorted_tri, const float* __restrict__ leaf_lo,
                                  const float* __restrict__ leaf_hi, const int2* __restrict__ children,
                                  const int* __restrict__ parent_of_node, const int* __restrict__ parent_of_leaf,
                                  const uint32_t* __restrict__ bounds, int n, int* __restrict__ arrivals,
                                  float* __restrict__ nodes /* 16 floats per node */)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float ext = 0.f;
    for (int a = 0; a < 3; ++a) ext = fmaxf(ext, ordered_to_float(bounds[3 + a]) - ordered_to_float(bounds[a]));
    const float pad = 4e-6f * ext + FLT_MIN;          // covers float rounding of the ray and of the slab test
    int node = parent_of_leaf[p];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(arrivals + node, 1) == 0) return;
        const int2 ch = children[node];
        float* out = nodes + (size_t)node * 16;
        const int link[2] = {ch.x, ch.y};
        int enc[2];
        for (int k = 0; k < 2; ++k) {
            Box b;
            if (link[k] < 0) {
                const uint32_t tri = sorted_tri[~link[k]];
                for (int a = 0; a < 3; ++a) { b.lo[a] = leaf_lo[3ll * tri + a] - pad; b.hi[a] = leaf_hi[3ll * tri + a] + pad; }
                enc[k] = ~int(tri);                   // leaves point at the ORIGINAL triangle index
            } else {
                b = node_union(nodes + (size_t)link[k] * 16);
                enc[k] = link[k];
            }
            for (int a = 0; a < 3; ++a) { out[6 * k + a] = b.lo[a]; out[6 * k + 3 + a] = b.hi[a]; }
        }
        out[12] = __int_as_float(enc[0]); out[13] = __int_as_float(enc[1]); out[14] = 0.f; out[15] = 0.f;
        node = parent_of_node[node];
    }
}

// ---------------------------------------------------------------------------
// ray-triangle and traversal
// ---------------------------------------------------------------------------
template <typename R> struct TriData { V3<R> v0, e1, e2; };

template <typename R> __device__ __forceinline__ TriData<R> load_tri(const MeshView& m, int tri);
template <> __device__ __forceinline__ TriData<double> load_tri<double>(const MeshView& m, int tri)
{
    const double2* p = reinterpret_cast<const double2*>(m.tri64 + (size_t)tri * kTri64Stride);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
    return {{a.x, a.y, b.x}, {b.y, c.x, c.y}, {d.x, d.y, e.x}};
}
template <> __device__ __forceinline__ TriData<float> load_tri<float>(const MeshView& m, int tri)
{
    const float4 a = __ldg(m.tri32 + (size_t)tri * 3), b = __ldg(m.tri32 + (size_t)tri * 3 + 1),
                 c = __ldg(m.tri32 + (size_t)tri * 3 + 2);
    return {{a.x, a.y, a.z}, {a.w, b.x, b.y}, {b.z, b.w, c.x}};
}

// Moller-Trumbore with the acceptance rules of drtb.h.  `tmin`/`best` are the
// closest hit so far (best < 0: an analytic primitive or nothing, which wins ties).
template <typename R>
__device__ __forceinline__ void tri_test(const MeshView& m, int tri, V3<R> o, V3<R> d, R& tmin, int& best)
{
    const TriData<R> T = load_tri<R>(m, tri);
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

template <typename R> __device__ __forceinline__ float upper_float(R t);
template <> __device__ __forceinline__ float upper_float<double>(double t) { return __double2float_ru(t); }
template <> __device__ __forceinline__ float upper_float<float>(float t) { return t; }

// Closest triangle along (o, d) that beats `tmin`; ordered traversal, near child first.
template <typename R>
__device__ __forceinline__ void bvh_closest(const MeshView& m, V3<R> o, V3<R> d, R& tmin, int& best,
                                            uint32_t& n_nodes, uint32_t& n_tests)
{
    if (m.n_tris == 1) { ++n_tests; tri_test(m, 0, o, d, tmin, best); return; }
    const float ox = float(o.x), oy = float(o.y), oz = float(o.z);
    const float ix = 1.0f / float(d.x), iy = 1.0f / float(d.y), iz = 1.0f / float(d.z);
    float tmax = upper_float<R>(tmin);
    int stack[kBvhStack];
    float stack_t[kBvhStack];
    int sp = 0, cur = 0;
    for (;;) {
        ++n_nodes;
        const float4 n0 = __ldg(m.nodes + 4ll * cur), n1 = __ldg(m.nodes + 4ll * cur + 1),
                     n2 = __ldg(m.nodes + 4ll * cur + 2), n3 = __ldg(m.nodes + 4ll * cur + 3);
        // child 0: lo = n0.xyz, hi = (n0.w, n1.x, n1.y); child 1: lo = (n1.z, n1.w, n2.x), hi = n2.yzw
        float a, b;
        a = (n0.x - ox) * ix; b = (n0.w - ox) * ix; float tn0 = fminf(a, b), tf0 = fmaxf(a, b);
        a = (n0.y - oy) * iy; b = (n1.x - oy) * iy; tn0 = fmaxf(tn0, fminf(a, b)); tf0 = fminf(tf0, fmaxf(a, b));
        a = (n0.z - oz) * iz; b = (n1.y - oz) * iz; tn0 = fmaxf(tn0, fminf(a, b)); tf0 = fminf(tf0, fmaxf(a, b));
        a = (n1.z - ox) * ix; b = (n2.y - ox) * ix; float tn1 = fminf(a, b), tf1 = fmaxf(a, b);
        a = (n1.w - oy) * iy; b = (n2.z - oy) * iy; tn1 = fmaxf(tn1, fminf(a, b)); tf1 = fminf(tf1, fmaxf(a, b));
        a = (n2.x - oz) * iz; b = (n2.w - oz) * iz; tn1 = fmaxf(tn1, fminf(a, b)); tf1 = fminf(tf1, fmaxf(a, b));
        // conservative: widen [tn, tf] by a few ulp before comparing
        tn0 = fmaxf(tn0 - fabsf(tn0) * 4e-7f, 0.0f); tn1 = fmaxf(tn1 - fabsf(tn1) * 4e-7f, 0.0f);
        tf0 += fabsf(tf0) * 4e-7f; tf1 += fabsf(tf1) * 4e-7f;
        bool h0 = tn0 <= tf0 && tn0 <= tmax, h1 = tn1 <= tf1 && tn1 <= tmax;
        const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
        if (h0 && c0 < 0) { ++n_tests; tri_test(m, ~c0, o, d, tmin, best); h0 = false; }
        if (h1 && c1 < 0) { ++n_tests; tri_test(m, ~c1, o, d, tmin, best); h1 = false; }
        tmax = upper_float<R>(tmin);
        if (h0 && h1) {
            const bool zero_first = tn0 <= tn1;
            if (sp < kBvhStack) { stack[sp] = zero_first ? c1 : c0; stack_t[sp] = zero_first ? tn1 : tn0; ++sp; }
            cur = zero_first ? c0 : c1;
            continue;
        }
        if (h0) { cur = c0; continue; }
        if (h1) { cur = c1; continue; }
        bool found = false;
        while (sp > 0) {
            --sp;
            if (stack_t[sp] <= tmax) { cur = stack[sp]; found = true; break; }
        }
        if (!found) break;
    }
}

// test aid (DRTB_FLAG_NO_BVH): the linear scan the BVH must agree with
template <typename R>
__device__ __forceinline__ void brute_closest(const MeshView& m, V3<R> o, V3<R> d, R& tmin, int& best, uint32_t& n_tests)
{
    for (int i = 0; i < m.n_tris; ++i) { ++n_tests; tri_test(m, i, o, d, tmin, best); }
}

} // namespace drtb
