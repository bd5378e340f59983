// sinks.cuh — where the adjoint's contributions go: the gradient sinks of the render, wavefront and
// explicit-ray kernels (what VariableNode::backward's `m_grad += g` is upstream, vector.hpp:185-188).
#pragma once
#include "path.cuh"

namespace drtb {

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
// Large parameter sets (mesh scenes, per-triangle albedos: up to 3 * 2^20 scalars in HBM).  Contributions of the
// lanes of a warp to the SAME scalar are added inside the warp first (the samples of a pixel hit the same
// triangle at their first vertex, so every one of its three channels is a many-way collision) and one lane issues
// one red.global.add.f64 for the group: __match_any_sync on the scalar's index among the lanes that reached this
// add together, then POINTER JUMPING along each group's lanes -- every lane knows the next higher lane of its group,
// adds that lane's running sum and takes over its pointer, so after ceil(log2(largest group)) rounds the first lane
// of a group holds the group's sum (a first version walked the group rank by rank with __fns: up to 31 rounds of
// ~40 instructions, 59 % of wf_adjoint's instructions, profiles/r02_wf_adjoint_f64_compact_summary.txt).  Distinct
// scalars (the usual case at deeper vertices) cost the match and one vote on top of the atomic.
// DRTB_FLAG_DETERMINISTIC (`fixed`): the same grouping, but every contribution is first rounded to 64-bit fixed point
// (2^-32) and the sums are integer sums -- associative, so the result does not depend on which lanes met in a warp or
// on the order the atomics land in; fixed_to_double_kernel converts the buffer in place after the last batch.
constexpr double kFixedScale = 4294967296.0;      // 2^32
struct AtomicSink {
    double* grad;
    bool fixed = false;
    template <typename T> static __device__ __forceinline__ T group_sum(unsigned active, unsigned group, int rounds, int lane, T mine)
    {
        const unsigned above = group & ~((2u << lane) - 1u);       // lane 31: 2u << 31 == 0, nothing above
        int next = above ? __ffs(above) - 1 : -1;
        T sum = mine;
        const int steps = 32 - __clz(rounds - 1);                   // ceil(log2(rounds)), rounds >= 2: warp-uniform
        for (int s = 0; s < steps; ++s) {
            const int src = next >= 0 ? next : lane;
            const T other = __shfl_sync(active, sum, src);
            const int after = __shfl_sync(active, next, src);
            if (next >= 0) { sum += other; next = after; }
        }
        return sum;
    }
    template <typename R> __device__ __forceinline__ void add(int p, int c, R v)
    {
        const unsigned active = __activemask();
        const int key = 3 * p + c;
        const unsigned group = __match_any_sync(active, key);
        const int rounds = __reduce_max_sync(active, __popc(group));     // 1: no two lanes share a scalar
        const int lane = threadIdx.x & 31;
        const int first = __ffs(group) - 1;
        if (fixed) {
            long long sum = __double2ll_rn(double(v) * kFixedScale);
            if (rounds > 1) sum = group_sum<long long>(active, group, rounds, lane, sum);
            if (lane == first) atomicAdd(reinterpret_cast<unsigned long long*>(grad) + key, (unsigned long long)sum);
            return;
        }
        double sum = double(v);
        if (rounds > 1) sum = group_sum<double>(active, group, rounds, lane, sum);
        if (lane == first) atomicAdd(grad + key, sum);
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

} // namespace drtb
