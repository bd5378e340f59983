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
// lanes of a warp to the SAME scalar are added inside the warp first (the 32 samples of a pixel hit the same
// triangle at their first vertex, so every one of its three channels is a 32-way collision) and one lane issues
// one red.global.add.f64 for the group: __match_any_sync on the scalar's index among the lanes that reached this
// add together, the group's values summed in lane order through shuffles.  Distinct scalars (the usual case at
// deeper vertices) cost the match and one vote on top of the atomic.
// DRTB_FLAG_DETERMINISTIC (`fixed`): the same grouping, but every contribution is first rounded to 64-bit fixed point
// (2^-32) and the sums are integer sums -- associative, so the result does not depend on which lanes met in a warp or
// on the order the atomics land in; fixed_to_double_kernel converts the buffer in place after the last batch.
constexpr double kFixedScale = 4294967296.0;      // 2^32
struct AtomicSink {
    double* grad;
    bool fixed = false;
    template <typename R> __device__ __forceinline__ void add(int p, int c, R v)
    {
        const unsigned active = __activemask();
        const int key = 3 * p + c;
        const unsigned group = __match_any_sync(active, key);
        const int size = __popc(group);
        const int rounds = __reduce_max_sync(active, size);        // 1: no two lanes share a scalar
        const int lane = threadIdx.x & 31;
        const int first = __ffs(group) - 1;
        if (fixed) {
            const long long q = __double2ll_rn(double(v) * kFixedScale);
            long long sum = q;
            for (int r = 1; r < rounds; ++r) {
                const int src = r < size ? int(__fns(group, 0, r + 1)) : lane;
                const long long other = __shfl_sync(active, q, src);
                if (lane == first && r < size) sum += other;
            }
            if (lane == first) atomicAdd(reinterpret_cast<unsigned long long*>(grad) + key, (unsigned long long)sum);
            return;
        }
        double sum = double(v);
        for (int r = 1; r < rounds; ++r) {                          // warp-uniform trip count
            const int src = r < size ? int(__fns(group, 0, r + 1)) : lane;
            const double other = __shfl_sync(active, double(v), src);
            if (lane == first && r < size) sum += other;
        }
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
