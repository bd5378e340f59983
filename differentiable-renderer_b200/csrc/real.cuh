// real.cuh — the two arithmetic instantiations of the path kernels.
//
//   Real<double>: the parity instantiation.  The reference computes in IEEE
//                 double (src/render.cpp:22); every value here is a double
//                 carried to <= 2 ulp, so results differ from the reference
//                 only through FMA contraction, Newton-iterated rcp/rsqrt and
//                 the documented algebraic shortcuts (sin(asin x) = x,
//                 cos(2 pi u) via sincospi) -- all ~1e-15 relative.
//   Real<float> : the throughput instantiation (MUFU-based rcp/rsqrt/sqrt).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace drtb {

template <typename R> struct Real;

template <> struct Real<double> {
    static constexpr double kPi    = 3.14159265358979323846;   // constants.hpp:9
    static constexpr double kInvPi = 0.31830988618379067154;
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
    static __device__ __forceinline__ double abs(double a) { return ::fabs(a); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    // The compiler's IEEE double division / sqrt cost ~25-40 instructions each
    // with a divergent slow-path call.  These are MUFU-seeded (2^-22) Newton
    // iterations with no branches: <= 1-2 ulp, i.e. ~2e-16 relative, twelve
    // orders of magnitude inside the 1e-4 parity tolerance.  Operands here are
    // never subnormal (scene-scale geometry), x == 0 is handled where it can occur.
    static __device__ __forceinline__ double rcp(double x)
    {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));   // e0 <= 2^-23
        const double e = ::fma(-x, r, 1.0);
        return ::fma(r, ::fma(e, e, e), r);                      // r (1 + e + e^2): e0^3 = 2^-69
    }
    static __device__ __forceinline__ double div(double a, double b)
    {
        const double r = rcp(b), q = a * r;
        return ::fma(::fma(-q, b, a), r, q);                     // one residual step
    }
    static __device__ __forceinline__ double rsqrt(double x)
    {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); // e0 <= 2^-22
        const double e = ::fma(-x * y, y, 1.0);                  // 1 - x y^2
        return ::fma(y, ::fma(0.375, e, 0.5) * e, y);            // y (1 + e/2 + 3e^2/8): ~e0^3
    }
    static __device__ __forceinline__ double sqrt(double x)      // x >= 0
    {
        const double s = x * rsqrt(x);
        return x > 0.0 ? s : 0.0;                                // 0 * inf guard
    }
    // sin/cos(2*pi*u): the reference forms phi = 2*pi*u in double then calls
    // libm cos/sin (bxdf.hpp:74, 48-49); sincospi(2u) has exact range reduction
    // and differs from that by ~1 ulp of phi.
    static __device__ __forceinline__ void sincos2pi(double u, double* s, double* c) { ::sincospi(2.0 * u, s, c); }
    // random::uniform(): double(k) / RAND_MAX (random.hpp:9), correctly rounded:
    // q0 = RN(k/M) up to 1 ulp, one FMA residual step makes it exact (Markstein).
    static __device__ __forceinline__ double uniform(uint32_t k)
    {
        const double M = 2147483647.0, inv = 1.0 / 2147483647.0;
        double a = double(k);
        double q = a * inv;
        double r = ::fma(-q, M, a);
        return ::fma(r, inv, q);
    }
};

template <> struct Real<float> {
    static constexpr float kPi    = 3.14159265358979323846f;
    static constexpr float kInvPi = 0.31830988618379067154f;
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float abs(float a) { return ::fabsf(a); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
    static __device__ __forceinline__ float rcp(float x) { return __fdividef(1.0f, x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ float rsqrt(float x) { return ::rsqrtf(x); }
    static __device__ __forceinline__ float sqrt(float x)
    {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));    // MUFU.SQRT, 2^-23 relative
        return r;
    }
    static __device__ __forceinline__ void sincos2pi(float u, float* s, float* c) { ::sincospif(2.0f * u, s, c); }
    // Top 24 bits of the 31-bit draw: u in [0, 1 - 2^-24], never 1.0f
    // (float(k/2147483647.0) would round the top ~64 draws to 1 and make
    // pdf = cos(theta)/pi = 0).
    static __device__ __forceinline__ float uniform(uint32_t k)
    {
        return float(k >> 7) * (1.0f / 16777216.0f);
    }
};

} // namespace drtb
