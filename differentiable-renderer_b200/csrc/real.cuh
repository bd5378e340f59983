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
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

namespace drtb {

template <typename R> struct Real;

// Double constants that are not encodable as a 32-bit immediate live in the
// constant bank: an FP64 instruction can take c[bank][offset] as an operand
// directly, whereas a literal costs two UMOVs each time it is rematerialised
// (34 per ray segment for the sincos polynomials alone in the first builds).
struct ConstF64 {
    double sin_c[8], cos_c[8];
    double half_pi, pi, inv_pi, inv_m, m, origin_eps, k375;     // k375 = 3/8 of the Newton corrections (rsqrt, sqrt)
    double tab_s[3], tab_c[2], tab_step;         // sincos_tab: -1/7!, 1/5!, -1/3! ; -1/6!, 1/4! ; 2 pi / M
};
// (static: every translation unit of the library carries its own copy -- no relocatable device code)
static __constant__ ConstF64 kC64 = {
    {2.8114572543455206e-15, -7.6471637318198164e-13, 1.6059043836821613e-10, -2.5052108385441720e-08,
     2.7557319223985893e-06, -1.9841269841269841e-04, 8.3333333333333332e-03, -1.6666666666666666e-01},
    {-1.5619206968586225e-16, 4.7794773323873853e-14, -1.1470745597729725e-11, 2.0876756987868100e-09,
     -2.7557319223985888e-07, 2.4801587301587302e-05, -1.3888888888888889e-03, 4.1666666666666664e-02},
    1.5707963267948966, 3.14159265358979323846, 0.31830988618379067154, 1.0 / 2147483647.0, 2147483647.0, 1e-3, 0.375,
    {-1.984126984126984e-04, 8.333333333333333e-03, -1.6666666666666666e-01},
    {-1.388888888888889e-03, 4.1666666666666664e-02}, 2.925836159896768e-09};

// sin/cos(2 pi k / M) for the 31-bit draw k (M = 2^31 - 1): the draw's top bits
// pick (sin, cos)(a_i) from a table, a_i = 2 pi (i 2^23) / M, the low bits are a
// small angle delta = 2 pi j / M, |delta| <= pi/256, and the angle-sum formulas
// need only degree-7 / degree-6 Taylor polynomials of it (truncation < 2e-20).
// No quadrant logic, no float-to-int conversion: 12 FP64 instructions and one
// 16-byte shared-memory load against ~24 + 11 integer ones for sincos2pi().
// The table is filled on the host in long double (drtb_create) and staged into
// shared memory per block (BlockScene).
constexpr int kSinCosShift = 23;
constexpr int kSinCosEntries = (1 << (31 - kSinCosShift)) + 1;      // 257 (index 256: k + 2^22 carries)
static __device__ double2 g_sincos_tab[kSinCosEntries];
// Fills THIS translation unit's copy of the table: (sin, cos)(2 pi i 2^23 / M) in long double on the host.
static inline cudaError_t upload_sincos_tab()
{
    static double2 tab[kSinCosEntries];
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int i = 0; i < kSinCosEntries; ++i) {
        const long double ang = two_pi * ((long double)i * (long double)(1u << kSinCosShift)) / 2147483647.0L;
        tab[i] = make_double2(double(sinl(ang)), double(cosl(ang)));
    }
    return cudaMemcpyToSymbol(g_sincos_tab, tab, sizeof(tab));
}
template <typename R> struct SinCosTab { };                          // float: sincospif
template <> struct SinCosTab<double> { double2 t[kSinCosEntries]; };

template <> struct Real<double> {
    static __device__ __forceinline__ double pi() { return kC64.pi; }           // constants.hpp:9
    static __device__ __forceinline__ double inv_pi() { return kC64.inv_pi; }
    static __device__ __forceinline__ double origin_eps() { return kC64.origin_eps; }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
    static __device__ __forceinline__ double abs(double a) { return ::fabs(a); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    // Sign logic on the high word: an FP64 compare/negate/select occupies the
    // (half-rate) FP64 pipe, an integer op on the sign word does not.  These
    // treat positive subnormals below 2^-1022 as zero -- nothing on this path
    // gets within 300 orders of magnitude of that.
    static __device__ __forceinline__ bool is_pos(double x) { return __double2hiint(x) > 0; }
    static __device__ __forceinline__ bool is_nonzero(double x) { return (__double2hiint(x) << 1) != 0; }
    // g > 0 ? -h : h   (sign(g) == + flips h)
    static __device__ __forceinline__ double flip_if_pos(double h, double g)
    {
        const int gh = __double2hiint(g);
        const int flip = (gh > 0) ? int(0x80000000u) : 0;
        return __hiloint2double(__double2hiint(h) ^ flip, __double2loint(h));
    }
    // a * b, except that an exact +-0 comes out negative (-2^-1000): for every product above 2^-947 in magnitude the
    // addend is below half an ulp and the FMA rounds it away (a t that small cannot occur: segments start 1e-3 off the
    // surface they leave).  Lets the plane tests reject t = 0 (shape.hpp:55, t > 0) without a test of their own
    // (Closest<double, true>::offer_nz).  The low word of the constant is zero, so it is an immediate operand.
    static __device__ __forceinline__ double mul_nz(double a, double b) { return ::fma(a, b, -0x1p-1000); }
    // x with its sign flipped iff c > 0 and h < 0 (sphere_test: the nearer root), from the two sign words.
    // c == +0 counts as positive and h == -0 as negative: see sphere_test for why that cannot matter.
    static __device__ __forceinline__ double neg_if_outside_ahead(double x, double c, double h)
    {
        const uint32_t flip = uint32_t(__double2hiint(h)) & ~uint32_t(__double2hiint(c)) & 0x80000000u;
        return __hiloint2double(int(uint32_t(__double2hiint(x)) ^ flip), __double2loint(x));
    }
    static __device__ __forceinline__ double select(bool c, double a, double b)
    {
        return __hiloint2double(c ? __double2hiint(a) : __double2hiint(b), c ? __double2loint(a) : __double2loint(b));
    }
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
        return ::fma(y, ::fma(kC64.k375, e, 0.5) * e, y);            // y (1 + e/2 + 3e^2/8): ~e0^3
    }
    // x >= 0; NaN for x < 0.  g = x y0 ~ sqrt(x) is corrected directly,
    // g (1 + e/2 + 3e^2/8) with e = 1 - g y0: five FP64 instructions, no final
    // multiply.  x == 0: the seed is taken at max(x, 2^-1022) -- an unsigned max on
    // the high word, which leaves negative x (high word >= 2^31) alone -- so y0 is
    // finite, g = 0, e = 1 and the result is an exact 0 without a select.
    static __device__ __forceinline__ double sqrt(double x)
    {
        const uint32_t xh = max(uint32_t(__double2hiint(x)), 0x00100000u);
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(int(xh), __double2loint(x))));   // e0 <= 2^-22
        const double g = x * y;
        const double e = ::fma(-g, y, 1.0);
        return ::fma(g, ::fma(kC64.k375, e, 0.5) * e, g);
    }
    static __device__ __forceinline__ double pow(double a, double b) { return ::pow(a, b); }   // SpecularBxDF only
    // sin/cos(2*pi*u), u in [0, 1).  The reference forms phi = 2*pi*u in double and
    // calls libm cos/sin (bxdf.hpp:74, 48-49).  Here: x = 4u = q + r with q the
    // nearest integer and |r| <= 1/2, i.e. phi = q*pi/2 + r*pi/2 exactly; Taylor
    // polynomials of sin/cos(r*pi/2) on |r*pi/2| <= pi/4 (truncation < 5e-17),
    // then a quadrant rotation.  ~35 instructions against ~64 for sincospi().
    static __device__ __forceinline__ void sincos2pi(double u, double* s, double* c)
    {
        const double x = 4.0 * u;
        const int q = __double2int_rn(x);
        const double r = x - double(q);                          // exact, |r| <= 0.5
        const double t = r * kC64.half_pi;                       // r * pi/2, |t| <= pi/4
        const double t2 = t * t;
        double ps = kC64.sin_c[0];                               //  1/17!, -1/15!, ... -1/3!
#pragma unroll
        for (int i = 1; i < 8; ++i) ps = ::fma(ps, t2, kC64.sin_c[i]);
        const double sn = ::fma(ps * t2, t, t);
        double pc = kC64.cos_c[0];                               // -1/18!, 1/16!, ... 1/4!
#pragma unroll
        for (int i = 1; i < 8; ++i) pc = ::fma(pc, t2, kC64.cos_c[i]);
        pc = ::fma(pc, t2, -0.5);
        const double cs = ::fma(pc, t2, 1.0);
        // rotate by q quarter turns: (sin, cos)(a + q pi/2)
        const bool swap = q & 1;
        const double a = select(swap, cs, sn), b = select(swap, sn, cs);
        const int sflip = (q & 2) ? int(0x80000000u) : 0;                 // sin: - for q = 2, 3
        const int cflip = ((q + 1) & 2) ? int(0x80000000u) : 0;           // cos: - for q = 1, 2
        *s = __hiloint2double(__double2hiint(a) ^ sflip, __double2loint(a));
        *c = __hiloint2double(__double2hiint(b) ^ cflip, __double2loint(b));
    }
    static constexpr bool kTable = true;
    static __device__ __forceinline__ void sincos_tab(const SinCosTab<double>& tab, uint32_t k, double* s, double* c)
    {
        const uint32_t i = (k + (1u << (kSinCosShift - 1))) >> kSinCosShift;
        const int j = int(k - (i << kSinCosShift));              // [-2^22, 2^22)
        const double2 sc = tab.t[i];
        const double d = double(j) * kC64.tab_step;
        const double d2 = d * d;
        double ps = ::fma(d2, kC64.tab_s[0], kC64.tab_s[1]);
        ps = ::fma(ps, d2, kC64.tab_s[2]);
        const double sn = ::fma(ps * d2, d, d);                  // sin(delta)
        double pc = ::fma(d2, kC64.tab_c[0], kC64.tab_c[1]);
        pc = ::fma(pc, d2, -0.5);
        const double cm1 = pc * d2;                              // cos(delta) - 1
        *s = ::fma(sc.x, cm1, ::fma(sc.y, sn, sc.x));
        *c = ::fma(sc.y, cm1, ::fma(-sc.x, sn, sc.y));
    }
    // random::uniform(): double(k) / RAND_MAX (random.hpp:9), correctly rounded:
    // q0 = RN(k/M) up to 1 ulp, one FMA residual step makes it exact (Markstein).
    static __device__ __forceinline__ double uniform(uint32_t k)
    {
        const double M = kC64.m, inv = kC64.inv_m;
        double a = double(k);
        double q = a * inv;
        double r = ::fma(-q, M, a);
        return ::fma(r, inv, q);
    }
    // The same to 1 ulp (k * RN(1/M)) for draws that feed continuous quantities
    // (pixel jitter, theta, phi); Russian roulette keeps the exact one.
    static __device__ __forceinline__ double uniform_fast(uint32_t k) { return double(k) * kC64.inv_m; }
};

template <> struct Real<float> {
    static __device__ __forceinline__ float pi() { return 3.14159265358979323846f; }
    static __device__ __forceinline__ float inv_pi() { return 0.31830988618379067154f; }
    static __device__ __forceinline__ float origin_eps() { return 1e-3f; }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float abs(float a) { return ::fabsf(a); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
    static __device__ __forceinline__ bool is_pos(float x) { return x > 0.0f; }
    static __device__ __forceinline__ bool is_nonzero(float x) { return x != 0.0f; }
    static __device__ __forceinline__ float flip_if_pos(float h, float g) { return g > 0.0f ? -h : h; }
    static __device__ __forceinline__ float select(bool c, float a, float b) { return c ? a : b; }
    static __device__ __forceinline__ float mul_nz(float a, float b) { return a * b; }
    static __device__ __forceinline__ float neg_if_outside_ahead(float x, float c, float h) { return (c > 0.0f && h < 0.0f) ? -x : x; }
    // one MUFU.RCP (2^-23 relative); __fdividef(1, x) wraps it in range handling that cost the float kernels 4 % of
    // their instructions (profiles/r02_mixed_f32_pass_summary.txt)
    static __device__ __forceinline__ float rcp(float x)
    {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float div(float a, float b) { return a * rcp(b); }
    static __device__ __forceinline__ float rsqrt(float x) { return ::rsqrtf(x); }
    static __device__ __forceinline__ float sqrt(float x)
    {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));    // MUFU.SQRT, 2^-23 relative
        return r;
    }
    static __device__ __forceinline__ float pow(float a, float b) { return ::powf(a, b); }
    static __device__ __forceinline__ void sincos2pi(float u, float* s, float* c) { ::sincospif(2.0f * u, s, c); }
    static constexpr bool kTable = false;
    static __device__ __forceinline__ void sincos_tab(const SinCosTab<float>&, uint32_t k, float* s, float* c)
    {
        sincos2pi(uniform(k), s, c);
    }
    // Top 24 bits of the 31-bit draw: u in [0, 1 - 2^-24], never 1.0f
    // (float(k/2147483647.0) would round the top ~64 draws to 1 and make
    // pdf = cos(theta)/pi = 0).
    static __device__ __forceinline__ float uniform(uint32_t k)
    {
        return float(k >> 7) * (1.0f / 16777216.0f);
    }
    static __device__ __forceinline__ float uniform_fast(uint32_t k) { return uniform(k); }
};

} // namespace drtb
