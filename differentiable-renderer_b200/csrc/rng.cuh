// rng.cuh — counter-based per-(pixel, sample, slot) stream.
//
// Replaces the reference's only randomness, random::uniform() =
// double(rand())/RAND_MAX over ONE global sequential libc stream
// (include/drt/random.hpp:7-10).  Draw number `slot` of path `key` is
//     k = splitmix64(key * 0x100000001B3 + slot) mod (2^31 - 1)
// i.e. exactly an integer glibc's rand() could have returned, minus RAND_MAX
// itself (k <= RAND_MAX-1 keeps u < 1, which Pathtracer::trace needs when
// absorb == 1: pathtracer.hpp:128-130 would otherwise divide by p = 0).
// Paths replay deterministically from (key, slot) alone.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define DRTB_HD __host__ __device__ __forceinline__
#else
#define DRTB_HD inline
#endif

namespace drtb {

constexpr uint64_t kKeyMul   = 0x100000001B3ull;
constexpr uint64_t kSeedMul  = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kGolden   = 0x9E3779B97F4A7C15ull;   // splitmix64's increment
constexpr uint32_t kMersenne = 2147483647u;          // RAND_MAX on glibc

DRTB_HD uint64_t splitmix64(uint64_t x)
{
    x += kGolden;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// x mod (2^31 - 1) without a 64-bit division: 2^31 == 1 (mod M), so the
// 31-bit digits of x can simply be added.
DRTB_HD uint32_t mod_mersenne31(uint64_t x)
{
    x = (x & kMersenne) + (x >> 31);                 // < 2^33 + 2^31
    uint32_t y = uint32_t(x & kMersenne) + uint32_t(x >> 31);   // < 2^31 + 5
    return y >= kMersenne ? y - kMersenne : y;
}

// The same split in two: splitmix64(x) = splitmix64_mix(x + golden), so a path
// can carry ctr = key * kKeyMul + golden + slot and pay one 64-bit add per draw
// instead of two.
DRTB_HD uint64_t splitmix64_mix(uint64_t x)
{
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
DRTB_HD uint32_t stream_draw_ctr(uint64_t ctr) { return mod_mersenne31(splitmix64_mix(ctr)); }

// base = key * kKeyMul, hoisted once per path.
DRTB_HD uint32_t stream_draw_base(uint64_t base, uint32_t slot)
{
    return mod_mersenne31(splitmix64(base + slot));
}

DRTB_HD uint32_t stream_draw(uint64_t key, uint32_t slot)
{
    return stream_draw_base(key * kKeyMul, slot);
}

} // namespace drtb
