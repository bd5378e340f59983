// drt/random.hpp — host-side random::uniform() (reference random.hpp:7-10) plus
// the key source of the counter-based GPU stream.
//
// The reference draws everything from ONE global sequential libc rand().  The
// GPU path cannot (and should not): every path owns the stream
//     k(key, slot) = splitmix64(key * 0x100000001B3 + slot) mod (2^31 - 1)
// (drtb_stream_draw in drtb.h).  random::uniform() below is kept for host-side
// user code and for the host conveniences Camera::sample / BxDF::sample.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>

namespace drt { namespace random {

inline double uniform() { return double(std::rand()) / RAND_MAX; }

// Keys handed to single-ray Pathtracer::trace calls: a process-wide counter
// far away from the per-pixel keys drt::render() uses ((y*W + x)*spp + i).
inline std::uint64_t next_ray_key()
{
    static std::atomic<std::uint64_t> counter{0};
    return (std::uint64_t(1) << 62) + counter.fetch_add(1, std::memory_order_relaxed);
}

} } // namespace drt::random
