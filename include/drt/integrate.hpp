// drt/integrate.hpp — the Monte Carlo integration operator
// (reference integrate.hpp:9-66), for host-side user code.
//
//   integrate<T,N>(forward, sampler, n)         sum of forward(x)/pdf, taped
//   integrate<T,N>(forward, sampler, n, true)   value detached; backward()
//                                               re-samples with FRESH random
//                                               numbers (decorrelated gradient)
#pragma once
#include <cstddef>
#include <tuple>
#include "vector.hpp"

namespace drt {

template <typename T, std::size_t N, typename Forward, typename Sampler>
Vector<T, N, true> integrate(const Forward& forward, const Sampler& sampler, std::size_t n_samples,
                             bool unbiased = false)
{
    if (!unbiased) {
        Vector<T, N, true> total(T(0));
        for (std::size_t i = 0; i < n_samples; ++i) {
            auto drawn = sampler();
            total += forward(std::get<0>(drawn)) / std::get<1>(drawn);
        }
        return total;
    }
    Vector<T, N> value(T(0));
    for (std::size_t i = 0; i < n_samples; ++i) {
        auto drawn = sampler();
        value += detach(forward(std::get<0>(drawn))) / T(std::get<1>(drawn));
    }
    Forward fwd = forward;
    Sampler smp = sampler;
    return Vector<T, N, true>(value, [fwd, smp, n_samples](const Vector<T, N>& g) {
        for (std::size_t i = 0; i < n_samples; ++i) {
            auto drawn = smp();
            auto y = fwd(std::get<0>(drawn));
            backward(y, g / T(std::get<1>(drawn)));
        }
    });
}

} // namespace drt
