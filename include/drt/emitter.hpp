// drt/emitter.hpp — Emitter / AreaEmitter (reference emitter.hpp:7-25).
#pragma once
#include "vector.hpp"

namespace drt {

template <typename T>
class Emitter {
public:
    virtual ~Emitter() = default;
    virtual Vector<T, 3, true> emission() const = 0;
};

// Constant radiance over the surface, both sides, all directions.
template <typename T>
class AreaEmitter : public Emitter<T> {
    Vector<T, 3, true> radiance_;

public:
    AreaEmitter(Vector<T, 3, true> emission) : radiance_(emission) {}
    Vector<T, 3, true> emission() const override { return radiance_; }
};

} // namespace drt
