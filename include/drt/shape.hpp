// drt/shape.hpp — Shape / Plane / Sphere (reference shape.hpp:11-111).
//
// intersect() and normal() are HOST conveniences with the reference's exact
// semantics (t > 0 acceptance, sphere quadratic with a == 1, plane normal
// returned un-normalised); the GPU path flattens shapes through describe().
#pragma once
#include <cmath>
#include <memory>
#include "../drtb.h"
#include "bxdf.hpp"
#include "emitter.hpp"
#include "vector.hpp"

namespace drt {

template <typename T>
class Shape {
    std::shared_ptr<BxDF<T>> surface_;
    std::shared_ptr<Emitter<T>> light_;

public:
    Shape(std::shared_ptr<BxDF<T>> bxdf = nullptr, std::shared_ptr<Emitter<T>> emitter = nullptr)
        : surface_(std::move(bxdf)), light_(std::move(emitter)) {}
    virtual ~Shape() = default;

    virtual bool intersect(Vector<T, 3> orig, Vector<T, 3> dir, double& t) const = 0;
    virtual Vector<T, 3> normal(Vector<T, 3> point) const = 0;
    // geometry only: fills type and v[4] of the flattened primitive (drtb.h)
    virtual void describe(drtb_prim& out) const = 0;

    BxDF<T>* bxdf() { return surface_.get(); }
    Emitter<T>* emitter() { return light_.get(); }
    const BxDF<T>* bxdf() const { return surface_.get(); }
    const Emitter<T>* emitter() const { return light_.get(); }
};

// { p : dot(p, normal) == offset }
template <typename T>
class Plane : public Shape<T> {
    Vector<T, 3> n_;
    double d_;

public:
    Plane(Vector<T, 3> normal, double offset, std::shared_ptr<BxDF<T>> bxdf = nullptr,
          std::shared_ptr<Emitter<T>> emitter = nullptr)
        : Shape<T>(std::move(bxdf), std::move(emitter)), n_(normal), d_(offset) {}

    bool intersect(Vector<T, 3> orig, Vector<T, 3> dir, double& t) const override
    {
        t = double((dot(orig, n_) - d_) / dot(dir, -n_));
        return t > 0;
    }
    Vector<T, 3> normal(Vector<T, 3>) const override { return n_; }
    void describe(drtb_prim& out) const override
    {
        out.type = DRTB_PLANE;
        out.v[0] = double(n_[0]); out.v[1] = double(n_[1]); out.v[2] = double(n_[2]); out.v[3] = d_;
    }
    const Vector<T, 3>& plane_normal() const { return n_; }
    double offset() const { return d_; }
};

template <typename T>
class Sphere : public Shape<T> {
    Vector<T, 3> c_;
    double r_;

public:
    Sphere(Vector<T, 3> center, double radius, std::shared_ptr<BxDF<T>> bxdf = nullptr,
           std::shared_ptr<Emitter<T>> emitter = nullptr)
        : Shape<T>(std::move(bxdf), std::move(emitter)), c_(center), r_(radius) {}

    // nearest positive root of |o + t d - c|^2 = r^2 with the quadratic's leading
    // coefficient taken as 1 (the reference never divides by |d|^2)
    bool intersect(Vector<T, 3> orig, Vector<T, 3> dir, double& t) const override
    {
        const Vector<T, 3> oc = orig - c_;
        const double b = 2 * double(dot(oc, dir));
        const double disc = b * b - 4 * (double(dot(oc, oc)) - r_ * r_);
        if (disc < 0) return false;
        const double root = std::sqrt(disc);
        const double near_t = (-b - root) / 2, far_t = (-b + root) / 2;
        if (near_t > 0) { t = near_t; return true; }
        if (far_t > 0) { t = far_t; return true; }
        return false;
    }
    Vector<T, 3> normal(Vector<T, 3> point) const override { return normalize(point - c_); }
    void describe(drtb_prim& out) const override
    {
        out.type = DRTB_SPHERE;
        out.v[0] = double(c_[0]); out.v[1] = double(c_[1]); out.v[2] = double(c_[2]); out.v[3] = r_;
    }
    const Vector<T, 3>& center() const { return c_; }
    double radius() const { return r_; }
};

} // namespace drt
