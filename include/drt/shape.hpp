// drt/shape.hpp — Shape / Plane / Sphere (reference shape.hpp:11-111), plus Triangle.
//
// intersect() and normal() are HOST conveniences with the reference's exact
// semantics (t > 0 acceptance, sphere quadratic with a == 1, plane normal
// returned un-normalised); the GPU path flattens shapes through describe().
//
// Triangle<T> is NEW: the reference's extension point is Shape<T> (shape.hpp:11-35) and it
// ships planes and spheres only.  Semantics as fixed in include/drtb.h (drtb_mesh):
// Moller-Trumbore, both sides hit, t > 0 with no epsilon, inclusive barycentric bounds,
// the unit geometric normal used as given.  gpu::flatten gathers the triangles of a Scene
// into one drtb_mesh, for which the library builds its BVH on the GPU.
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
    // triangles go to the mesh (drtb_mesh) instead: v0, v1, v2 as 9 doubles
    virtual bool describe_triangle(double*) const { return false; }

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

// Triangle (v0, v1, v2); NEW relative to the reference, see the header comment.
template <typename T>
class Triangle : public Shape<T> {
    Vector<T, 3> v0_, v1_, v2_, e1_, e2_;      // the edges are v1 - v0 and v2 - v0, as the library forms them

public:
    Triangle(Vector<T, 3> v0, Vector<T, 3> v1, Vector<T, 3> v2, std::shared_ptr<BxDF<T>> bxdf = nullptr,
             std::shared_ptr<Emitter<T>> emitter = nullptr)
        : Shape<T>(std::move(bxdf), std::move(emitter)), v0_(v0), v1_(v1), v2_(v2), e1_(v1 - v0), e2_(v2 - v0) {}

    bool intersect(Vector<T, 3> orig, Vector<T, 3> dir, double& t) const override
    {
        const Vector<T, 3> p = cross(dir, e2_);
        const double det = double(dot(e1_, p));
        if (det == 0) return false;
        const double inv = 1.0 / det;
        const Vector<T, 3> tv = orig - v0_;
        const double u = double(dot(tv, p)) * inv;
        const Vector<T, 3> q = cross(tv, e1_);
        const double v = double(dot(dir, q)) * inv;
        t = double(dot(e2_, q)) * inv;
        return u >= 0 && u <= 1 && v >= 0 && u + v <= 1 && t > 0;
    }
    Vector<T, 3> normal(Vector<T, 3>) const override { return normalize(cross(e1_, e2_)); }
    void describe(drtb_prim& out) const override { out.type = 2; }          // not an analytic primitive: see describe_triangle
    bool describe_triangle(double* v) const override
    {
        for (int c = 0; c < 3; ++c) {
            v[c] = double(v0_[c]); v[3 + c] = double(v1_[c]); v[6 + c] = double(v2_[c]);
        }
        return true;
    }
};

} // namespace drt
