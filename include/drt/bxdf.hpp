// drt/bxdf.hpp — BxDF / DiffuseBxDF / SpecularBxDF (reference bxdf.hpp:12-124).
//
// operator() and sample() are HOST conveniences kept for API compatibility;
// drt::render() and Pathtracer::trace() never call them -- they flatten the
// material (kind + parameter handle) and the CUDA kernels do the sampling.
// MirrorBxDF is not provided: it does not compile in the reference either
// (bxdf.hpp:135 returns a double as a Vector).
#pragma once
#include <array>
#include <cmath>
#include <tuple>
#include "constants.hpp"
#include "random.hpp"
#include "vector.hpp"

namespace drt {

enum class BxDFKind { Diffuse = 0, Specular = 1 };

template <typename T>
class BxDF {
public:
    virtual ~BxDF() = default;
    virtual Vector<T, 3, true> operator()(const Vector<T, 3>& normal, const Vector<T, 3>& dir_in,
                                          const Vector<T, 3>& dir_out) const = 0;
    virtual std::tuple<Vector<T, 3>, double> sample(const Vector<T, 3>& normal,
                                                    const Vector<T, 3>& dir_in) const = 0;
    // what the flattener needs (the reference keeps these private)
    virtual BxDFKind kind() const = 0;
    virtual const Vector<T, 3, true>& color() const = 0;
    virtual double exponent() const { return 0.0; }
};

namespace internal {

// Orthonormal-ish frame around `normal` (which is NOT normalised here, exactly
// like the reference: the Cornell box's green wall has a non-unit normal).
template <typename T>
std::array<Vector<T, 3>, 3> make_frame(const Vector<T, 3>& normal)
{
    const bool use_x = std::abs(double(normal[0])) < std::abs(double(normal[1]));
    Vector<T, 3> axis = use_x ? Vector<T, 3>{1, 0, 0} : Vector<T, 3>{0, 1, 0};
    Vector<T, 3> tangent = normalize(axis - normal * dot(axis, normal));
    Vector<T, 3> bitangent = normalize(cross(normal, tangent));
    return {tangent, bitangent, normal};
}

template <typename T>
Vector<T, 3> angle_to_dir(double theta, double phi, const std::array<Vector<T, 3>, 3>& f)
{
    const double s = std::sin(theta);
    return (std::cos(phi) * s) * f[0] + (std::sin(phi) * s) * f[1] + std::cos(theta) * f[2];
}

} // namespace internal

template <typename T>
class DiffuseBxDF : public BxDF<T> {
    Vector<T, 3, true> albedo_;

public:
    DiffuseBxDF(const Vector<T, 3, true>& color) : albedo_(color) {}

    Vector<T, 3, true> operator()(const Vector<T, 3>&, const Vector<T, 3>&, const Vector<T, 3>&) const override
    {
        return albedo_ / pi;
    }
    // cosine-weighted hemisphere: theta = asin(sqrt(u1)), phi = 2 pi u2, pdf = cos(theta)/pi
    std::tuple<Vector<T, 3>, double> sample(const Vector<T, 3>& normal, const Vector<T, 3>&) const override
    {
        const double theta = std::asin(std::sqrt(random::uniform()));
        const double phi = 2 * pi * random::uniform();
        return std::make_tuple(internal::angle_to_dir(theta, phi, internal::make_frame(normal)),
                               std::cos(theta) / pi);
    }
    BxDFKind kind() const override { return BxDFKind::Diffuse; }
    const Vector<T, 3, true>& color() const override { return albedo_; }
};

// Normalised Blinn-Phong-like lobe around the half vector (reference
// bxdf.hpp:85-124).  The CUDA path samples and evaluates it on the device
// (DRTB_SPECULAR, csrc/path.cuh specular_sample); these host methods mirror it.
template <typename T>
class SpecularBxDF : public BxDF<T> {
    Vector<T, 3, true> tint_;
    double shininess_;

public:
    SpecularBxDF(const Vector<T, 3, true>& color, double exponent) : tint_(color), shininess_(exponent) {}

    Vector<T, 3, true> operator()(const Vector<T, 3>& normal, const Vector<T, 3>& dir_in,
                                  const Vector<T, 3>& dir_out) const override
    {
        const double c = double(dot(normal, normalize(dir_in + dir_out)));
        const double lobe = (shininess_ + 2) / (2 * pi) * std::pow(c, shininess_) * std::sqrt(1 - c * c);
        return lobe * tint_;
    }
    std::tuple<Vector<T, 3>, double> sample(const Vector<T, 3>& normal, const Vector<T, 3>& dir_in) const override
    {
        const double theta = std::acos(std::sqrt(std::pow(random::uniform(), 2 / (shininess_ + 2))));
        const double phi = 2 * pi * random::uniform();
        Vector<T, 3> h = internal::angle_to_dir(theta, phi, internal::make_frame(normal));
        if (dot(h, dir_in) < 0) h = reflect(h, normal);
        const double pdf = (shininess_ + 2) / (2 * pi) * std::pow(std::cos(theta), shininess_ + 1) * std::sin(theta);
        return std::make_tuple(reflect(dir_in, h), pdf);
    }
    BxDFKind kind() const override { return BxDFKind::Specular; }
    const Vector<T, 3, true>& color() const override { return tint_; }
    double exponent() const override { return shininess_; }
};

} // namespace drt
