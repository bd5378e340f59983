// drt/camera.hpp — pinhole Camera (reference camera.hpp:10-70).
#pragma once
#include <cmath>
#include <cstddef>
#include <tuple>
#include "random.hpp"
#include "vector.hpp"

namespace drt {

template <typename T>
class Camera {
    std::size_t w_, h_;
    double vfov_;
    Vector<T, 3> eye_, fwd_, right_, up_;

public:
    Camera(std::size_t width, std::size_t height, double vfov = 1.3963, Vector<T, 3> eye = Vector<T, 3>(0),
           Vector<T, 3> forward = Vector<T, 3>{0, 0, -1}, Vector<T, 3> right = Vector<T, 3>{1, 0, 0},
           Vector<T, 3> up = Vector<T, 3>{0, 1, 0})
        : w_(width), h_(height), vfov_(vfov), eye_(eye), fwd_(forward), right_(right), up_(up) {}

    void look_at(Vector<T, 3> eye, Vector<T, 3> at, Vector<T, 3> up = Vector<T, 3>{0, 1, 0})
    {
        eye_ = eye;
        fwd_ = normalize(at - eye);
        right_ = normalize(cross(fwd_, up));
        up_ = cross(right_, fwd_);
    }

    std::size_t width() const { return w_; }
    std::size_t height() const { return h_; }
    double aspect() const { return double(w_) / h_; }
    double vfov() const { return vfov_; }
    Vector<T, 3> eye() const { return eye_; }
    const Vector<T, 3>& forward() const { return fwd_; }
    const Vector<T, 3>& right() const { return right_; }
    const Vector<T, 3>& up() const { return up_; }

    // HOST convenience: a jittered ray through pixel (x, y), row 0 at the top;
    // returns (direction, pdf = 1).  drt::render() generates its rays on the GPU.
    std::tuple<Vector<T, 3>, double> sample(std::size_t x, std::size_t y) const
    {
        const double half = std::tan(vfov_ / 2.);
        const double s = (x + random::uniform()) / w_;
        const double t = (y + random::uniform()) / h_;
        Vector<T, 3> d = fwd_;
        d += ((2. * s - 1.) * aspect() * half) * right_;
        d += ((2. * t - 1.) * half) * -up_;
        return std::make_tuple(normalize(d), 1.0);
    }
};

} // namespace drt
