// drt/vector.hpp — value vectors and differentiable handles, source-compatible
// with the reference's include/drt/vector.hpp (same spellings: Vector<T,N>,
// Vector<T,N,true>, detach/grad/requires_grad/backward, dot/norm/normalize/
// cross/reflect) but built differently: one closure-carrying tape node instead
// of the reference's AutogradNode/ConstantNode/VariableNode/BackwardNode class
// tree (reference vector.hpp:120-213).
//
// On the GPU path none of this tape is used per sample: drt::render()
// (drt/render.hpp) flattens the scene, the CUDA adjoint kernel produces
// d(loss)/d(param) directly, and the result is added to the leaf's grad() --
// which is all a caller of `red.grad()` ever observed (reference
// vector.hpp:185-191).  The tape remains for host-side user math around the
// renderer (losses, custom backward functions as in the reference README).
#pragma once

#include <array>
#include <cmath>
#include <cstddef>
#include <functional>
#include <initializer_list>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <type_traits>
#include <typeinfo>
#include <utility>

namespace drt {

template <typename T, std::size_t N, bool Autograd = false>
class Vector;

// ----------------------------------------------------------------------------
// Plain value vector
// ----------------------------------------------------------------------------
template <typename T, std::size_t N>
class Vector<T, N, false> {
    std::array<T, N> c_{};

    template <typename F>
    Vector& zip(const Vector& o, F f)
    {
        for (std::size_t i = 0; i < N; ++i) c_[i] = f(c_[i], o.c_[i]);
        return *this;
    }
    template <typename F>
    Vector& map(F f)
    {
        for (std::size_t i = 0; i < N; ++i) c_[i] = f(c_[i]);
        return *this;
    }

public:
    using iterator = typename std::array<T, N>::iterator;
    using const_iterator = typename std::array<T, N>::const_iterator;

    Vector() = default;
    explicit Vector(T fill) { c_.fill(fill); }
    Vector(std::initializer_list<T> xs)
    {
        if (xs.size() != N) throw std::runtime_error("incorrect number of initializers for `Vector`");
        std::size_t i = 0;
        for (const T& x : xs) c_[i++] = x;
    }

    T& operator[](std::size_t i) { return c_[i]; }
    const T& operator[](std::size_t i) const { return c_[i]; }
    iterator begin() { return c_.begin(); }
    iterator end() { return c_.end(); }
    const_iterator begin() const { return c_.begin(); }
    const_iterator end() const { return c_.end(); }
    constexpr std::size_t size() const { return N; }
    T* data() { return c_.data(); }
    const T* data() const { return c_.data(); }

    Vector& operator+=(const Vector& o) { return zip(o, [](T a, T b) { return a + b; }); }
    Vector& operator-=(const Vector& o) { return zip(o, [](T a, T b) { return a - b; }); }
    Vector& operator*=(const Vector& o) { return zip(o, [](T a, T b) { return a * b; }); }
    Vector& operator/=(const Vector& o) { return zip(o, [](T a, T b) { return a / b; }); }
    Vector& operator*=(T s) { return map([s](T a) { return a * s; }); }
    Vector& operator/=(T s) { return map([s](T a) { return a / s; }); }
};

template <typename S, typename T>
using if_scalar_t = std::enable_if_t<std::is_convertible_v<S, T>>;

template <typename T, std::size_t N> Vector<T, N> operator+(Vector<T, N> a, const Vector<T, N>& b) { return a += b; }
template <typename T, std::size_t N> Vector<T, N> operator-(Vector<T, N> a, const Vector<T, N>& b) { return a -= b; }
template <typename T, std::size_t N> Vector<T, N> operator*(Vector<T, N> a, const Vector<T, N>& b) { return a *= b; }
template <typename T, std::size_t N> Vector<T, N> operator/(Vector<T, N> a, const Vector<T, N>& b) { return a /= b; }
template <typename T, std::size_t N, typename S, typename = if_scalar_t<S, T>>
Vector<T, N> operator*(Vector<T, N> a, S s) { return a *= T(s); }
template <typename T, std::size_t N, typename S, typename = if_scalar_t<S, T>>
Vector<T, N> operator*(S s, Vector<T, N> a) { return a *= T(s); }
template <typename T, std::size_t N, typename S, typename = if_scalar_t<S, T>>
Vector<T, N> operator/(Vector<T, N> a, S s) { return a /= T(s); }
// unary minus is a multiplication by -1, like the reference's (vector.hpp:320-325)
template <typename T, std::size_t N> Vector<T, N> operator-(const Vector<T, N>& a) { return T(-1) * a; }

template <typename T, std::size_t N>
T dot(const Vector<T, N>& a, const Vector<T, N>& b)
{
    T acc = T();                                   // left-to-right from T(), as the reference sums
    for (std::size_t i = 0; i < N; ++i) acc = acc + a[i] * b[i];
    return acc;
}
template <typename T, std::size_t N> T norm(const Vector<T, N>& a) { using std::sqrt; return sqrt(dot(a, a)); }
template <typename T, std::size_t N> Vector<T, N> normalize(const Vector<T, N>& a) { return a / norm(a); }
template <typename T>
Vector<T, 3> cross(const Vector<T, 3>& a, const Vector<T, 3>& b)
{
    return Vector<T, 3>{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
template <typename T, std::size_t N>
Vector<T, N> reflect(const Vector<T, N>& v, const Vector<T, N>& n) { return -v + (2 * dot(n, v)) * n; }

// ----------------------------------------------------------------------------
// Differentiable handle
// ----------------------------------------------------------------------------
namespace internal {

// One node type for leaves, constants and results alike.
template <typename T, std::size_t N>
struct TapeNode {
    Vector<T, N> value;
    Vector<T, N> grad = Vector<T, N>(T());          // leaves accumulate here (zero-initialised,
                                                    // unlike the reference's VariableNode::m_grad)
    bool leaf = false;                              // created with requires_grad = true
    std::function<void(const Vector<T, N>&)> pull;  // results: pushes a cotangent to the inputs

    bool tracked() const { return leaf || bool(pull); }
    void backward(const Vector<T, N>& g)
    {
        if (leaf) grad += g;
        else if (pull) pull(g);
    }
};

} // namespace internal

template <typename T, std::size_t N>
class Vector<T, N, true> {
    using Node = internal::TapeNode<T, N>;
    std::shared_ptr<Node> n_;

public:
    explicit Vector(T fill, bool requires_grad = false) : Vector(Vector<T, N>(fill), requires_grad) {}
    Vector(std::initializer_list<T> xs, bool requires_grad = false) : Vector(Vector<T, N>(xs), requires_grad) {}
    Vector(const Vector<T, N>& v, bool requires_grad = false) : n_(std::make_shared<Node>())
    {
        n_->value = v;
        n_->leaf = requires_grad;
    }
    // custom backward function: Vector<T,N,true>(value, [=](const Vector<T,N>& g) {...})
    template <typename Backward,
              typename = std::enable_if_t<std::is_invocable_v<const Backward&, const Vector<T, N>&>>>
    Vector(const Vector<T, N>& v, const Backward& backward) : n_(std::make_shared<Node>())
    {
        n_->value = v;
        n_->pull = backward;
    }

    T& operator[](std::size_t i) { return n_->value[i]; }
    const T& operator[](std::size_t i) const { return n_->value[i]; }
    constexpr std::size_t size() const { return N; }

    Vector<T, N>& detach() { return n_->value; }
    const Vector<T, N>& detach() const { return n_->value; }
    Vector<T, N>& grad()
    {
        if (!n_->leaf) throw std::runtime_error("Vector has no gradient (not a variable)");
        return n_->grad;
    }
    const Vector<T, N>& grad() const
    {
        if (!n_->leaf) throw std::runtime_error("Vector has no gradient (not a variable)");
        return n_->grad;
    }
    bool requires_grad() const { return n_->tracked(); }
    void backward(const Vector<T, N>& g) const { n_->backward(g); }

    // identity of the underlying parameter: copies of a handle alias one node
    // (what lets six shapes share `white`); drt::render() keys parameters on it
    const void* id() const { return n_.get(); }
    bool is_leaf() const { return n_->leaf; }

    Vector& operator+=(const Vector& o) { return *this = *this + o; }
    Vector& operator-=(const Vector& o) { return *this = *this - o; }
    Vector& operator*=(const Vector& o) { return *this = *this * o; }
    Vector& operator/=(const Vector& o) { return *this = *this / o; }
    Vector& operator*=(T s) { return *this = *this * s; }
    Vector& operator/=(T s) { return *this = *this / s; }
};

template <typename T, std::size_t N> Vector<T, N>& detach(Vector<T, N>& v) { return v; }
template <typename T, std::size_t N> const Vector<T, N>& detach(const Vector<T, N>& v) { return v; }
template <typename T, std::size_t N> Vector<T, N>& detach(Vector<T, N, true>& v) { return v.detach(); }
template <typename T, std::size_t N> const Vector<T, N>& detach(const Vector<T, N, true>& v) { return v.detach(); }
template <typename T, std::size_t N> constexpr bool requires_grad(const Vector<T, N>&) { return false; }
template <typename T, std::size_t N> bool requires_grad(const Vector<T, N, true>& v) { return v.requires_grad(); }
template <typename T, std::size_t N> void backward(Vector<T, N>&, const Vector<T, N>&) {}
template <typename T, std::size_t N> void backward(Vector<T, N, true>& v, const Vector<T, N>& g) { v.backward(g); }

namespace internal {
template <typename T, std::size_t N> Vector<T, N, true> lift(const Vector<T, N>& v) { return Vector<T, N, true>(v); }
template <typename T, std::size_t N> const Vector<T, N, true>& lift(const Vector<T, N, true>& v) { return v; }

// result = f(a, b) with cotangent rule `rule(g, a_value, b_value) -> (ga, gb)`
template <typename T, std::size_t N, typename A, typename B, typename Rule>
Vector<T, N, true> record2(const Vector<T, N>& value, const A& a, const B& b, Rule rule)
{
    if (!requires_grad(a) && !requires_grad(b)) return Vector<T, N, true>(value);
    Vector<T, N, true> ha = lift(a), hb = lift(b);
    return Vector<T, N, true>(value, [ha, hb, rule](const Vector<T, N>& g) {
        auto parts = rule(g, ha.detach(), hb.detach());
        ha.backward(parts.first);
        hb.backward(parts.second);
    });
}
} // namespace internal

#define DRT_BINARY_OP(OP, RULE)                                                                       \
    template <typename T, std::size_t N, bool A1, bool A2, typename = std::enable_if_t<A1 || A2>>     \
    Vector<T, N, true> operator OP(const Vector<T, N, A1>& a, const Vector<T, N, A2>& b)              \
    {                                                                                                 \
        using V = Vector<T, N>;                                                                       \
        return internal::record2<T, N>(detach(a) OP detach(b), a, b,                                  \
                                       [](const V& g, const V& x, const V& y) { (void)x; (void)y; return RULE; }); \
    }
DRT_BINARY_OP(+, std::make_pair(g, g))
DRT_BINARY_OP(-, std::make_pair(g, -g))
DRT_BINARY_OP(*, std::make_pair(y * g, x * g))
DRT_BINARY_OP(/, std::make_pair(g / y, -x * g / (y * y)))
#undef DRT_BINARY_OP

template <typename T, std::size_t N, typename S, typename = if_scalar_t<S, T>>
Vector<T, N, true> operator*(S s, Vector<T, N, true> v)
{
    Vector<T, N> r = T(s) * v.detach();
    if (!v.requires_grad()) return Vector<T, N, true>(r);
    const T k = T(s);
    return Vector<T, N, true>(r, [v, k](const Vector<T, N>& g) { v.backward(k * g); });
}
template <typename T, std::size_t N, typename S, typename = if_scalar_t<S, T>>
Vector<T, N, true> operator*(Vector<T, N, true> v, S s) { return s * v; }
template <typename T, std::size_t N, typename S, typename = if_scalar_t<S, T>>
Vector<T, N, true> operator/(Vector<T, N, true> v, S s)
{
    Vector<T, N> r = v.detach() / T(s);
    if (!v.requires_grad()) return Vector<T, N, true>(r);
    const T k = T(s);
    return Vector<T, N, true>(r, [v, k](const Vector<T, N>& g) { v.backward(g / k); });
}
template <typename T, std::size_t N> Vector<T, N, true> operator-(const Vector<T, N, true>& a) { return T(-1) * a; }

template <typename T, std::size_t N, bool Ag>
std::ostream& operator<<(std::ostream& os, const Vector<T, N, Ag>& v)
{
    os << "Vector<" << typeid(T).name() << ", " << N << (Ag ? ", true" : "") << ">{";
    for (std::size_t i = 0; i < N; ++i) os << (i ? ", " : "") << v[i];
    return os << "}";
}

} // namespace drt
