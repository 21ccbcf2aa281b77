// dxmc/vectormath.hpp — the vector helpers OpenDXMC calls: rotate, cross, scale, add
// (R:src/libopendxmc/dxmc_specialization.cpp:80-88, R:src/libopendxmc/beamactorcontainer.cpp:52,66,188).
#pragma once
#include <array>
#include <cmath>
namespace dxmc::vectormath {
using Vec3 = std::array<double, 3>;
inline Vec3 add(const Vec3& a, const Vec3& b) { return { a[0] + b[0], a[1] + b[1], a[2] + b[2] }; }
inline Vec3 subtract(const Vec3& a, const Vec3& b) { return { a[0] - b[0], a[1] - b[1], a[2] - b[2] }; }
inline Vec3 scale(const Vec3& a, double s) { return { a[0] * s, a[1] * s, a[2] * s }; }
inline Vec3 scale(double s, const Vec3& a) { return scale(a, s); }
inline double dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline Vec3 cross(const Vec3& a, const Vec3& b) { return { a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0] }; }
inline Vec3 cross(const std::array<Vec3, 2>& c) { return cross(c[0], c[1]); }
inline double length(const Vec3& a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalized(const Vec3& a)
{
    const double l = length(a);
    return l > 0 ? scale(a, 1.0 / l) : a;
}
inline void normalize(Vec3& a) { a = normalized(a); }
// Rodrigues rotation of v about `axis` by `angle` [rad]
inline Vec3 rotate(const Vec3& v, const Vec3& axis, double angle)
{
    const Vec3 k = normalized(axis);
    const double c = std::cos(angle), s = std::sin(angle);
    const Vec3 kv = cross(k, v);
    const double kd = dot(k, v) * (1 - c);
    return { v[0] * c + kv[0] * s + k[0] * kd, v[1] * c + kv[1] * s + k[1] * kd, v[2] * c + kv[2] * s + k[2] * kd };
}
inline std::size_t argmin3(const Vec3& a)
{
    std::size_t m = 0;
    for (std::size_t i = 1; i < 3; ++i)
        if (std::abs(a[i]) < std::abs(a[m]))
            m = i;
    return m;
}
}
