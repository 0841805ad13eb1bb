// madrona::math for the B200 engine.  Mirrors the public surface of the
// reference's include/madrona/math.hpp:20-385 (names, member order, operator
// set) so simulator sources compile unchanged, and keeps each formula's
// operation order (math.inl) so float results stay within 1e-4 rel of the
// CPU oracle.  Compiled by NVRTC (simulator code), nvcc (engine kernels) and
// g++ (host tools); no fast-math intrinsics: 1/sqrtf, not rsqrtf.
#pragma once

#include <cstdint>
#include <cmath>
#include <cfloat>

#include <madrona/types.hpp>
#include <madrona/macros.hpp>

namespace madrona {
namespace math {

struct Vector2;
struct Vector3;
struct Vector4;
struct Quat;
struct Mat3x3;
struct Mat3x4;

constexpr inline float pi {3.14159265358979323846264338327950288f};
constexpr inline float pi_d2 {pi / 2.f};
constexpr inline float pi_m2 {pi * 2.f};

MB2_HD inline constexpr float toRadians(float degrees)
{
    constexpr float mult = pi / 180.f;
    return mult * degrees;
}

MB2_HD inline float sqr(float x) { return x * x; }

MB2_HD inline bool solveQuadraticUnsafe(float a, float b, float c,
                                        float *t1, float *t2)
{
    float det = b * b - 4.f * a * c;
    if (det < 0.f) return false;
    float s = sqrtf(det);
    float r = 1.f / (2.f * a);
    *t1 = (-b - s) * r;
    *t2 = (-b + s) * r;
    return true;
}

struct Vector2 {
    float x;
    float y;

    MB2_HD float dot(const Vector2 &o) const { return x * o.x + y * o.y; }
    MB2_HD float length2() const { return x * x + y * y; }
    MB2_HD float length() const { return sqrtf(length2()); }
    MB2_HD float invLength() const { return 1.f / sqrtf(length2()); }

    MB2_HD float &operator[](CountT i) { return i == 0 ? x : y; }
    MB2_HD float operator[](CountT i) const { return i == 0 ? x : y; }

    MB2_HD constexpr Vector2 &operator+=(const Vector2 &o) { x += o.x; y += o.y; return *this; }
    MB2_HD constexpr Vector2 &operator-=(const Vector2 &o) { x -= o.x; y -= o.y; return *this; }
    MB2_HD constexpr Vector2 &operator+=(float o) { x += o; y += o; return *this; }
    MB2_HD constexpr Vector2 &operator-=(float o) { x -= o; y -= o; return *this; }
    MB2_HD constexpr Vector2 &operator*=(float o) { x *= o; y *= o; return *this; }
    MB2_HD constexpr Vector2 &operator/=(float o) { float r = 1.f / o; x *= r; y *= r; return *this; }

    MB2_HD static Vector2 min(Vector2 a, Vector2 b) { return { fminf(a.x, b.x), fminf(a.y, b.y) }; }
    MB2_HD static Vector2 max(Vector2 a, Vector2 b) { return { fmaxf(a.x, b.x), fmaxf(a.y, b.y) }; }
};

MB2_HD constexpr inline Vector2 operator-(Vector2 v) { return { -v.x, -v.y }; }
MB2_HD constexpr inline Vector2 operator+(Vector2 a, const Vector2 &b) { a += b; return a; }
MB2_HD constexpr inline Vector2 operator-(Vector2 a, const Vector2 &b) { a -= b; return a; }
MB2_HD constexpr inline Vector2 operator+(Vector2 a, float b) { a += b; return a; }
MB2_HD constexpr inline Vector2 operator-(Vector2 a, float b) { a -= b; return a; }
MB2_HD constexpr inline Vector2 operator*(Vector2 a, float b) { a *= b; return a; }
MB2_HD constexpr inline Vector2 operator/(Vector2 a, float b) { a /= b; return a; }
MB2_HD constexpr inline Vector2 operator+(float a, Vector2 b) { return b + a; }
MB2_HD constexpr inline Vector2 operator-(float a, Vector2 b) { return -b + a; }
MB2_HD constexpr inline Vector2 operator*(float a, Vector2 b) { return b * a; }
MB2_HD constexpr inline Vector2 operator/(float a, Vector2 b) { return { a / b.x, a / b.y }; }

struct Vector3 {
    float x;
    float y;
    float z;

    MB2_HD float dot(const Vector3 &o) const { return x * o.x + y * o.y + z * o.z; }
    MB2_HD Vector3 cross(const Vector3 &o) const
    {
        return { y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x };
    }

    // Two unit vectors perpendicular to *this (which must be normalized).
    MB2_HD inline void frame(Vector3 *a, Vector3 *b) const;

    MB2_HD float length2() const { return x * x + y * y + z * z; }
    MB2_HD float length() const { return sqrtf(length2()); }
    MB2_HD float invLength() const { return 1.f / sqrtf(length2()); }

    MB2_HD inline float distance(const Vector3 &o) const;
    MB2_HD inline float distance2(const Vector3 &o) const;

    [[nodiscard]] MB2_HD inline Vector3 normalize() const;

    MB2_HD constexpr Vector2 xy() const { return { x, y }; }
    MB2_HD constexpr Vector2 yz() const { return { y, z }; }
    MB2_HD constexpr Vector2 xz() const { return { x, z }; }
    MB2_HD constexpr Vector2 yx() const { return { y, x }; }
    MB2_HD constexpr Vector2 zy() const { return { z, y }; }
    MB2_HD constexpr Vector2 zx() const { return { z, x }; }

    MB2_HD float &operator[](CountT i) { return i == 0 ? x : (i == 1 ? y : z); }
    MB2_HD float operator[](CountT i) const { return i == 0 ? x : (i == 1 ? y : z); }

    MB2_HD constexpr Vector3 &operator+=(const Vector3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    MB2_HD constexpr Vector3 &operator-=(const Vector3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    MB2_HD constexpr Vector3 &operator+=(float o) { x += o; y += o; z += o; return *this; }
    MB2_HD constexpr Vector3 &operator-=(float o) { x -= o; y -= o; z -= o; return *this; }
    MB2_HD constexpr Vector3 &operator*=(float o) { x *= o; y *= o; z *= o; return *this; }
    MB2_HD constexpr Vector3 &operator/=(float o) { float r = 1.f / o; x *= r; y *= r; z *= r; return *this; }

    MB2_HD static Vector3 min(Vector3 a, Vector3 b) { return { fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z) }; }
    MB2_HD static Vector3 max(Vector3 a, Vector3 b) { return { fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z) }; }

    MB2_HD static constexpr Vector3 zero() { return { 0, 0, 0 }; }
    MB2_HD static constexpr Vector3 one() { return { 1, 1, 1 }; }
    MB2_HD static constexpr Vector3 all(float v) { return { v, v, v }; }
};

MB2_HD constexpr inline Vector3 operator-(Vector3 v) { return { -v.x, -v.y, -v.z }; }
MB2_HD constexpr inline Vector3 operator+(Vector3 a, const Vector3 &b) { a += b; return a; }
MB2_HD constexpr inline Vector3 operator-(Vector3 a, const Vector3 &b) { a -= b; return a; }
MB2_HD constexpr inline Vector3 operator+(Vector3 a, float b) { a += b; return a; }
MB2_HD constexpr inline Vector3 operator-(Vector3 a, float b) { a -= b; return a; }
MB2_HD constexpr inline Vector3 operator*(Vector3 a, float b) { a *= b; return a; }
MB2_HD constexpr inline Vector3 operator/(Vector3 a, float b) { a /= b; return a; }
MB2_HD constexpr inline Vector3 operator+(float a, Vector3 b) { return b + a; }
MB2_HD constexpr inline Vector3 operator-(float a, Vector3 b) { return -b + a; }
MB2_HD constexpr inline Vector3 operator*(float a, Vector3 b) { return b * a; }
MB2_HD constexpr inline Vector3 operator/(float a, Vector3 b) { return { a / b.x, a / b.y, a / b.z }; }

MB2_HD inline float dot(Vector2 a, Vector2 b) { return a.dot(b); }
MB2_HD inline float dot(Vector3 a, Vector3 b) { return a.dot(b); }
MB2_HD inline Vector3 cross(Vector3 a, Vector3 b) { return a.cross(b); }
MB2_HD inline Vector3 normalize(Vector3 v) { return v.normalize(); }
MB2_HD inline Vector3 reflect(Vector3 direction, Vector3 normal)
{
    return direction - 2.f * dot(direction, normal) * normal;
}

MB2_HD void Vector3::frame(Vector3 *a, Vector3 *b) const
{
    // NB: like the reference (math.inl:256-267) the results are not renormalised.
    Vector3 pick = fabsf(x) < 0.8 ? Vector3 { 1, 0, 0 } : Vector3 { 0, 1, 0 };
    *a = cross(pick);
    *b = cross(*a);
}

MB2_HD float Vector3::distance(const Vector3 &o) const { return (*this - o).length(); }
MB2_HD float Vector3::distance2(const Vector3 &o) const { return (*this - o).length2(); }
MB2_HD Vector3 Vector3::normalize() const { return *this * invLength(); }

struct Vector4 {
    float x;
    float y;
    float z;
    float w;

    MB2_HD Vector3 xyz() const { return { x, y, z }; }
    MB2_HD float &operator[](CountT i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    MB2_HD float operator[](CountT i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    MB2_HD static Vector4 fromVec3W(Vector3 v, float w) { return { v.x, v.y, v.z, w }; }
    MB2_HD Vector4 operator*(float s) const { return { x * s, y * s, z * s, w * s }; }
    MB2_HD Vector4 operator+(const Vector4 &o) const { return { x + o.x, y + o.y, z + o.z, w + o.w }; }
    MB2_HD static constexpr Vector4 zero() { return { 0, 0, 0, 0 }; }
    MB2_HD static constexpr Vector4 one() { return { 1, 1, 1, 1 }; }
};

struct Quat {
    float w;
    float x;
    float y;
    float z;

    MB2_HD float length2() const { return w * w + x * x + y * y + z * z; }
    MB2_HD float length() const { return sqrtf(length2()); }
    MB2_HD float invLength() const { return 1.f / sqrtf(length2()); }

    [[nodiscard]] MB2_HD Quat normalize() const
    {
        float s = invLength();
        return { w * s, x * s, y * s, z * s };
    }
    [[nodiscard]] MB2_HD Quat inv() const { return { w, -x, -y, -z }; }

    // v + 2 (w (q x v) + q x (q x v))
    MB2_HD Vector3 rotateVec(Vector3 v) const
    {
        Vector3 q { x, y, z };
        Vector3 qv = cross(q, v);
        Vector3 qqv = cross(q, qv);
        return v + 2.f * ((qv * w) + qqv);
    }

    MB2_HD static Quat angleAxis(float angle, Vector3 normal)
    {
        float c = cosf(angle / 2.f);
        float s = sinf(angle / 2.f);
        return { c, normal.x * s, normal.y * s, normal.z * s };
    }
    MB2_HD static Quat fromAngularVec(Vector3 v) { return { 0, v.x, v.y, v.z }; }
    MB2_HD static inline Quat fromBasis(Vector3 a, Vector3 b, Vector3 c);
    MB2_HD static constexpr Quat id() { return { 1.f, 0.f, 0.f, 0.f }; }

    MB2_HD Quat &operator+=(Quat o) { w += o.w; x += o.x; y += o.y; z += o.z; return *this; }
    MB2_HD Quat &operator-=(Quat o) { w -= o.w; x -= o.x; y -= o.y; z -= o.z; return *this; }
    MB2_HD inline Quat &operator*=(Quat o);
    MB2_HD Quat &operator*=(float f) { w *= f; x *= f; y *= f; z *= f; return *this; }
};

MB2_HD inline Quat operator+(Quat a, Quat b) { return a += b; }
MB2_HD inline Quat operator-(Quat a, Quat b) { return a -= b; }
MB2_HD inline Quat operator*(Quat a, Quat b)
{
    return {
        (a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z),
        (a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y),
        (a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x),
        (a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w),
    };
}
MB2_HD inline Quat operator*(Quat a, float b) { return a *= b; }
MB2_HD inline Quat operator*(float b, Quat a) { return a *= b; }
MB2_HD Quat &Quat::operator*=(Quat o) { return *this = (*this * o); }

// Rotation matrix (columns a,b,c) -> quaternion, largest-component branch
// (the classic trace method; reference: math.inl:751-835).
MB2_HD Quat Quat::fromBasis(Vector3 a, Vector3 b, Vector3 c)
{
    float t[4] = { a.x + b.y + c.z, a.x - b.y - c.z,
                   b.y - a.x - c.z, c.z - a.x - b.y };
    int big = 0;
    float big_v = t[0];
    for (int i = 1; i < 4; i++) {
        if (t[i] > big_v) { big_v = t[i]; big = i; }
    }
    float v = sqrtf(big_v + 1.f) * 0.5f;
    float m = 0.25f / v;
    switch (big) {
    case 0: return { v, (b.z - c.y) * m, (c.x - a.z) * m, (a.y - b.x) * m };
    case 1: return { (b.z - c.y) * m, v, (a.y + b.x) * m, (c.x + a.z) * m };
    case 2: return { (c.x - a.z) * m, (a.y + b.x) * m, v, (b.z + c.y) * m };
    default: return { (a.y - b.x) * m, (c.x + a.z) * m, (b.z + c.y) * m, v };
    }
}

struct Diag3x3 {
    float d0;
    float d1;
    float d2;

    MB2_HD Diag3x3 inv() const { return { 1.f / d0, 1.f / d1, 1.f / d2 }; }
    MB2_HD static Diag3x3 fromVec(Vector3 v) { return { v.x, v.y, v.z }; }
    MB2_HD static constexpr Diag3x3 uniform(float s) { return { s, s, s }; }
    MB2_HD static constexpr Diag3x3 id() { return { 1.f, 1.f, 1.f }; }

    MB2_HD Diag3x3 &operator*=(Diag3x3 o) { d0 *= o.d0; d1 *= o.d1; d2 *= o.d2; return *this; }
    MB2_HD Diag3x3 &operator*=(float o) { d0 *= o; d1 *= o; d2 *= o; return *this; }
    MB2_HD Diag3x3 &operator/=(float o) { d0 /= o; d1 /= o; d2 /= o; return *this; }

    MB2_HD float &operator[](CountT i) { return i == 0 ? d0 : (i == 1 ? d1 : d2); }
    MB2_HD float operator[](CountT i) const { return i == 0 ? d0 : (i == 1 ? d1 : d2); }
};

MB2_HD inline Diag3x3 operator*(Diag3x3 a, Diag3x3 b) { a *= b; return a; }
MB2_HD inline Diag3x3 operator*(Diag3x3 a, float b) { a *= b; return a; }
MB2_HD inline Diag3x3 operator*(float a, Diag3x3 b) { b *= a; return b; }
MB2_HD inline Vector3 operator*(Diag3x3 d, Vector3 v) { return { d.d0 * v.x, d.d1 * v.y, d.d2 * v.z }; }
MB2_HD inline Diag3x3 operator/(Diag3x3 a, float b) { a /= b; return a; }
MB2_HD inline Diag3x3 operator/(float a, Diag3x3 b) { return { a / b.d0, a / b.d1, a / b.d2 }; }

struct Mat3x3 {
    struct Transpose {
        const Mat3x3 *src;
        MB2_HD inline Vector3 operator[](CountT i) const;
    };

    Vector3 cols[3];

    MB2_HD float determinant() const
    {
        Vector3 c0 = cols[0], c1 = cols[1], c2 = cols[2];
        return c0.x * (c1.y * c2.z - c2.y * c1.z) -
               c0.y * (c1.x * c2.z - c2.x * c1.z) +
               c0.z * (c1.x * c2.y - c2.x * c1.y);
    }
    MB2_HD Transpose transpose() const { return { this }; }

    MB2_HD static inline Mat3x3 fromQuat(Quat r);
    MB2_HD static inline Mat3x3 fromRS(Quat r, Diag3x3 s);

    MB2_HD Vector3 &operator[](CountT i) { return cols[i]; }
    MB2_HD Vector3 operator[](CountT i) const { return cols[i]; }

    MB2_HD Mat3x3 &operator+=(const Mat3x3 &o) { cols[0] += o.cols[0]; cols[1] += o.cols[1]; cols[2] += o.cols[2]; return *this; }
    MB2_HD Mat3x3 &operator-=(const Mat3x3 &o) { cols[0] -= o.cols[0]; cols[1] -= o.cols[1]; cols[2] -= o.cols[2]; return *this; }

    MB2_HD Vector3 operator*(Vector3 v) const { return cols[0] * v.x + cols[1] * v.y + cols[2] * v.z; }
    MB2_HD Mat3x3 operator*(const Mat3x3 &o) const { return { *this * o.cols[0], *this * o.cols[1], *this * o.cols[2] }; }
    MB2_HD Mat3x3 &operator*=(const Mat3x3 &o) { return *this = (*this * o); }
    MB2_HD Mat3x3 &operator*=(float s) { cols[0] *= s; cols[1] *= s; cols[2] *= s; return *this; }
};

MB2_HD Vector3 Mat3x3::Transpose::operator[](CountT i) const
{
    return { src->cols[0][i], src->cols[1][i], src->cols[2][i] };
}

MB2_HD inline Vector3 operator*(Mat3x3::Transpose t, Vector3 v)
{
    return { dot(t.src->cols[0], v), dot(t.src->cols[1], v), dot(t.src->cols[2], v) };
}

MB2_HD Mat3x3 Mat3x3::fromQuat(Quat r)
{
    float x2 = r.x * r.x, y2 = r.y * r.y, z2 = r.z * r.z;
    float xz = r.x * r.z, xy = r.x * r.y, yz = r.y * r.z;
    float wx = r.w * r.x, wy = r.w * r.y, wz = r.w * r.z;
    return {{
        { 1.f - 2.f * (y2 + z2), 2.f * (xy + wz), 2.f * (xz - wy) },
        { 2.f * (xy - wz), 1.f - 2.f * (x2 + z2), 2.f * (yz + wx) },
        { 2.f * (xz + wy), 2.f * (yz - wx), 1.f - 2.f * (x2 + y2) },
    }};
}

MB2_HD Mat3x3 Mat3x3::fromRS(Quat r, Diag3x3 s)
{
    float x2 = r.x * r.x, y2 = r.y * r.y, z2 = r.z * r.z;
    float xz = r.x * r.z, xy = r.x * r.y, yz = r.y * r.z;
    float wx = r.w * r.x, wy = r.w * r.y, wz = r.w * r.z;
    Diag3x3 ds = 2.f * s;
    return {{
        { s.d0 - ds.d0 * (y2 + z2), ds.d0 * (xy + wz), ds.d0 * (xz - wy) },
        { ds.d1 * (xy - wz), s.d1 - ds.d1 * (x2 + z2), ds.d1 * (yz + wx) },
        { ds.d2 * (xz + wy), ds.d2 * (yz - wx), s.d2 - ds.d2 * (x2 + y2) },
    }};
}

MB2_HD inline Mat3x3 operator+(Mat3x3 a, const Mat3x3 &b) { return (a += b); }
MB2_HD inline Mat3x3 operator-(Mat3x3 a, const Mat3x3 &b) { return (a -= b); }
MB2_HD inline Mat3x3 operator*(const Mat3x3 &m, Diag3x3 d) { return {{ m[0] * d.d0, m[1] * d.d1, m[2] * d.d2 }}; }
MB2_HD inline Mat3x3 operator*(Diag3x3 d, const Mat3x3 &m) { return {{ d * m[0], d * m[1], d * m[2] }}; }
MB2_HD inline Mat3x3 operator*(Mat3x3 a, Mat3x3::Transpose b) { return { a * b[0], a * b[1], a * b[2] }; }
MB2_HD inline Mat3x3 operator*(Mat3x3::Transpose a, Mat3x3 b) { return { a * b[0], a * b[1], a * b[2] }; }
MB2_HD inline Mat3x3 operator*(float s, const Mat3x3 &m) { return {{ s * m[0], s * m[1], s * m[2] }}; }
MB2_HD inline Mat3x3 operator*(const Mat3x3 &m, float s) { return s * m; }
MB2_HD inline Mat3x3 operator/(const Mat3x3 &m, float s) { return {{ m[0] / s, m[1] / s, m[2] / s }}; }

MB2_HD inline Mat3x3 outerProduct(Vector3 a, Vector3 b)
{
    return {{ a * b.x, a * b.y, a * b.z }};
}

// Symmetric 3x3: diag = (m11,m22,m33), off = (m12,m13,m23).
struct Symmetric3x3 {
    Vector3 diag;
    Vector3 off;

    MB2_HD static inline Symmetric3x3 AAT(Mat3x3 A);
    MB2_HD static inline Symmetric3x3 AXAT(Mat3x3 A, Symmetric3x3 X);
    MB2_HD static Symmetric3x3 vvT(Vector3 v)
    {
        return { { v.x * v.x, v.y * v.y, v.z * v.z }, { v.x * v.y, v.x * v.z, v.y * v.z } };
    }

    MB2_HD Vector3 operator[](CountT i) const
    {
        return i == 0 ? Vector3 { diag.x, off.x, off.y } :
               (i == 1 ? Vector3 { off.x, diag.y, off.z } :
                         Vector3 { off.y, off.z, diag.z });
    }

    MB2_HD Symmetric3x3 &operator+=(const Symmetric3x3 &o) { diag += o.diag; off += o.off; return *this; }
    MB2_HD Symmetric3x3 &operator-=(const Symmetric3x3 &o) { diag -= o.diag; off -= o.off; return *this; }
    MB2_HD inline Symmetric3x3 &operator*=(const Symmetric3x3 &o);
    MB2_HD Symmetric3x3 &operator*=(float s) { diag *= s; off *= s; return *this; }
};

MB2_HD Symmetric3x3 Symmetric3x3::AAT(Mat3x3 A)
{
    // rows of A
    Vector3 r0 { A[0].x, A[1].x, A[2].x };
    Vector3 r1 { A[0].y, A[1].y, A[2].y };
    Vector3 r2 { A[0].z, A[1].z, A[2].z };
    return { { dot(r0, r0), dot(r1, r1), dot(r2, r2) },
             { dot(r0, r1), dot(r0, r2), dot(r1, r2) } };
}

MB2_HD Symmetric3x3 Symmetric3x3::AXAT(Mat3x3 A, Symmetric3x3 X)
{
    // rows of A, columns of (symmetric) X; result = A X A^T.
    Vector3 r0 { A[0].x, A[1].x, A[2].x };
    Vector3 r1 { A[0].y, A[1].y, A[2].y };
    Vector3 r2 { A[0].z, A[1].z, A[2].z };
    Vector3 x0 = X[0], x1 = X[1], x2 = X[2];
    Vector3 t0 { dot(r0, x0), dot(r0, x1), dot(r0, x2) };   // row 0 of A X
    Vector3 t1 { dot(r1, x0), dot(r1, x1), dot(r1, x2) };
    Vector3 t2 { dot(r2, x0), dot(r2, x1), dot(r2, x2) };
    return { { dot(r0, t0), dot(r1, t1), dot(r2, t2) },
             { dot(r1, t0), dot(r2, t0), dot(r2, t1) } };
}

MB2_HD inline Symmetric3x3 operator+(Symmetric3x3 a, Symmetric3x3 b) { return a += b; }
MB2_HD inline Symmetric3x3 operator-(Symmetric3x3 a, Symmetric3x3 b) { return a -= b; }
MB2_HD inline Symmetric3x3 operator*(Symmetric3x3 a, float b) { return a *= b; }
MB2_HD inline Symmetric3x3 operator*(float a, Symmetric3x3 b) { return b *= a; }
MB2_HD inline Vector3 operator*(Symmetric3x3 m, Vector3 v)
{
    return { m.diag.x * v.x + m.off.x * v.y + m.off.y * v.z,
             m.off.x * v.x + m.diag.y * v.y + m.off.z * v.z,
             m.off.y * v.x + m.off.z * v.y + m.diag.z * v.z };
}

struct Mat3x4 {
    Vector3 cols[4];

    MB2_HD Vector3 txfmPoint(Vector3 p) const
    {
        return cols[0] * p.x + cols[1] * p.y + cols[2] * p.z + cols[3];
    }
    MB2_HD Vector3 txfmDir(Vector3 p) const
    {
        return cols[0] * p.x + cols[1] * p.y + cols[2] * p.z;
    }
    MB2_HD Mat3x4 compose(const Mat3x4 &o) const
    {
        return {{ txfmDir(o.cols[0]), txfmDir(o.cols[1]), txfmDir(o.cols[2]), txfmPoint(o.cols[3]) }};
    }

    MB2_HD static Mat3x4 fromRows(Vector4 r0, Vector4 r1, Vector4 r2)
    {
        return {{ { r0.x, r1.x, r2.x }, { r0.y, r1.y, r2.y },
                  { r0.z, r1.z, r2.z }, { r0.w, r1.w, r2.w } }};
    }
    MB2_HD static Mat3x4 fromTRS(Vector3 t, Quat r, Diag3x3 s = { 1.f, 1.f, 1.f })
    {
        Mat3x3 rs = Mat3x3::fromRS(r, s);
        return {{ rs[0], rs[1], rs[2], t }};
    }
    MB2_HD static constexpr Mat3x4 identity()
    {
        return {{ { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 }, { 0, 0, 0 } }};
    }
};

struct Mat4x4 {
    Vector4 cols[4];

    MB2_HD Vector4 txfmPoint(Vector4 p) const
    {
        return cols[0] * p.x + cols[1] * p.y + cols[2] * p.z + cols[3] * p.w;
    }
    MB2_HD Mat4x4 compose(const Mat4x4 &o) const
    {
        return {{ txfmPoint(o.cols[0]), txfmPoint(o.cols[1]), txfmPoint(o.cols[2]), txfmPoint(o.cols[3]) }};
    }
    MB2_HD static constexpr Mat4x4 identity()
    {
        return {{ { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } }};
    }
};

struct AABB {
    Vector3 pMin;
    Vector3 pMax;

    MB2_HD float surfaceArea() const
    {
        Vector3 d = pMax - pMin;
        return 2.f * (d.x * d.y + d.x * d.z + d.y * d.z);
    }
    MB2_HD Vector3 centroid() const { return 0.5f * (pMin + pMax); }
    MB2_HD int maxDimension() const
    {
        Vector3 d = pMax - pMin;
        if (d.x > d.y && d.x > d.z) return 0;
        return d.y > d.z ? 1 : 2;
    }
    MB2_HD bool overlaps(const AABB &o) const
    {
        return pMin.x < o.pMax.x && o.pMin.x < pMax.x &&
               pMin.y < o.pMax.y && o.pMin.y < pMax.y &&
               pMin.z < o.pMax.z && o.pMin.z < pMax.z;
    }
    MB2_HD bool intersects(const AABB &o) const
    {
        return pMin.x <= o.pMax.x && o.pMin.x <= pMax.x &&
               pMin.y <= o.pMax.y && o.pMin.y <= pMax.y &&
               pMin.z <= o.pMax.z && o.pMin.z <= pMax.z;
    }
    MB2_HD bool contains(const AABB &o) const
    {
        return pMin.x <= o.pMin.x && pMin.y <= o.pMin.y && pMin.z <= o.pMin.z &&
               pMax.x >= o.pMax.x && pMax.y >= o.pMax.y && pMax.z >= o.pMax.z;
    }
    MB2_HD bool contains(const Vector3 &p) const
    {
        return pMin.x <= p.x && pMin.y <= p.y && pMin.z <= p.z &&
               pMax.x >= p.x && pMax.y >= p.y && pMax.z >= p.z;
    }
    // if / else-if per axis exactly as the reference (math.inl:1649-1668):
    // a point below pMin never also raises pMax.
    MB2_HD void expand(const Vector3 &p)
    {
        for (int i = 0; i < 3; i++) {
            if (p[i] < pMin[i]) pMin[i] = p[i];
            else if (p[i] > pMax[i]) pMax[i] = p[i];
        }
    }
    MB2_HD float distance2(const AABB &o) const
    {
        float d2 = 0.f;
        for (int i = 0; i < 3; i++) {
            float diff = fmaxf(pMin[i], o.pMin[i]) - fminf(pMax[i], o.pMax[i]);
            if (diff > 0) d2 += diff * diff;
        }
        return d2;
    }
    MB2_HD Vector3 offset(const Vector3 &p) const
    {
        Vector3 o = p - pMin;
        for (int i = 0; i < 3; i++) {
            if (pMax[i] > pMin[i]) o[i] /= pMax[i] - pMin[i];
        }
        return o;
    }

    // Slab test, reference math.inl:1670-1735 ("max of mins, min of maxes" with
    // fminf / fmaxf, which skip NaNs, and a NON-strict final comparison): kept
    // expression for expression -- the NaN / infinity / equality cases decide
    // which leaves a grazing ray visits.
    MB2_HD bool rayIntersects(Vector3 ray_o, Diag3x3 inv_ray_d,
                              float ray_t_min, float ray_t_max,
                              float &hit_t, float &far_t)
    {
        Vector3 t_lower = inv_ray_d * (pMin - ray_o);
        Vector3 t_upper = inv_ray_d * (pMax - ray_o);
        Vector3 mins = Vector3::min(t_lower, t_upper);
        Vector3 maxes = Vector3::max(t_lower, t_upper);
        float t_box_min = fmaxf(mins.x, fmaxf(mins.y, fmaxf(mins.z, ray_t_min)));
        float t_box_max = fminf(maxes.x, fminf(maxes.y, fminf(maxes.z, ray_t_max)));
        if (t_box_min <= t_box_max) {
            hit_t = t_box_min;
            far_t = t_box_max;
            return true;
        }
        return false;
    }
    MB2_HD bool rayIntersects(Vector3 ray_o, Diag3x3 inv_ray_d,
                              float ray_t_min, float ray_t_max)
    {
        float a, b;
        return rayIntersects(ray_o, inv_ray_d, ray_t_min, ray_t_max, a, b);
    }

    // Box after translate/rotate/scale (RTCD p.86; reference math.inl:1737-1769):
    // accumulate per matrix entry the min/max of entry*pMin / entry*pMax.
    [[nodiscard]] MB2_HD AABB applyTRS(const Vector3 &translation,
                                       const Quat &rotation,
                                       const Diag3x3 &scale = { 1, 1, 1 }) const
    {
        Mat3x3 m = Mat3x3::fromRS(rotation, scale);
        AABB out { translation, translation };
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) {
                float e = m[j][i] * pMin[j];
                float f = m[j][i] * pMax[j];
                if (e < f) {
                    out.pMin[i] += e;
                    out.pMax[i] += f;
                } else {
                    out.pMin[i] += f;
                    out.pMax[i] += e;
                }
            }
        }
        return out;
    }

    MB2_HD float operator[](CountT i) const
    {
        return i < 3 ? pMin[i] : pMax[i - 3];
    }

    MB2_HD static AABB invalid()
    {
        return { { FLT_MAX, FLT_MAX, FLT_MAX }, { -FLT_MAX, -FLT_MAX, -FLT_MAX } };
    }
    MB2_HD static AABB point(const Vector3 &p) { return { p, p }; }
    MB2_HD static AABB merge(const AABB &a, const AABB &b)
    {
        return { Vector3::min(a.pMin, b.pMin), Vector3::max(a.pMax, b.pMax) };
    }
};

struct AABB2D {
    Vector2 pMin;
    Vector2 pMax;
    MB2_HD Vector2 centroid() const { return (pMin + pMax) / 2.f; }
    MB2_HD float area() const { Vector2 d = pMax - pMin; return d.x * d.y; }
};

constexpr inline Vector3 up { 0, 0, 1 };
constexpr inline Vector3 fwd { 0, 1, 0 };
constexpr inline Vector3 right { 1, 0, 0 };

}

constexpr inline math::Vector3 worldUp = math::up;
constexpr inline math::Vector3 worldFwd = math::fwd;
constexpr inline math::Vector3 worldRight = math::right;

}
