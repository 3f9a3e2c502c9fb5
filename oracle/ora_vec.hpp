// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under geometricvofext_b200/ may include, link or load this.
//
// Vector algebra with the operation order of OpenFOAM's VectorSpace/Vector
// templates (OF v2312, not vendored in the reference tree -- recalled):
//   a & b = a.x*b.x + a.y*b.y + a.z*b.z          (left-to-right sum)
//   a ^ b = (ay*bz - az*by, az*bx - ax*bz, ax*by - ay*bx)
//   mag(a) = sqrt(magSqr(a)),  magSqr(a) = ax*ax + ay*ay + az*az
//   s*a, a*s, a/s are component-wise (a/s is a DIVISION per component)
// Compiled with -ffp-contract=off so no multiply-add is ever fused.
#pragma once
#include <cmath>

namespace ora {

typedef double scalar;
typedef int label;

// OpenFOAM floating-point constants (src/OpenFOAM/primitives/Scalar, recalled)
static const scalar SMALL = 1.0e-15;
static const scalar VSMALL = 1.0e-300;
static const scalar ROOTVSMALL = 1.0e-150;
static const scalar GREAT = 1.0e+15;
static const scalar VGREAT = 1.0e+300;

inline scalar mag(scalar s) { return std::fabs(s); }
inline scalar sign(scalar s) { return (s >= 0) ? 1 : -1; }
inline scalar pos0(scalar s) { return (s >= 0) ? 1 : 0; }
inline scalar neg0(scalar s) { return (s <= 0) ? 1 : 0; }
inline scalar sqr(scalar s) { return s * s; }
inline scalar pow3(scalar s) { return s * sqr(s); }
inline scalar smax(scalar a, scalar b) { return (a > b) ? a : b; }  // Foam::max
inline scalar smin(scalar a, scalar b) { return (a < b) ? a : b; }  // Foam::min

struct vec {
    scalar x, y, z;
    vec() : x(0), y(0), z(0) {}
    vec(scalar X, scalar Y, scalar Z) : x(X), y(Y), z(Z) {}
    vec& operator+=(const vec& b) { x += b.x; y += b.y; z += b.z; return *this; }
    vec& operator-=(const vec& b) { x -= b.x; y -= b.y; z -= b.z; return *this; }
    vec& operator*=(scalar s) { x *= s; y *= s; z *= s; return *this; }
    vec& operator/=(scalar s) { x /= s; y /= s; z /= s; return *this; }
};
typedef vec point;

inline vec operator+(const vec& a, const vec& b) { return vec(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec operator-(const vec& a, const vec& b) { return vec(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec operator-(const vec& a) { return vec(-a.x, -a.y, -a.z); }
inline vec operator*(scalar s, const vec& a) { return vec(s * a.x, s * a.y, s * a.z); }
inline vec operator*(const vec& a, scalar s) { return vec(a.x * s, a.y * s, a.z * s); }
inline vec operator/(const vec& a, scalar s) { return vec(a.x / s, a.y / s, a.z / s); }
inline scalar operator&(const vec& a, const vec& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec operator^(const vec& a, const vec& b)
{
    return vec(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline scalar magSqr(const vec& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline scalar mag(const vec& a) { return std::sqrt(magSqr(a)); }

}  // namespace ora
