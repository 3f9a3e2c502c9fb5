// svof_math.cuh -- FP64 3-vector algebra for the SimPLIC kernels (sm_100a).
//
// The parity target (BASELINE.json: interface-cell set and cut-face topology
// bit-exact, alpha within 1e-12 per step) is only reachable if the device
// evaluates the reference's expressions in the reference's order, so:
//   * every product/sum is written out in OpenFOAM's VectorSpace order
//     ((x + y) + z; cross product component order as in Vector.H), and
//   * this translation unit is compiled with -fmad=false: no DFMA contraction.
// There is nothing to vectorise or tensorise here: the work is scalar FP64 on
// a few thousand polygons; throughput comes from occupancy, not from tcgen05.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace svof {

// OpenFOAM constants (OF v2312 primitives/Scalar)
#define SV_SMALL 1.0e-15
#define SV_VSMALL 1.0e-300
#define SV_ROOTVSMALL 1.0e-150
#define SV_GREAT 1.0e+15
#define SV_VGREAT 1.0e+300
#define SV_TSMALL (10.0 * SV_SMALL)  /* cutFace.C:154, cutCell.C:355,618 */
#define SV_ATOL (100.0 * SV_SMALL)   /* advectionTemplates.C:125,228      */

struct d3 {
    double x, y, z;
};

__host__ __device__ __forceinline__ d3 mk3(double x, double y, double z)
{
    d3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
__host__ __device__ __forceinline__ d3 zero3() { return mk3(0.0, 0.0, 0.0); }
__host__ __device__ __forceinline__ d3 operator+(const d3& a, const d3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ d3 operator-(const d3& a, const d3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ d3 operator-(const d3& a) { return mk3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ d3 operator*(double s, const d3& a) { return mk3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ d3 operator*(const d3& a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ d3 operator/(const d3& a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ void operator+=(d3& a, const d3& b)
{
    a.x += b.x;
    a.y += b.y;
    a.z += b.z;
}
__host__ __device__ __forceinline__ void operator-=(d3& a, const d3& b)
{
    a.x -= b.x;
    a.y -= b.y;
    a.z -= b.z;
}
__host__ __device__ __forceinline__ void operator/=(d3& a, double s)
{
    a.x /= s;
    a.y /= s;
    a.z /= s;
}
// a & b
__host__ __device__ __forceinline__ double dot(const d3& a, const d3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// a ^ b
__host__ __device__ __forceinline__ d3 cross(const d3& a, const d3& b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ double magSqr(const d3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__host__ __device__ __forceinline__ double mag(const d3& a) { return sqrt(magSqr(a)); }

__host__ __device__ __forceinline__ double sgn(double s) { return (s >= 0) ? 1.0 : -1.0; }   // Foam::sign
__host__ __device__ __forceinline__ double pos0(double s) { return (s >= 0) ? 1.0 : 0.0; }
__host__ __device__ __forceinline__ double neg0(double s) { return (s <= 0) ? 1.0 : 0.0; }
__host__ __device__ __forceinline__ double dmax(double a, double b) { return (a > b) ? a : b; }  // Foam::max
__host__ __device__ __forceinline__ double dmin(double a, double b) { return (a < b) ? a : b; }  // Foam::min

__device__ __forceinline__ d3 ld3(const double* __restrict__ p, int64_t i)
{
    const double* q = p + 3 * i;
    return mk3(__ldg(q), __ldg(q + 1), __ldg(q + 2));
}
__device__ __forceinline__ void st3(double* p, int64_t i, const d3& v)
{
    double* q = p + 3 * i;
    q[0] = v.x;
    q[1] = v.y;
    q[2] = v.z;
}

// order-preserving map double -> uint64 so min/max reduce with integer atomics
// (exact and order independent => deterministic)
__host__ __device__ __forceinline__ unsigned long long dkey(double v)
{
#ifdef __CUDA_ARCH__
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long u;
    memcpy(&u, &v, 8);
#endif
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dunkey(unsigned long long k)
{
    unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double v;
    memcpy(&v, &u, 8);
    return v;
#endif
}

}  // namespace svof
