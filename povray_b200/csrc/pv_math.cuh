// FP64 vector / transform helpers with the reference's operation order
// (source/core/math/vector.h:395-590, matrix.cpp:415-507).  The TU is compiled with -fmad=false, so
// a*b+c below is a rounded multiply followed by a rounded add, like the -ffp-contract=off oracle.
#pragma once
#include "pv_common.cuh"

namespace pvgpu {

__device__ __forceinline__ V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(const V3& a, const V3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(const V3& a, const V3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(const V3& a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(const V3& a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(double s, const V3& a) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator/(const V3& a, double s) { return mk(a.x / s, a.y / s, a.z / s); }
// dot(): ((ax*bx) + (ay*by)) + (az*bz)            vector.h:568
__device__ __forceinline__ double dot(const V3& a, const V3& b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
__device__ __forceinline__ double length_sqr(const V3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ double length(const V3& a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
// normalize(): divide by the length, component-wise (no reciprocal)        vector.h:531
__device__ __forceinline__ V3 normalized(const V3& a)
{
    double l = length(a);
    return (l != 0.0) ? mk(a.x / l, a.y / l, a.z / l) : a;
}
__device__ __forceinline__ V3 cross(const V3& a, const V3& b) { return mk((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)); }   // vector.h:575
__device__ __forceinline__ double comp(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__device__ __forceinline__ V3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ V3 ld3f(const float* p) { return mk((double)p[0], (double)p[1], (double)p[2]); }
// BasicRay::Evaluate: Origin + Direction * depth                           coretypes.h:470
__device__ __forceinline__ V3 evaluate(const V3& o, const V3& d, double t) { return mk(o.x + d.x * t, o.y + d.y * t, o.z + d.z * t); }

__device__ __forceinline__ double sqr(double x) { return x * x; }

// MTransPoint / MTransDirection with matrix m (row-major [4][4])           matrix.cpp:415,455
__device__ __forceinline__ V3 m_point(const double* m, const V3& v)
{
    return mk(v.x * m[0] + v.y * m[4] + v.z * m[8] + m[12],
              v.x * m[1] + v.y * m[5] + v.z * m[9] + m[13],
              v.x * m[2] + v.y * m[6] + v.z * m[10] + m[14]);
}
__device__ __forceinline__ V3 m_direction(const double* m, const V3& v)
{
    return mk(v.x * m[0] + v.y * m[4] + v.z * m[8],
              v.x * m[1] + v.y * m[5] + v.z * m[9],
              v.x * m[2] + v.y * m[6] + v.z * m[10]);
}
// MInvTransNormal(result, v, matrix): uses the transposed 3x3              matrix.cpp:495
__device__ __forceinline__ V3 m_transposed(const double* m, const V3& v)
{
    return mk(v.x * m[0] + v.y * m[1] + v.z * m[2],
              v.x * m[4] + v.y * m[5] + v.z * m[6],
              v.x * m[8] + v.y * m[9] + v.z * m[10]);
}
// TRANSFORM wrappers (matrix.h:97-102)
__device__ __forceinline__ V3 trans_point(const pvgpu_transform& t, const V3& v) { return m_point(t.matrix, v); }
__device__ __forceinline__ V3 inv_trans_point(const pvgpu_transform& t, const V3& v) { return m_point(t.inverse, v); }
__device__ __forceinline__ V3 trans_direction(const pvgpu_transform& t, const V3& v) { return m_direction(t.matrix, v); }
__device__ __forceinline__ V3 inv_trans_direction(const pvgpu_transform& t, const V3& v) { return m_direction(t.inverse, v); }
__device__ __forceinline__ V3 trans_normal(const pvgpu_transform& t, const V3& v) { return m_transposed(t.inverse, v); }
__device__ __forceinline__ V3 inv_trans_normal(const pvgpu_transform& t, const V3& v) { return m_transposed(t.matrix, v); }

// FLOOR (texture.h:73)
__device__ __forceinline__ double pv_floor(double x) { return (x >= 0.0) ? floor(x) : (0.0 - floor(0.0 - x) - 1.0); }

struct Ray3 { V3 o, d; };

}  // namespace pvgpu
