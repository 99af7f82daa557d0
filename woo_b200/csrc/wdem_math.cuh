// wdem_math.cuh — float64 3-vector / quaternion helpers for the sm_100a DEM kernels.
// Operation order is part of the contract: the translation unit is compiled with --fmad=false and the order of
// every 3-term reduction is fixed by the three switches below (the same names and defaults are used by
// oracle/dem_oracle.cpp; DESIGN.md section 2 "Eigen semantics" has the analysis).  IEEE div/sqrt (fp64 default).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// The reference evaluates its 3-vector algebra with Eigen; Eigen is not available here, so the order in which Eigen adds the three
// products of dot()/squaredNorm(), of Matrix3*Vector3 / Matrix3*Matrix3 and of Matrix3::trace() is restated from its sources.
//   WD_SUM3_ORDER    dot / squaredNorm / norm of 3-vectors, Transpose<Matrix3>*Vector3:
//                    0 = (c0+c1)+c2  [default: Eigen >= 3.3 on x86-64/SSE2, redux LinearVectorizedTraversal: one Packet2d + scalar tail]
//                    1 = c0+(c1+c2)  [Eigen without vectorisation / Eigen 3.2: redux_novec_unroller splits by halves]
//   WD_TRACE_ORDER   Matrix3::trace() inside Quaternion(Matrix3) (strided diagonal: never vectorised):
//                    1 = c0+(c1+c2)  [default: redux_novec_unroller by halves]     0 = (c0+c1)+c2
//   WD_MATPROD_ORDER rows of Matrix3*Vector3 and Matrix3*Matrix3 (coefficient-based lazy product):
//                    1 = rows 0,1 (c0+c1)+c2 (Packet2d pmul/pmadd chain), row 2 c0+(c1+c2) (scalar tail through a strided row: novec redux)  [default]
//                    0 = every row (c0+c1)+c2        2 = every row c0+(c1+c2) [no vectorisation]
// tests/test_order_sensitivity.py runs the parity scenes under the flipped orders (libwdem_alt.so / libdem_oracle_alt.so) and shows
// that the real-contact set is unchanged and forces move by < 1e-10 relative, i.e. that parity does not hinge on this choice.
#ifndef WD_SUM3_ORDER
#define WD_SUM3_ORDER 0
#endif
#ifndef WD_TRACE_ORDER
#define WD_TRACE_ORDER 1
#endif
#ifndef WD_MATPROD_ORDER
#define WD_MATPROD_ORDER 1
#endif

namespace wd {

__host__ __device__ __forceinline__ double sum3(double a, double b, double c) {
#if WD_SUM3_ORDER == 0
  return (a + b) + c;
#else
  return a + (b + c);
#endif
}
__host__ __device__ __forceinline__ double trace3(double a, double b, double c) {
#if WD_TRACE_ORDER == 0
  return (a + b) + c;
#else
  return a + (b + c);
#endif
}
// row `ROW` of a coefficient-based 3x3 product
template <int ROW>
__host__ __device__ __forceinline__ double rowsum3(double a, double b, double c) {
#if WD_MATPROD_ORDER == 0
  return (a + b) + c;
#elif WD_MATPROD_ORDER == 1
  return ROW < 2 ? (a + b) + c : a + (b + c);
#else
  return a + (b + c);
#endif
}

struct d3 {
  double x, y, z;
};
__host__ __device__ __forceinline__ d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ d3 operator-(d3 a) { return mk3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ d3 operator*(d3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ d3 operator*(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ d3 operator/(d3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ double dot(d3 a, d3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ double sqn(d3 a) { return dot(a, a); }
__host__ __device__ __forceinline__ double nrm(d3 a) { return sqrt(sqn(a)); }
__host__ __device__ __forceinline__ d3 cross(d3 a, d3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ double comp(d3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__host__ __device__ __forceinline__ void setc(d3& a, int i, double v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }
__host__ __device__ __forceinline__ d3 xyz(double4 v) { return mk3(v.x, v.y, v.z); }

struct q4 { double w, x, y, z; };

// Eigen Quaternion(Matrix3) from columns c0,c1,c2 (L6Geom.cpp:68,215: node.ori=Quaternionr(trsf))
__host__ __device__ __forceinline__ q4 quatFromCols(d3 c0, d3 c1, d3 c2) {
  // m(i,j) = cj[i]
  const double m00 = c0.x, m10 = c0.y, m20 = c0.z, m01 = c1.x, m11 = c1.y, m21 = c1.z, m02 = c2.x, m12 = c2.y, m22 = c2.z;
  q4 q;
  double t = trace3(m00, m11, m22);
  if (t > 0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m21 - m12) * t;
    q.y = (m02 - m20) * t;
    q.z = (m10 - m01) * t;
  } else {
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > (i == 0 ? m00 : m11)) i = 2;
    if (i == 0) {  // j=1,k=2
      t = sqrt(m00 - m11 - m22 + 1.0);
      q.x = 0.5 * t; t = 0.5 / t;
      q.w = (m21 - m12) * t; q.y = (m10 + m01) * t; q.z = (m20 + m02) * t;
    } else if (i == 1) {  // j=2,k=0
      t = sqrt(m11 - m22 - m00 + 1.0);
      q.y = 0.5 * t; t = 0.5 / t;
      q.w = (m02 - m20) * t; q.z = (m21 + m12) * t; q.x = (m01 + m10) * t;
    } else {  // i=2: j=0,k=1
      t = sqrt(m22 - m00 - m11 + 1.0);
      q.z = 0.5 * t; t = 0.5 / t;
      q.w = (m10 - m01) * t; q.x = (m02 + m20) * t; q.y = (m12 + m21) * t;
    }
  }
  return q;
}
// Eigen QuaternionBase::_transformVector (Contact.cpp:30-31, DynDt.cpp:29)
__host__ __device__ __forceinline__ d3 quatRotate(q4 q, d3 v) {
  d3 qv = mk3(q.x, q.y, q.z);
  d3 uv = cross(qv, v);
  uv = uv + uv;
  return (v + q.w * uv) + cross(qv, uv);
}

// Eigen quaternion product a*b
__host__ __device__ __forceinline__ q4 quatMul(q4 a, q4 b) {
  q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z; r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z; r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
__host__ __device__ __forceinline__ q4 quatConj(q4 q) { q4 r; r.w = q.w; r.x = -q.x; r.y = -q.y; r.z = -q.z; return r; }
__host__ __device__ __forceinline__ q4 quatNormalized(q4 q) {
  const double n = sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
  q4 r; r.w = q.w / n; r.x = q.x / n; r.y = q.y / n; r.z = q.z / n; return r;
}

struct m33 { double m[9]; };  // row-major, host-prepared uniforms
// Eigen QuaternionBase::toRotationMatrix
__host__ __device__ __forceinline__ m33 quatToMat(q4 q) {
  m33 r;
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0] = 1 - (tyy + tzz); r.m[1] = txy - twz; r.m[2] = txz + twy;
  r.m[3] = txy + twz; r.m[4] = 1 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1 - (txx + tyy);
  return r;
}
__host__ __device__ __forceinline__ d3 mulMV(const m33& M, d3 v) {
  return mk3(rowsum3<0>(M.m[0] * v.x, M.m[1] * v.y, M.m[2] * v.z), rowsum3<1>(M.m[3] * v.x, M.m[4] * v.y, M.m[5] * v.z),
             rowsum3<2>(M.m[6] * v.x, M.m[7] * v.y, M.m[8] * v.z));
}

}  // namespace wd
