// dem_oracle.cpp — CPU restatement ("oracle") of woodem/woo's per-timestep DEM hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product (woo_b200/) never does.
//
// PARITY STATUS: "parity unpinned" at the bit level.  The reference cannot be compiled here
// (needs Eigen3/Boost/TBB; none installed, no network) and ships no golden data for this path.
// The oracle is pinned against the reference's own known-answer tests (tests/test_oracle_kat.py:
// Hertz collision time test_hertz.py:30-47, pWaveDt utils.py:80-81, Cundall kn models.py:16-19,
// PBC cellDist test_pbc.py:130-171, Wall+sphere Leapfrog and EnergyTracker entries test_energy.py:69-90), against golden vectors produced by the reference's own
// Python contact models (py/models.py CundallModel / HertzModel, imported from /root/reference by scripts/make_golden_models.py -> tests/golden/ref_models.json)
// and otherwise by review against the cited lines.  tests/golden/scene_*.npz pin the oracle's own output against drift between rounds.  The order of Eigen's fixed-size 3-term reductions is restated from
// the Eigen sources and switchable (WD_SUM3_ORDER / WD_TRACE_ORDER / WD_MATPROD_ORDER below); tests/test_order_sensitivity.py shows that the
// real-contact set and the 1e-10 force tolerance do not depend on it.  scripts/replay_in_woo.py replays tests/golden/ in a real Woo build.
//
// Every function cites the reference file:line (relative to /root/reference) it follows.
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#include <map>
#include <unordered_map>
#include <string>
#include <algorithm>
#include <limits>
#include <stdexcept>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {
typedef double Real;
static const Real Inf = std::numeric_limits<Real>::infinity();
static const Real NaN = std::numeric_limits<Real>::quiet_NaN();

// ---------------------------------------------------------------------------------------------
// Minimal fixed-size linear algebra with the operation order of the Eigen expressions used on
// the path.  Reductions: left-to-right.
// ---------------------------------------------------------------------------------------------
struct V3 {
  Real x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(Real a, Real b, Real c) : x(a), y(b), z(c) {}
  Real& operator[](int i) { return (&x)[i]; }
  const Real& operator[](int i) const { return (&x)[i]; }
};
static inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }
static inline V3 operator*(const V3& a, Real s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(Real s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline V3 operator/(const V3& a, Real s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline V3& operator+=(V3& a, const V3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
static inline V3& operator-=(V3& a, const V3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
static inline bool operator==(const V3& a, const V3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
static inline bool operator!=(const V3& a, const V3& b) { return !(a == b); }
// Order of Eigen's 3-term reductions: the same three switches (names, meaning, defaults) as woo_b200/csrc/wdem_math.cuh; analysis in
// DESIGN.md section 2.  oracle/Makefile also builds libdem_oracle_alt.so with every order flipped (tests/test_order_sensitivity.py).
#ifndef WD_SUM3_ORDER
#define WD_SUM3_ORDER 0      // dot/squaredNorm/norm, Transpose<Matrix3>*Vector3: 0 = (c0+c1)+c2 (Eigen>=3.3 SSE2 Packet2d + tail), 1 = c0+(c1+c2) (novec halves)
#endif
#ifndef WD_TRACE_ORDER
#define WD_TRACE_ORDER 1     // Matrix3::trace() in Quaternion(Matrix3): 1 = c0+(c1+c2) (strided diagonal: novec halves), 0 = (c0+c1)+c2
#endif
#ifndef WD_MATPROD_ORDER
#define WD_MATPROD_ORDER 1   // Matrix3*Vector3, Matrix3*Matrix3: 1 = rows 0,1 (c0+c1)+c2 + row 2 c0+(c1+c2); 0 = all (c0+c1)+c2; 2 = all c0+(c1+c2)
#endif
static inline Real sum3(Real a, Real b, Real c) {
#if WD_SUM3_ORDER == 0
  return (a + b) + c;
#else
  return a + (b + c);
#endif
}
static inline Real trace3(Real a, Real b, Real c) {
#if WD_TRACE_ORDER == 0
  return (a + b) + c;
#else
  return a + (b + c);
#endif
}
static inline Real rowsum3(int row, Real a, Real b, Real c) {
#if WD_MATPROD_ORDER == 0
  (void)row; return (a + b) + c;
#elif WD_MATPROD_ORDER == 1
  return row < 2 ? (a + b) + c : a + (b + c);
#else
  (void)row; return a + (b + c);
#endif
}
static inline Real dot(const V3& a, const V3& b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline Real sqNorm(const V3& a) { return dot(a, a); }
static inline Real norm(const V3& a) { return std::sqrt(sqNorm(a)); }
static inline V3 cross(const V3& a, const V3& b) {
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline V3 normalized(const V3& a) { return a / norm(a); }
static inline Real pow2(Real x) { return x * x; }
static inline Real pow3(Real x) { return x * x * x; }

struct M3 {  // column-major: c[j] is column j
  V3 c[3];
  Real operator()(int i, int j) const { return c[j][i]; }
  Real& operator()(int i, int j) { return c[j][i]; }
  static M3 Zero() { return M3(); }
  static M3 Identity() { M3 m; m.c[0].x = m.c[1].y = m.c[2].z = 1; return m; }
};
static inline V3 mulMV(const M3& m, const V3& v) {  // m*v, row-wise dot
  return V3(rowsum3(0, m(0, 0) * v.x, m(0, 1) * v.y, m(0, 2) * v.z), rowsum3(1, m(1, 0) * v.x, m(1, 1) * v.y, m(1, 2) * v.z),
            rowsum3(2, m(2, 0) * v.x, m(2, 1) * v.y, m(2, 2) * v.z));
}
static inline V3 mulMtV(const M3& m, const V3& v) {  // m^T*v = (col_j . v)
  return V3(dot(m.c[0], v), dot(m.c[1], v), dot(m.c[2], v));
}
static inline M3 mulMM(const M3& a, const M3& b) {
  M3 r;
  for (int j = 0; j < 3; j++) r.c[j] = mulMV(a, b.c[j]);
  return r;
}
static inline M3 addMM(const M3& a, const M3& b) { M3 r; for (int j = 0; j < 3; j++) r.c[j] = a.c[j] + b.c[j]; return r; }
static inline M3 subMM(const M3& a, const M3& b) { M3 r; for (int j = 0; j < 3; j++) r.c[j] = a.c[j] - b.c[j]; return r; }
static inline M3 scaleM(const M3& a, Real s) { M3 r; for (int j = 0; j < 3; j++) r.c[j] = a.c[j] * s; return r; }
static inline M3 transposeM(const M3& a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = a(j, i); return r; }
static inline Real detM(const M3& m) {
  return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
         m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
}
static inline M3 inverseM(const M3& m) {  // cofactor formula (host-only use: Cell, Leapfrog GRADV2 matrices)
  M3 r;
  Real d = detM(m);
  Real id = 1. / d;
  r(0, 0) = (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) * id;
  r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * id;
  r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * id;
  r(1, 0) = (m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2)) * id;
  r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * id;
  r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) * id;
  r(2, 0) = (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0)) * id;
  r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) * id;
  r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) * id;
  return r;
}
// src/supp/base/Math.hpp:223-228
static inline V3 leviCivita(const M3& m) { return V3(m(1, 2) - m(2, 1), m(2, 0) - m(0, 2), m(0, 1) - m(1, 0)); }

struct Quat {
  Real w, x, y, z;
  Quat() : w(1), x(0), y(0), z(0) {}
  Quat(Real w_, Real x_, Real y_, Real z_) : w(w_), x(x_), y(y_), z(z_) {}
};
// Eigen Quaternion(Matrix3) — quaternionbase_assign_impl<Other,3,3>; used at L6Geom.cpp:68,215
static inline Quat quatFromMat(const M3& m) {
  Quat q;
  Real t = trace3(m(0, 0), m(1, 1), m(2, 2));
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m(2, 1) - m(1, 2)) * t;
    q.y = (m(0, 2) - m(2, 0)) * t;
    q.z = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    Real qv[3];
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m(k, j) - m(j, k)) * t;
    qv[j] = (m(j, i) + m(i, j)) * t;
    qv[k] = (m(k, i) + m(i, k)) * t;
    q.x = qv[0]; q.y = qv[1]; q.z = qv[2];
  }
  return q;
}
// Eigen QuaternionBase::_transformVector; used at Contact.cpp:30-31, DynDt.cpp:29
static inline V3 quatRotate(const Quat& q, const V3& v) {
  V3 qv(q.x, q.y, q.z);
  V3 uv = cross(qv, v);
  uv += uv;
  return (v + q.w * uv) + cross(qv, uv);
}
static inline Quat quatMul(const Quat& a, const Quat& b) {
  return Quat(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x);
}
static inline void quatNormalize(Quat& q) {
  Real n = std::sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
// Leapfrog::DotQ (Leapfrog.cpp:324-331)
static inline Quat dotQ(const V3& w, const Quat& Q) {
  Quat d;
  d.w = (-Q.x * w.x - Q.y * w.y - Q.z * w.z) / 2;
  d.x = (Q.w * w.x - Q.z * w.y + Q.y * w.z) / 2;
  d.y = (Q.z * w.x + Q.w * w.y - Q.x * w.z) / 2;
  d.z = (-Q.y * w.x + Q.x * w.y + Q.w * w.z) / 2;
  return d;
}
static inline M3 quatToMat(const Quat& q) {  // Eigen toRotationMatrix
  M3 r;
  Real tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  Real twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y,
       tyz = tz * q.y, tzz = tz * q.z;
  r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
  r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Data model.  DemData flags: src/dem/Particle.hpp:198-246
// ---------------------------------------------------------------------------------------------
enum { DOF_X = 1, DOF_Y = 2, DOF_Z = 4, DOF_RX = 8, DOF_RY = 16, DOF_RZ = 32, DOF_ALL = 63, DOF_XYZ = 7, DOF_RXRYRZ = 56,
       CLUMP_CLUMPED = 64, CLUMP_CLUMP = 128, ENERGY_SKIP = 256, GRAVITY_SKIP = 512, TRACER_SKIP = 1024, DAMPING_SKIP = 2048 };
enum { SHAPE_SPHERE = 0, SHAPE_WALL = 1, SHAPE_FACET = 2 };
enum { MAT_FRICT = 0, MAT_HERTZ = 1 };
enum { LAW_IDEALELPL = 0, LAW_HERTZ_DMT = 1 };
enum { HOMO_NONE = 0, HOMO_POS = 1, HOMO_VEL = 2, HOMO_VEL_2ND = 3, HOMO_GRADV2 = 4 };  // src/core/Cell.hpp

struct Material { int kind; Real density, young, ktDivKn, tanPhi, surfEnergy, alpha; };

struct Aabb { V3 mn, mx; bool has = false; };

struct Sphere {
  int id; V3 pos; Quat ori; V3 vel, angVel, force, torque; Real mass, inertia, radius; unsigned flags, mask; int mat;
  Aabb bb; V3 lastPos;
  bool dead = false;   // removed by ParticleContainer::remove (the slot stays: ids are not renumbered)
  int clump = -1;      // index of the clump this sphere is a member of (DemData::isClumped), -1: free
  int impose = -1;     // index of the set of impositions on the node (DemData.impose), -1: none
};
struct Wall {
  int id; V3 pos; int axis, sense; V3 vel, angVel; unsigned mask; int mat; V3 force, torque; Aabb bb; V3 lastPos; int impose = -1;
  bool moving = false;   // Particle::guessMoving at ParticleContainer::add (ParticleContainer.cpp:193-198, Particle.cpp:638-640): a node with nonzero velocity
                         // enters DemField.nodes and is integrated by Leapfrog; a fixed wall at rest never is
};
struct Facet {
  int id; V3 v[3]; V3 vvel[3]; V3 fakeVel; Real halfThick; Real n21lim[3]; unsigned mask; int mat; Aabb bb; V3 lastPos[3];
  int impose = -1;     // set of impositions carried by all three nodes (py/triangulated.py:59-64), -1: none
};
// Master node of a clump of boundary particles that moves them rigidly (ParticleContainer::addClumped -> ClumpData::makeClump, Clump.cpp:187-261): a massless,
// fully blocked node with prescribed vel/angVel (examples/mill.py:38-42) whose members are facets.  relPos are the members' node positions in the clump's frame.
struct FacetGroup { V3 pos; Quat ori; V3 vel, angVel; std::vector<int> facets; std::vector<V3> relPos; int impose = -1; };
// Impose (Particle.hpp:130-160, Impose.hpp, Impose.cpp): hooks called by Leapfrog on a node — velocity after the accelerations are applied and again at the end of
// the node's step, force at the reset, readForce before the reset.  One descriptor per Impose object; a node refers to a SET of them (CombinedImpose, Impose.cpp:283-300).
enum { IMP_HARMONIC = 1, IMP_ALIGNED_HARMONIC = 2, IMP_CONSTANT_FORCE = 3, IMP_RADIAL_FORCE = 4, IMP_LOCAL6 = 5, IMP_VEL_READ_FORCE = 6, IMP_READ_FORCE = 7, IMP_VARIABLE_VEL3D = 8,
       IMP_VARIABLE_ALIGNED_ROTATION = 9, IMP_CIRCULAR_ORBIT = 10, IMP_STABLE_CIRCULAR_ORBIT = 11, IMP_INTERPOLATED_MOTION = 12 };
enum { IMPW_VELOCITY = 1, IMPW_FORCE = 2, IMPW_READ_FORCE = 8 };
struct ImposeDesc {
  int kind = 0, flags = 0; double v[16] = {0}; Quat ori; std::vector<double> times, vals, vals2;
  size_t ipos = 0, ipos2 = 0;                 // linearInterpolate's lookup cookies
  double sumF = 0, dist = 0; V3 F, T;         // VelocityAndReadForce.sumF/.dist, ReadForce.F/.T
  int nNodes = 0;                             // nodes carrying it
  int what() const {
    switch (kind) {
      case IMP_HARMONIC: case IMP_ALIGNED_HARMONIC: case IMP_VARIABLE_VEL3D: case IMP_VARIABLE_ALIGNED_ROTATION: case IMP_CIRCULAR_ORBIT: case IMP_STABLE_CIRCULAR_ORBIT: case IMP_INTERPOLATED_MOTION: return IMPW_VELOCITY;
      case IMP_CONSTANT_FORCE: case IMP_RADIAL_FORCE: return IMPW_FORCE;
      case IMP_LOCAL6: return IMPW_VELOCITY | IMPW_FORCE;
      case IMP_VEL_READ_FORCE: return IMPW_VELOCITY | IMPW_READ_FORCE;
      case IMP_READ_FORCE: return IMPW_READ_FORCE;
    }
    return 0;
  }
};
// linearInterpolate (supp/smoothing/LinearInterpolate.hpp:13-52) for values of `w` components per time point
static inline void linInterp(Real t, const std::vector<double>& tt, const std::vector<double>& val, int w, size_t& _pos, double* out) {
  size_t pos = _pos; const size_t n = tt.size();
  const double *A, *B; Real rel;
  if (t <= tt[0]) { _pos = 0; A = &val[0]; B = &val[(n > 1 ? 1 : 0) * w]; rel = 0.; }
  else if (t >= tt[n - 1]) { _pos = n - 2; A = n > 1 ? &val[pos * w] : &val[(n - 1) * w]; B = &val[(n - 1) * w]; rel = 1.; }
  else {
    pos = std::min(pos, n - 2);
    while ((tt[pos] > t) || (tt[pos + 1] < t)) { if (tt[pos] > t) pos--; else pos++; }
    _pos = pos; A = &val[pos * w]; B = &val[(pos + 1) * w]; rel = (t - tt[pos]) / (tt[pos + 1] - tt[pos]);
  }
  for (int k = 0; k < w; k++) out[k] = (rel == 0.) ? A[k] : (rel == 1.) ? B[k] : A[k] + (B[k] - A[k]) * rel;
}
// Clump of spheres with mass (ClumpData, Clump.hpp:38-60): the master node carries mass / principal inertia and is integrated by Leapfrog (aspherical rotation when
// the inertia components differ); member nodes are slaves (DemData::isClumped): positions, orientations and velocities follow, forces are collected from them.
struct Clump {
  V3 pos; Quat ori; V3 vel, angVel, angMom = V3(NaN, NaN, NaN); Real mass = 0; V3 inertia; V3 force, torque; unsigned flags = 0;
  std::vector<int> members /*sphere indices*/; std::vector<V3> relPos; std::vector<Quat> relOri;
  int impose = -1;     // set of impositions on the master node
};
struct PRef { int kind, idx; };

struct Contact {
  int idA, idB;        // canonical order: boundary (wall/facet) first; sphere-sphere: lower id first
  int cellDist[3];
  bool hasGeom = false, hasPhys = false;
  int stepCreated = -1;
  bool overlapBB = true;  // AABBs overlap as of the last collider full run
  // L6Geom (src/dem/L6Geom.hpp:16-21) + CGeom::node (Contact.hpp:21)
  M3 trsf; V3 contPt; Quat ori; V3 gvel, gangVel; Real uN = NaN; Real lens[2] = {0, 0}; Real contA = NaN;
  // FrictPhys (Contact.hpp:32-33, FrictMat.hpp:33-35) + HertzPhys (Hertz.hpp:20-26)
  V3 force, torque; Real kn = NaN, kt = NaN, tanPhi = NaN;
  Real kt0 = NaN, alpha_sqrtMK = NaN, R = NaN, K = NaN, gamma = NaN, alpha = NaN;
  bool isReal() const { return hasGeom && hasPhys; }
  void reset() { hasGeom = hasPhys = false; }  // Contact.cpp:18-22 (stepCreated kept)
};

struct Params {  // mirrors the engine attributes of SURVEY §8b
  // Scene / DemField
  Real dt = NaN, dtSafety = 0.9; Real gravity[3] = {0, 0, 0}; unsigned loneMask = 2; Real distFactor = -1; int deterministic = 1;
  // Leapfrog (Leapfrog.hpp:50-53)
  Real damping = 0.2; int reset = 1;
  // AabbCollider (Collision.hpp:74,94)
  Real verletDist = -0.05; Real wallHalfThick = 0;
  // Cg2 (L6Geom.hpp:35-45)
  int noRatch = 0, iniLensTouch = 1, trsfRenorm = 100;
  // ContactLoop
  int applyForces = 1;
  // law selection + Law2 IdealElPl (IdealElPl.hpp:17-24)
  int law = LAW_IDEALELPL; int noSlip = 0, noBreak = 0, noFrict = 0;
  Real relRollStiff = 0, relTwistStiff = 0, rollTanPhi = 0;
  // Cp2_HertzMat_HertzPhys (Hertz.hpp:58-59), Law2 DMT (Hertz.hpp:96)
  Real poisson = 0.2, en = NaN; int noAttraction = 1;
  // DynDt (DynDt.hpp:21-23)
  int dynDtPeriod = 100; Real maxRelInc = 1e-4;
  // periodic cell
  int periodic = 0;
};

struct Cell {  // src/core/Cell.{hpp,cpp}
  M3 hSize = M3::Identity(), trsf = M3::Identity(), gradV, nextGradV, pprevHsize = M3::Identity(), W;
  V3 spinVec, size, cosv; M3 shearTrsf = M3::Identity(), unshearTrsf = M3::Identity(); bool hasShear = false;
  int homoDeform = HOMO_GRADV2;
  // Cell.cpp:81-120
  void integrateAndUpdate(Real dt) {
    M3 trsfInc = scaleM(gradV, dt);
    trsf = addMM(trsf, mulMM(trsfInc, trsf));
    M3 prevHsize = hSize;
    if (homoDeform == HOMO_GRADV2) {
      M3 A = inverseM(subMM(M3::Identity(), scaleM(scaleM(gradV, dt), 1 / 2.)));
      M3 B = addMM(M3::Identity(), scaleM(scaleM(gradV, dt), 1 / 2.));
      hSize = mulMM(mulMM(A, B), hSize);
    } else hSize = addMM(hSize, mulMM(trsfInc, hSize));
    pprevHsize = scaleM(addMM(prevHsize, hSize), .5);
    if (detM(hSize) == 0) throw std::runtime_error("Cell is degenerate (zero volume).");
    M3 Hnorm;
    for (int i = 0; i < 3; i++) {
      V3 base = hSize.c[i];
      size[i] = norm(base);
      base = base / size[i];
      Hnorm.c[i] = base;
    }
    for (int i = 0; i < 3; i++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
      cosv[i] = sqNorm(cross(Hnorm.c[i1], Hnorm.c[i2]));
    }
    shearTrsf = Hnorm;
    unshearTrsf = inverseM(shearTrsf);
    hasShear = (hSize(0, 1) != 0 || hSize(0, 2) != 0 || hSize(1, 0) != 0 || hSize(1, 2) != 0 || hSize(2, 0) != 0 || hSize(2, 1) != 0);
  }
  // Cell.cpp:60-66
  void setNextGradV() { gradV = nextGradV; W = scaleM(subMM(gradV, transposeM(gradV)), .5); spinVec = .5 * leviCivita(W); }
  V3 unshearPt(const V3& p) const { return mulMV(unshearTrsf, p); }
  // Cell.hpp:93
  V3 intrShiftPos(const int* cd) const { return mulMV(hSize, V3((Real)cd[0], (Real)cd[1], (Real)cd[2])); }
  // Cell.cpp:28-34
  V3 intrShiftVel(const int* cd) const {
    V3 c((Real)cd[0], (Real)cd[1], (Real)cd[2]);
    switch (homoDeform) {
      case HOMO_VEL: case HOMO_VEL_2ND: return mulMV(mulMM(gradV, hSize), c);
      case HOMO_GRADV2: return mulMV(mulMM(gradV, pprevHsize), c);
      default: return V3();
    }
  }
  // Cell.cpp:36-50
  V3 pprevFluctVel(const V3& currPos, const V3& pprevVel, Real dt) const {
    switch (homoDeform) {
      case HOMO_NONE: case HOMO_POS: return pprevVel;
      case HOMO_VEL: case HOMO_VEL_2ND: return pprevVel - mulMV(gradV, currPos);
      case HOMO_GRADV2: return pprevVel - mulMV(gradV, currPos - dt / 2 * pprevVel);
      default: throw std::runtime_error("invalid homoDeform");
    }
  }
  // Cell.cpp:52-57
  V3 pprevFluctAngVel(const V3& w) const { return homoDeform == HOMO_GRADV2 ? w - spinVec : w; }
  // Cell.cpp:68-79
  V3 shearAlignedExtents(const V3& perp) const {
    if (!hasShear) return perp;
    V3 ret = perp;
    for (int ax = 0; ax < 3; ax++) {
      int ax1 = (ax + 1) % 3, ax2 = (ax + 2) % 3;
      ret[ax1] += .5 * perp[ax1] * (1 / cosv[ax] - 1);
      ret[ax2] += .5 * perp[ax2] * (1 / cosv[ax] - 1);
    }
    return ret;
  }
  Real volume() const { return detM(hSize); }
};

static inline uint64_t pairKey(int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; }

// InsertionSortCollider::Bounds (InsertionSortCollider.hpp:121-141)
struct Bound {
  Real coord; int id; bool isMin, hasBB;
  bool gt(const Bound& b) const { if (id == b.id && coord == b.coord) return !isMin; return coord > b.coord; }
  bool lt(const Bound& b) const { if (id == b.id && coord == b.coord) return isMin; return coord < b.coord; }
};

struct Scene {
  // Scene.trackEnergy / EnergyTracker (core/EnergyTracker.hpp:30-60): named accumulators; "kinetic" and "elast" are resettable (zeroed in the step prologue, Scene.cpp:416)
  enum { EN_KINETIC = 0, EN_GRAV, EN_DAMP, EN_ELAST, EN_PLAST, EN_BROKEN, EN_VISCN, EN_VISCT, EN_COUNT };
  bool trackEnergy = false; double energy[EN_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0};
  // ContactLoop.evalStress + the "gradV" energy (ContactLoop.cpp:116,218-231,243-254)
  bool evalStress = false; M3 prevStress; Real prevVol = NaN; double gradVWork = 0;
  void eadd(int ix, Real v) {  // EnergyTracker::add (EnergyTracker.cpp:26-34): NaN is ignored; callable from parallel sections
    if (std::isnan(v)) return;
    #pragma omp atomic
    energy[ix] += v;
  }
  Params P; Cell cell;
  std::vector<Material> mats; std::vector<Sphere> sph; std::vector<Wall> walls; std::vector<Facet> facets; std::vector<FacetGroup> groups; std::vector<Clump> clumps; std::vector<PRef> par;
  std::map<uint64_t, Contact> con;          // ordered by canonical key
  std::vector<Contact*> lin; bool linDirty = true;
  std::vector<uint64_t> pending;            // ContactContainer::requestRemoval queue
  long step = 0; Real time = 0; Real nextDt = NaN;
  // collider state
  std::vector<Bound> BB[3]; std::vector<Real> minima, maxima; int nFullRuns = 0; bool conDirty = false; bool verletSet = false;
  std::vector<ImposeDesc> imposes; std::vector<std::vector<int>> imposeSets;
  std::vector<int> freeIds; std::vector<int> lastAdded;   // ParticleContainer::freeIds (ParticleContainer.cpp:23-36,125); ids handed out by the last orc_add_spheres
  long dynDtStepLast = -1, dynDtStepPrev = -1;
  std::string err;

  // ------------------------------------------------------------------------------------------
  // helpers
  // ------------------------------------------------------------------------------------------
  unsigned maskOf(int id) const { const PRef& p = par[id]; return p.kind == SHAPE_SPHERE ? sph[p.idx].mask : p.kind == SHAPE_WALL ? walls[p.idx].mask : facets[p.idx].mask; }
  int matOf(int id) const { const PRef& p = par[id]; return p.kind == SHAPE_SPHERE ? sph[p.idx].mat : p.kind == SHAPE_WALL ? walls[p.idx].mat : facets[p.idx].mat; }
  Aabb& bbOf(int id) { PRef& p = par[id]; return p.kind == SHAPE_SPHERE ? sph[p.idx].bb : p.kind == SHAPE_WALL ? walls[p.idx].bb : facets[p.idx].bb; }

  // Collider::mayCollide (Collision.cpp:21-43); facets sharing a node cannot happen here (facet-facet never passes masks)
  bool mayCollide(int a, int b) const {
    if (a == b) return false;
    unsigned mA = maskOf(a), mB = maskOf(b);
    if (!(mA & mB)) return false;
    if ((mA & mB & P.loneMask) != 0) return false;
    // particles inside the same clump (Collision.cpp:31-40)
    if (par[a].kind == SHAPE_SPHERE && par[b].kind == SHAPE_SPHERE) { const int ca = sph[par[a].idx].clump, cb = sph[par[b].idx].clump; if (ca >= 0 && ca == cb) return false; }
    // only sphere-sphere, wall-sphere, facet-sphere have functors on this path (ContactLoop.cpp:149 "if(!cgf) continue")
    return true;
  }
  bool hasFunctor(int a, int b) const {
    int ka = par[a].kind, kb = par[b].kind;
    return (ka == SHAPE_SPHERE || kb == SHAPE_SPHERE);
  }
  // canonical (pA,pB): shapes ordered as the functor declares (Wall.hpp:55-56, Facet.hpp:86-87); sphere-sphere by id
  void canonical(int& a, int& b) const {
    int ka = par[a].kind, kb = par[b].kind;
    if (ka == SHAPE_SPHERE && kb == SHAPE_SPHERE) { if (a > b) std::swap(a, b); return; }
    if (ka == SHAPE_SPHERE) std::swap(a, b);
  }
  Contact* findContact(int a, int b) {
    canonical(a, b);
    auto it = con.find(pairKey(a, b));
    return it == con.end() ? nullptr : &it->second;
  }

  // ------------------------------------------------------------------------------------------
  // Bo1_*_Aabb + AabbCollider::updateAabb  (Sphere.cpp:52-82, Wall.cpp:25-38, Facet.cpp:275-298, Collision.cpp:125-162)
  // ------------------------------------------------------------------------------------------
  void updateAabbAll() {
    const Real vd = P.verletDist;
    for (auto& s : sph) {
      V3 half = ((P.distFactor > 0 ? P.distFactor : 1.) * s.radius) * V3(1, 1, 1);
      if (!P.periodic) { s.bb.mn = s.pos - half; s.bb.mx = s.pos + half; }
      else { half = cell.shearAlignedExtents(half); V3 u = cell.unshearPt(s.pos); s.bb.mn = u - half; s.bb.mx = u + half; }
      s.bb.has = true; s.lastPos = s.pos;
      if (vd > 0) { s.bb.mx += vd * V3(1, 1, 1); s.bb.mn -= vd * V3(1, 1, 1); }
    }
    for (auto& w : walls) {
      if (P.periodic && cell.hasShear) throw std::runtime_error("Walls not supported in skewed periodic boundary conditions.");
      w.bb.mn = V3(-Inf, -Inf, -Inf); w.bb.mx = V3(Inf, Inf, Inf);
      w.bb.mn[w.axis] = w.bb.mx[w.axis] = w.pos[w.axis];
      if (P.wallHalfThick > 0) { w.bb.mn[w.axis] -= P.wallHalfThick; w.bb.mx[w.axis] += P.wallHalfThick; }
      w.bb.has = true; w.lastPos = w.pos;
      if (vd > 0) { w.bb.mx += vd * V3(1, 1, 1); w.bb.mn -= vd * V3(1, 1, 1); }
    }
    for (auto& f : facets) {
      V3 ht(f.halfThick, f.halfThick, f.halfThick);
      V3 p0 = P.periodic ? cell.unshearPt(f.v[0]) : f.v[0];
      f.bb.mn = p0 - ht; f.bb.mx = p0 + ht;
      for (int i = 1; i < 3; i++) {
        V3 p = P.periodic ? cell.unshearPt(f.v[i]) : f.v[i];
        V3 lo = p - ht, hi = p + ht;
        for (int k = 0; k < 3; k++) { f.bb.mn[k] = std::min(f.bb.mn[k], lo[k]); f.bb.mx[k] = std::max(f.bb.mx[k], hi[k]); }
      }
      f.bb.has = true; for (int i = 0; i < 3; i++) f.lastPos[i] = f.v[i];
      if (vd > 0) { f.bb.mx += vd * V3(1, 1, 1); f.bb.mn -= vd * V3(1, 1, 1); }
    }
  }
  // AabbCollider::setVerletDist (Collision.cpp:164-184); equivRadius: sphere radius, NaN for Wall/Facet
  void setVerletDist() {
    if (P.verletDist >= 0) return;
    Real minR = Inf;
    for (auto& s : sph) minR = std::min(s.radius, minR);
    if (!std::isinf(minR)) P.verletDist = std::abs(P.verletDist) * minR; else P.verletDist = 0.;
  }
  // AabbCollider::aabbIsDirty (Collision.cpp:80-123), maxRot<0 for all shapes on this path
  bool anyAabbDirty() {
    const Real maxD2 = pow2(P.verletDist);
    for (auto& pr : par) {
      if (pr.kind == SHAPE_SPHERE) {
        Sphere& s = sph[pr.idx];
        if (!s.bb.has) return true;
        Real d2 = std::max(0., sqNorm(s.lastPos - s.pos));
        if (d2 + 0. > maxD2) return true;
      } else if (pr.kind == SHAPE_WALL) { if (!walls[pr.idx].bb.has) return true; Real d2 = std::max(0., sqNorm(walls[pr.idx].lastPos - walls[pr.idx].pos)); if (d2 > maxD2) return true; }
      else { Facet& f = facets[pr.idx]; if (!f.bb.has) return true; Real d2 = 0; for (int i = 0; i < 3; i++) d2 = std::max(d2, sqNorm(f.lastPos[i] - f.v[i])); if (d2 > maxD2) return true; }
    }
    return false;
  }

  // ------------------------------------------------------------------------------------------
  // InsertionSortCollider (aperiodic: faithful incremental sweep; periodic: overlap set rebuilt from scratch
  // with spatialOverlapPeri semantics — algorithm free)
  // ------------------------------------------------------------------------------------------
  bool spatialOverlap(int id1, int id2) const {  // InsertionSortCollider.cpp:67-73
    return (minima[3 * id1 + 0] <= maxima[3 * id2 + 0]) && (maxima[3 * id1 + 0] >= minima[3 * id2 + 0]) &&
           (minima[3 * id1 + 1] <= maxima[3 * id2 + 1]) && (maxima[3 * id1 + 1] >= minima[3 * id2 + 1]) &&
           (minima[3 * id1 + 2] <= maxima[3 * id2 + 2]) && (maxima[3 * id1 + 2] >= minima[3 * id2 + 2]);
  }
  static Real cellWrapRel(Real x, Real x0, Real x1, int& period) {  // InsertionSortCollider.cpp:624-628
    Real xNorm = (x - x0) / (x1 - x0);
    period = (int)std::floor(xNorm);
    return (xNorm - std::floor(xNorm)) * (x1 - x0);
  }
  // InsertionSortCollider.cpp:897-1026 ("new algorithm")
  bool spatialOverlapPeri_axis(int axis, int id1, int id2, Real min1, Real max1, Real min2, Real max2, Real dim, int& period) const {
    if (std::isinf(min1) || std::isinf(min2)) { period = 0; return true; }
    int pmn1_2, pmx1_2, pmn2_1, pmx2_1;
    Real mn1_2 = cellWrapRel(min1, min2, min2 + dim, pmn1_2); (void)cellWrapRel(max1, min2, min2 + dim, pmx1_2);
    Real mn2_1 = cellWrapRel(min2, min1, min1 + dim, pmn2_1); (void)cellWrapRel(max2, min1, min1 + dim, pmx2_1);
    if (pmn1_2 != pmx1_2 && pmn2_1 != pmx2_1)
      throw std::runtime_error("InsertionSortCollider: ambiguous periodic contact ##" + std::to_string(id1) + "+" + std::to_string(id2));
    else if (pmn2_1 == pmx2_1) { period = -pmn2_1; return mn2_1 < (max1 - min1); }
    else { period = pmn1_2; return mn1_2 < (max2 - min2); }
  }
  bool spatialOverlapPeri(int id1, int id2, int* periods) const {  // InsertionSortCollider.cpp:1042-1060
    for (int axis = 0; axis < 3; axis++)
      if (!spatialOverlapPeri_axis(axis, id1, id2, minima[3 * id1 + axis], maxima[3 * id1 + axis], minima[3 * id2 + axis], maxima[3 * id2 + axis], cell.size[axis], periods[axis])) return false;
    return true;
  }
  void makeContact(int id1, int id2, const int* periods) {  // canonicalised makeContactLater + Contact::swapOrder (Contact.cpp:12-16)
    int a = id1, b = id2; canonical(a, b);
    Contact C; C.idA = a; C.idB = b;
    int sgn = (a == id1) ? 1 : -1;
    for (int k = 0; k < 3; k++) C.cellDist[k] = periods ? sgn * periods[k] : 0;
    C.stepCreated = (int)step;
    con.emplace(pairKey(a, b), C); linDirty = true;
  }
  // InsertionSortCollider::handleBoundInversion (InsertionSortCollider.cpp:78-113)
  void handleBoundInversion(int id1, int id2, bool separating) {
    bool overlap = separating ? false : spatialOverlap(id1, id2);
    Contact* C = findContact(id1, id2);
    bool hasCon = (C != nullptr);
    if (!overlap && !hasCon) return;
    if (overlap && hasCon) { C->overlapBB = true; return; }
    if (overlap) {
      if (!mayCollide(id1, id2) || !hasFunctor(id1, id2)) return;
      makeContact(id1, id2, nullptr);
      return;
    }
    if (!C->isReal()) { int a = id1, b = id2; canonical(a, b); con.erase(pairKey(a, b)); linDirty = true; }
    else C->overlapBB = false;
  }
  // insertionSort_part serial (InsertionSortCollider.cpp:189-222)
  void insertionSort(std::vector<Bound>& v, bool doCollide) {
    long n = (long)v.size();
    for (long i = 1; i < n; i++) {
      if (!(v[i - 1].gt(v[i]))) continue;
      const Bound viInit = v[i]; long j = i - 1;
      while (j >= 0 && v[j].gt(viInit)) {
        v[j + 1] = v[j];
        if (viInit.isMin != v[j].isMin && doCollide && viInit.hasBB && v[j].hasBB && viInit.id != v[j].id)
          handleBoundInversion(viInit.id, v[j].id, /*separating*/ !viInit.isMin);
        j--;
      }
      v[j + 1] = viInit;
    }
  }
  // ContactContainer::removePending + shouldBeRemoved (ContactContainer.hpp:78-93, InsertionSortCollider.hpp:213-218)
  void removePending() {
    for (uint64_t k : pending) {
      auto it = con.find(k);
      if (it == con.end()) continue;
      Contact& C = it->second;
      if (C.isReal()) continue;  // re-created in between (cannot happen within one step)
      bool remove;
      if (!P.periodic) remove = !spatialOverlap(C.idA, C.idB);
      else { int per[3]; remove = !spatialOverlapPeri(C.idA, C.idB, per); }
      if (remove) { con.erase(it); linDirty = true; }
    }
    pending.clear();
  }
  void colliderRun() {  // InsertionSortCollider::run (InsertionSortCollider.cpp:386-613)
    bool fullRun = false;
    long nPar = (long)par.size();
    if (conDirty) { fullRun = true; conDirty = false; }
    if ((long)BB[0].size() != 2 * nPar) fullRun = true;
    // updateBboxes_doFullRun (InsertionSortCollider.cpp:308-358)
    setVerletDist();
    bool recompute = (P.verletDist == 0) ? true : anyAabbDirty();
    if (recompute) { updateAabbAll(); fullRun = true; }
    if (!fullRun) { removePending(); return; }
    nFullRuns++;
    bool doInitSort = false;
    if ((long)BB[0].size() != 2 * nPar) {
      if (2 * nPar < (long)BB[0].size()) for (int i = 0; i < 3; i++) BB[i].clear();
      if (2 * nPar - (long)BB[0].size() > 200 || BB[0].size() == 0) doInitSort = true;
      for (int i = 0; i < 3; i++) for (long id = (long)BB[i].size() / 2; id < nPar; id++) { BB[i].push_back(Bound{0, (int)id, true, true}); BB[i].push_back(Bound{0, (int)id, false, true}); }
    }
    minima.resize(3 * nPar); maxima.resize(3 * nPar);
    for (long id = 0; id < nPar; id++) { Aabb& b = bbOf((int)id); for (int k = 0; k < 3; k++) { minima[3 * id + k] = b.mn[k]; maxima[3 * id + k] = b.mx[k]; } }
    if (P.periodic) { removePending(); colliderPeriFromScratch(); return; }
    for (int j = 0; j < 3; j++) for (auto& b : BB[j]) { b.coord = b.isMin ? minima[3 * b.id + j] : maxima[3 * b.id + j]; b.hasBB = true; }
    removePending();
    if (!doInitSort) { for (int i = 0; i < 3; i++) insertionSort(BB[i], true); }
    else {
      for (int i = 0; i < 3; i++) std::sort(BB[i].begin(), BB[i].end(), [](const Bound& a, const Bound& b) { return a.lt(b); });
      std::vector<Bound>& V = BB[0];  // sortAxis=0
      for (long i = 0; i < 2 * nPar; i++) {
        if (!(V[i].isMin && V[i].hasBB)) continue;
        const int iid = V[i].id;
        for (long j = i + 1; j < 2 * nPar && V[j].id != iid; j++) {
          if (!V[j].isMin) continue;
          handleBoundInversion(iid, V[j].id, false);
        }
      }
    }
  }
  // Periodic variant: contact set := {pairs with spatialOverlapPeri && mayCollide} ∪ {real}; cellDist from spatialOverlapPeri at creation.
  // Candidate search: uniform grid over wrapped unsheared AABB centres (superset), exact test by spatialOverlapPeri.
  void colliderPeriFromScratch() {
    long nPar = (long)par.size();
    // mark all as non-overlapping, then re-mark
    for (auto& kv : con) kv.second.overlapBB = false;
    // grid
    Real maxExt = 0;
    for (long id = 0; id < nPar; id++) for (int k = 0; k < 3; k++) { Real e = maxima[3 * id + k] - minima[3 * id + k]; if (!std::isinf(e)) maxExt = std::max(maxExt, e); }
    int nc[3];
    for (int k = 0; k < 3; k++) { nc[k] = std::max(1, (int)std::floor(cell.size[k] / std::max(maxExt, 1e-300))); nc[k] = std::min(nc[k], 256); }
    auto cidx = [&](int ix, int iy, int iz) { return ((long)iz * nc[1] + iy) * nc[0] + ix; };
    std::vector<std::vector<int>> cells((size_t)nc[0] * nc[1] * nc[2]);
    std::vector<int> infIds;
    auto cellOf = [&](long id, int* c) -> bool {
      for (int k = 0; k < 3; k++) {
        Real mn = minima[3 * id + k], mx = maxima[3 * id + k];
        if (std::isinf(mn) || std::isinf(mx)) return false;
        Real ctr = .5 * (mn + mx); Real xn = ctr / cell.size[k]; xn = xn - std::floor(xn);
        c[k] = std::min(nc[k] - 1, std::max(0, (int)std::floor(xn * nc[k])));
      }
      return true;
    };
    for (long id = 0; id < nPar; id++) { int c[3]; if (cellOf(id, c)) cells[cidx(c[0], c[1], c[2])].push_back((int)id); else infIds.push_back((int)id); }
    auto tryPair = [&](int a, int b) {
      if (a == b) return;
      if (!mayCollide(a, b) || !hasFunctor(a, b)) return;  // mask check first: avoids the ambiguity throw for excluded pairs is NOT what the reference does (hpp #if 0) — but those pairs never exist on this path
      int per[3];
      int lo = std::min(a, b), hi = std::max(a, b);
      if (!spatialOverlapPeri(lo, hi, per)) return;
      Contact* C = findContact(lo, hi);
      if (C) { C->overlapBB = true; return; }
      makeContact(lo, hi, per);
    };
    for (int iz = 0; iz < nc[2]; iz++) for (int iy = 0; iy < nc[1]; iy++) for (int ix = 0; ix < nc[0]; ix++) {
      auto& A = cells[cidx(ix, iy, iz)];
      if (A.empty()) continue;
      // neighbour cells (with wrap); visit each unordered cell pair once via index comparison
      int seenN = 0; long seen[27];
      for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
        int jx = (ix + dx + nc[0]) % nc[0], jy = (iy + dy + nc[1]) % nc[1], jz = (iz + dz + nc[2]) % nc[2];
        long cj = cidx(jx, jy, jz), ci = cidx(ix, iy, iz);
        if (cj < ci) continue;
        bool dup = false; for (int s = 0; s < seenN; s++) if (seen[s] == cj) dup = true;
        if (dup) continue; seen[seenN++] = cj;
        auto& B = cells[cj];
        if (cj == ci) { for (size_t p = 0; p < A.size(); p++) for (size_t q = p + 1; q < A.size(); q++) tryPair(A[p], A[q]); }
        else { for (int a : A) for (int b : B) tryPair(a, b); }
      }
    }
    for (int w : infIds) for (long id = 0; id < nPar; id++) tryPair(w, (int)id);
    // non-real contacts whose boxes separated are removed (handleBoundInversionPeri: !overlap && hasCon && !real)
    for (auto it = con.begin(); it != con.end();) { if (!it->second.overlapBB && !it->second.isReal()) { it = con.erase(it); linDirty = true; } else ++it; }
  }

  // ------------------------------------------------------------------------------------------
  // L6Geom (src/dem/L6Geom.cpp)
  // ------------------------------------------------------------------------------------------
  static void setInitialLocalCoords(Contact& g, const V3& locX) {  // L6Geom.cpp:25-36
    V3 locY = cross(locX, std::abs(locX[1]) < std::abs(locX[2]) ? V3(0, 1, 0) : V3(0, 0, 1));
    locY -= locX * dot(locY, locX);
    locY = locY / norm(locY);
    V3 locZ = cross(locX, locY);
    g.trsf.c[0] = locX; g.trsf.c[1] = locY; g.trsf.c[2] = locZ;
  }
  // Cg2_Any_Any_L6Geom__Base::handleSpheresLikeContact (L6Geom.cpp:49-216), approxMask==0
  void handleSpheresLikeContact(Contact& C, const V3& pos1, const V3& vel1, const V3& angVel1, const V3& pos2, const V3& vel2,
                                const V3& angVel2, const V3& normal, const V3& contPt, Real uN, Real r1, Real r2) {
    if (!C.hasGeom) {
      C.hasGeom = true;
      C.gvel = V3(); C.gangVel = V3();
      setInitialLocalCoords(C, normal);
      C.uN = uN;
      if (P.iniLensTouch) { C.lens[0] = std::abs(r1); C.lens[1] = std::abs(r2); }
      else { C.lens[0] = std::abs(r1) + uN / 2; C.lens[1] = std::abs(r2) + uN / 2; }
      if (C.lens[0] < 0) C.lens[0] = C.lens[1];
      if (C.lens[1] < 0) C.lens[1] = C.lens[0];
      C.contA = M_PI * pow2((r1 > 0 && r2 > 0) ? std::min(r1, r2) : (r1 > 0 ? r1 : r2));
      C.contPt = contPt;
      C.ori = quatFromMat(C.trsf);
      return;
    }
    const V3& currNormal = normal; const V3 prevNormal = C.trsf.c[0];
    const V3 prevContPt = C.contPt;
    const Real dt = P.dt;
    V3 shiftVel2 = P.periodic ? cell.intrShiftVel(C.cellDist) : V3();
    V3 normRotVec = cross(prevNormal, currNormal);
    V3 midNormal = .5 * (prevNormal + currNormal);
    if (norm(midNormal) == 0) midNormal = currNormal; else midNormal = midNormal / norm(midNormal);
    V3 normTwistVec = midNormal * dt * .5 * dot(midNormal, angVel1 + angVel2);
    const M3 prevTrsf = C.trsf;
    M3 midTrsf;
    midTrsf.c[0] = midNormal;
    midTrsf.c[1] = prevTrsf.c[1] - cross(prevTrsf.c[1], normRotVec + normTwistVec) / 2.;
    midTrsf.c[2] = cross(midTrsf.c[0], midTrsf.c[1]);
    M3 currTrsf;
    currTrsf.c[0] = currNormal;
    currTrsf.c[1] = prevTrsf.c[1] - cross(midTrsf.c[1], normRotVec + normTwistVec);
    currTrsf.c[2] = cross(currTrsf.c[0], currTrsf.c[1]);
    if (P.trsfRenorm > 0 && (step % P.trsfRenorm) == 0) {
      currTrsf.c[0] = currTrsf.c[0] / norm(currTrsf.c[0]);
      currTrsf.c[1] -= currTrsf.c[0] * dot(currTrsf.c[1], currTrsf.c[0]);
      currTrsf.c[1] = currTrsf.c[1] / norm(currTrsf.c[1]);
      currTrsf.c[2] = cross(currTrsf.c[0], currTrsf.c[1]);
    }
    V3 midContPt = .5 * (prevContPt + contPt);
    V3 midPos1 = pos1 - (dt / 2.) * vel1;
    V3 midPos2 = pos2 - (dt / 2.) * (vel2 + shiftVel2);
    V3 c1x = (P.noRatch && r1 > 0) ? (r1 * midNormal) : (midContPt - midPos1);
    V3 c2x = (P.noRatch && r2 > 0) ? (-r2 * midNormal) : (midContPt - midPos2);
    V3 relVel = ((vel2 + shiftVel2) + cross(angVel2, c2x)) - (vel1 + cross(angVel1, c1x));
    C.trsf = currTrsf;
    C.gvel = mulMtV(midTrsf, relVel);
    C.gangVel = mulMtV(midTrsf, angVel2 - angVel1);
    C.uN = uN;
    C.contPt = contPt;
    C.ori = quatFromMat(C.trsf);
  }
  // Cg2_Sphere_Sphere_L6Geom::go (Sphere.cpp:92-117)
  bool cg2SphereSphere(Contact& C, const Sphere& s1, const Sphere& s2, const V3& shift2) {
    const Real r1 = s1.radius, r2 = s2.radius;
    V3 relPos = (s2.pos + shift2) - s1.pos;
    Real unDistSq = sqNorm(relPos) - pow2(std::abs(P.distFactor) * (r1 + r2));
    if (unDistSq > 0 && !C.isReal()) return false;
    Real dist = norm(relPos);
    Real uN = dist - (r1 + r2);
    V3 normal = relPos / dist;
    V3 contPt = s1.pos + (r1 + 0.5 * uN) * normal;
    handleSpheresLikeContact(C, s1.pos, s1.vel, s1.angVel, s2.pos + shift2, s2.vel, s2.angVel, normal, contPt, uN, r1, r2);
    return true;
  }
  // Cg2_Wall_Sphere_L6Geom::go (Wall.cpp:50-84)
  bool cg2WallSphere(Contact& C, const Wall& wall, const Sphere& sphere, const V3& shift2) {
    if (P.periodic && cell.hasShear) throw std::logic_error("Cg2_Wall_Sphere_L6Geom does not handle periodic boundary conditions with skew.");
    const Real radius = sphere.radius; const int ax = wall.axis, sense = wall.sense;
    const V3& wallPos = wall.pos; V3 spherePos = sphere.pos + shift2;
    Real dist = spherePos[ax] - wallPos[ax];
    if (!C.isReal() && std::abs(dist) > radius) return false;
    V3 contPt = spherePos; contPt[ax] = wallPos[ax];
    V3 normal(0, 0, 0);
    if (sense == 0) {
      if (!C.hasGeom) normal[ax] = dist > 0 ? 1. : -1.;
      else normal[ax] = C.trsf.c[0][ax];
    } else normal[ax] = (sense == 1 ? 1. : -1);
    Real uN = normal[ax] * dist - radius;
    handleSpheresLikeContact(C, wallPos, wall.vel, wall.angVel, spherePos, sphere.vel, sphere.angVel, normal, contPt, uN, -radius, radius);
    return true;
  }
  // CompUtils::closestSegmentPt (CompUtils.cpp:59-65)
  static V3 closestSegmentPt(const V3& P_, const V3& A, const V3& B) {
    V3 BA = B - A;
    if (sqNorm(BA) == 0.) return A;
    Real u = (dot(P_, BA) - dot(A, BA)) / (sqNorm(BA));
    return A + std::min(1., std::max(0., u)) * BA;
  }
  // CompUtils::triangleBarycentrics (CompUtils.cpp:94-112); Matrix2 inverse = adj/det
  static V3 triangleBarycentrics(const V3& x, const V3& A, const V3& B, const V3& Cc) {
    V3 vB = B - A, vC = Cc - A, vX = x - A;
    Real m00 = dot(vB, vB), m01 = dot(vC, vB), m10 = dot(vB, vC), m11 = dot(vC, vC);
    Real det = m00 * m11 - m01 * m10; Real invdet = 1. / det;
    Real i00 = m11 * invdet, i01 = -m01 * invdet, i10 = -m10 * invdet, i11 = m00 * invdet;
    Real b0 = dot(vX, vB), b1 = dot(vX, vC);
    Real u = i00 * b0 + i01 * b1, v = i10 * b0 + i11 * b1;
    return V3(1 - u - v, u, v);
  }
  static V3 facetNormal(const Facet& f) { return normalized(cross(f.v[1] - f.v[0], f.v[2] - f.v[0])); }  // Facet.cpp:74-77
  static void facetOuterVectors(const Facet& f, V3* o) {  // Facet.cpp:113-118
    V3 nn = cross(f.v[1] - f.v[0], f.v[2] - f.v[0]);
    o[0] = cross(f.v[1] - f.v[0], nn); o[1] = cross(f.v[2] - f.v[1], nn); o[2] = cross(f.v[0] - f.v[2], nn);
  }
  // Facet::interpolatePtLinAngVel (Facet.cpp:161-177)
  static void facetLinAngVel(const Facet& f, const V3& x, V3& linVel, V3& angVel) {
    V3 a = triangleBarycentrics(x, f.v[0], f.v[1], f.v[2]);
    linVel = (a[0] * f.vvel[0] + a[1] * f.vvel[1]) + a[2] * f.vvel[2];
    angVel = (cross(f.v[0] - x, f.vvel[0]) + cross(f.v[1] - x, f.vvel[1])) + cross(f.v[2] - x, f.vvel[2]);
    if (f.fakeVel == V3(0, 0, 0)) return;
    if (!std::isnan(f.fakeVel[0])) { linVel = linVel + f.fakeVel; return; }
    V3 fN = facetNormal(f);
    linVel = dot(linVel, fN) * fN;
  }
  // Facet::adjustBoundaryContactGeom_impl (Facet.cpp:239-265)
  static void adjustImpl(const V3& fN, const V3& edgePt, const V3& unitOutVec, const V3& otherCenter, Real nnLim, V3& normal, V3& contPt) {
    if (std::isnan(nnLim)) return;
    Real nn = dot(normal, fN);
    if ((nn >= 0 && nnLim > 0 && nn >= nnLim) || (nn < 0 && nnLim < 0 && nn <= nnLim)) return;
    Real noLim = (std::abs(nnLim) >= 1. ? 1. : std::sqrt(1. - nnLim * nnLim));
    Real nno0, nno1;
    if ((nnLim >= 0 && nn <= -noLim) || (nnLim <= 0 && nn >= noLim)) { nno0 = -nnLim; nno1 = 0; } else { nno0 = nnLim; nno1 = noLim; }
    normal = nno0 * fN + nno1 * unitOutVec;
    contPt = otherCenter - normal * (dot(otherCenter - edgePt, normal));
  }
  // Facet::adjustBoundaryContactGeom (Facet.cpp:220-237)
  static void adjustBoundary(const Facet& f, int e0, int e1, const V3& fN, const V3* outVec, const V3& otherCenter, V3& normal, V3& contPt) {
    if (std::isnan(f.n21lim[e0])) return;
    if (e1 < 0) { adjustImpl(fN, f.v[e0], normalized(outVec[e0]), otherCenter, f.n21lim[e0], normal, contPt); return; }
    if (std::isnan(f.n21lim[e1])) return;
    V3 out0 = normalized(outVec[e0]), out1 = normalized(outVec[e1]);
    Real p0 = dot(out0, normal), pMin = dot(out0, out1), pMax = 1.;
    Real t01 = (p0 - pMin) / (pMax - pMin);
    adjustImpl(fN, f.v[e1], (t01 * normalized(outVec[e0]) + (1 - t01) * normalized(outVec[e1])), otherCenter, (t01 * f.n21lim[e0] + (1 - t01) * f.n21lim[e1]), normal, contPt);
  }
  // Cg2_Facet_Sphere_L6Geom::go (Facet.cpp:333-391)
  bool cg2FacetSphere(Contact& C, const Facet& f, const Sphere& s, const V3& shift2) {
    const V3 sC = s.pos + shift2;
    V3 fNormal = facetNormal(f);
    Real planeDist = dot(sC - f.v[0], fNormal);
    if (std::abs(planeDist) > (s.radius + f.halfThick) && !C.isReal()) return false;
    V3 fC = sC - planeDist * fNormal;
    V3 outVec[3]; facetOuterVectors(f, outVec);
    Real ll[3];
    for (int i = 0; i < 3; i++) ll[i] = dot(outVec[i], fC - f.v[i]);
    short w = (ll[0] > 0 ? 1 : 0) + (ll[1] > 0 ? 2 : 0) + (ll[2] > 0 ? 4 : 0);
    V3 contPt;
    switch (w) {
      case 0: contPt = fC; break;
      case 1: contPt = closestSegmentPt(fC, f.v[0], f.v[1]); break;
      case 2: contPt = closestSegmentPt(fC, f.v[1], f.v[2]); break;
      case 4: contPt = closestSegmentPt(fC, f.v[2], f.v[0]); break;
      case 3: contPt = f.v[1]; break;
      case 5: contPt = f.v[0]; break;
      case 6: contPt = f.v[2]; break;
      default: throw std::logic_error("Cg2_Facet_Sphere_L3Geom: Impossible sphere-facet intersection (all points are outside the edges).");
    }
    V3 normal = sC - contPt;
    if (sqNorm(normal) > pow2(s.radius + f.halfThick) && !C.isReal()) return false;
    Real dist = norm(normal);
    if (dist != 0) normal = normal / dist;
    else normal = (C.hasGeom ? quatRotate(C.ori, V3(1, 0, 0)) : V3(1, 0, 0));
    static const short edgeNums[7][2] = {{-1, -1}, {0, -1}, {1, -1}, {0, 1}, {2, -1}, {2, 0}, {1, 2}};
    if (w > 0) adjustBoundary(f, edgeNums[w][0], edgeNums[w][1], fNormal, outVec, sC, normal, contPt);
    if (f.halfThick > 0) contPt += normal * f.halfThick;
    Real uN = dist - s.radius - f.halfThick;
    V3 linVel, angVel; facetLinAngVel(f, contPt, linVel, angVel);
    handleSpheresLikeContact(C, contPt, linVel, angVel, sC, s.vel, s.angVel, normal, contPt, uN, std::max(f.halfThick, s.radius), s.radius);
    return true;
  }

  // ------------------------------------------------------------------------------------------
  // Cp2 functors
  // ------------------------------------------------------------------------------------------
  void cp2Frict(Contact& C) {  // Cp2_FrictMat_FrictPhys::updateFrictPhys (FrictMat.cpp:20-44)
    const Material& m1 = mats[matOf(C.idA)]; const Material& m2 = mats[matOf(C.idB)];
    Real l0 = C.lens[0], l1 = C.lens[1], A = C.contA;
    C.kn = 1 / (1 / (m1.young * A / l0) + 1 / (m2.young * A / l1));
    C.kt = .5 * (m1.ktDivKn + m2.ktDivKn) * C.kn;
    C.tanPhi = std::min(m1.tanPhi, m2.tanPhi);
    C.hasPhys = true; C.force = V3(); C.torque = V3();
  }
  Real massOf(int id) const { const PRef& p = par[id]; return p.kind == SHAPE_SPHERE ? sph[p.idx].mass : 0.; }
  void cp2Hertz(Contact& C) {  // Cp2_HertzMat_HertzPhys::go (Hertz.cpp:34-77)
    const Material& mat1 = mats[matOf(C.idA)]; const Material& mat2 = mats[matOf(C.idB)];
    if (mat1.kind != MAT_HERTZ || mat2.kind != MAT_HERTZ) throw std::runtime_error("ContactLoop: no Contact.phys created (Cp2_HertzMat_HertzPhys needs HertzMat on both particles)");
    const Real r1 = C.lens[0], r2 = C.lens[1]; const Real E1 = mat1.young, E2 = mat2.young; const Real nu1 = P.poisson, nu2 = P.poisson;
    C.K = (4 / 3.) * 1. / ((1 - nu1 * nu1) / E1 + (1 - nu2 * nu2) / E2);
    C.R = 1. / (1. / r1 + 1. / r2);
    Real kn0 = C.K * std::sqrt(C.R);
    C.gamma = std::min(mat1.surfEnergy, mat2.surfEnergy);
    C.alpha = .5 * (mat1.alpha + mat2.alpha);
    Real G1 = E1 / (2 * (1 + nu1)); Real G2 = E2 / (2 * (1 + nu2));
    Real G = .5 * (G1 + G2);
    Real nu = .5 * (nu1 + nu2);
    C.kt0 = 2 * std::sqrt(4 * C.R) * G / (2 - nu);
    C.tanPhi = std::min(mat1.tanPhi, mat2.tanPhi);
    const Real en = P.en;
    if (en > 0 && en < 1.) {
      const Real m1 = massOf(C.idA), m2 = massOf(C.idB);
      Real mbar = (m1 <= 0 && m2 > 0) ? m2 : ((m1 > 0 && m2 <= 0) ? m1 : (m1 * m2) / (m1 + m2));
      Real viscAlpha = -std::sqrt(5) * std::log(en) / (std::sqrt(pow2(std::log(en)) + pow2(M_PI)));
      C.alpha_sqrtMK = std::max(0., viscAlpha * std::sqrt(mbar * kn0));
    } else C.alpha_sqrtMK = 0.0;
    C.hasPhys = true; C.force = V3(); C.torque = V3(); C.kn = NaN; C.kt = NaN;
  }

  // ------------------------------------------------------------------------------------------
  // Law2 functors
  // ------------------------------------------------------------------------------------------
  bool lawIdealElPl(Contact& C) {  // Law2_L6Geom_FrictPhys_IdealElPl::go (IdealElPl.cpp:15-125); iniEqlb, trackEnergy not restated
    Real uN = C.uN;
    if (uN > 0 && !P.noBreak) {
      if (P.noSlip && trackEnergy) {   // IdealElPl.cpp:28-40: the elastic potential of the breaking contact is lost
        const Real Fn = C.kn * uN; const Real t0 = C.force[1] + (P.dt * C.kt) * C.gvel[1], t1 = C.force[2] + (P.dt * C.kt) * C.gvel[2];
        eadd(EN_BROKEN, .5 * (pow2(Fn) / C.kn + (t0 * t0 + t1 * t1) / C.kt));
      }
      return false;
    }
    C.torque = V3();
    C.force[0] = C.kn * uN;
    Real& Ft0 = C.force[1]; Real& Ft1 = C.force[2];
    if (P.noFrict || C.tanPhi == 0.) { Ft0 = 0; Ft1 = 0; }
    else {
      const Real vT0 = C.gvel[1], vT1 = C.gvel[2];
      Ft0 += (P.dt * C.kt) * vT0; Ft1 += (P.dt * C.kt) * vT1;
      Real maxFt = std::abs(C.force[0]) * C.tanPhi;
      Real FtSq = Ft0 * Ft0 + Ft1 * Ft1;
      if (FtSq > maxFt * maxFt && !P.noSlip) {
        Real FtNorm = std::sqrt(FtSq);
        Real ratio = maxFt / FtNorm;
        if (FtNorm == 0.) ratio = 0.;
        if (trackEnergy) eadd(EN_PLAST, C.kt != 0. ? (maxFt) * (FtNorm - maxFt) / C.kt : 0.);   // IdealElPl.cpp:64-75
        Ft0 *= ratio; Ft1 *= ratio;
      }
      if (std::isnan(C.force[0]) || std::isnan(Ft0) || std::isnan(Ft1)) throw std::runtime_error("NaN force in contact ##" + std::to_string(C.idA) + "+" + std::to_string(C.idB));
    }
    if (trackEnergy) {   // IdealElPl.cpp:85-92
      Real elast = 0.5 * (pow2(C.force[0]) / C.kn + (C.kt != 0. ? (Ft0 * Ft0 + Ft1 * Ft1) / C.kt : 0.));
      if (std::isnan(elast)) elast = 0.;
      eadd(EN_ELAST, elast);
    }
    if (P.relRollStiff > 0. && P.rollTanPhi > 0.) {
      Real charLen = C.lens[0] + C.lens[1];
      if (charLen <= 0) throw std::runtime_error("charLen<=0 !?");
      Real kr = C.kn * charLen;
      if (P.relTwistStiff > 0) {
        C.torque[0] += P.dt * P.relTwistStiff * kr * C.gangVel[0];
        Real maxTt = std::abs(C.force[0] * P.rollTanPhi * charLen);
        if (std::abs(C.torque[0]) > maxTt) C.torque[0] = std::copysign(maxTt, C.torque[0]);
      }
      Real& Tr0 = C.torque[1]; Real& Tr1 = C.torque[2];
      Tr0 += (P.dt * kr) * C.gangVel[1]; Tr1 += (P.dt * kr) * C.gangVel[2];
      Real maxTr = std::max(0., std::abs(C.force[0]) * P.rollTanPhi * charLen);
      Real TrSq = Tr0 * Tr0 + Tr1 * Tr1;
      if (TrSq > maxTr * maxTr) { Real TrNorm = std::sqrt(TrSq); Real ratio = maxTr / TrNorm; if (TrNorm == 0.) ratio = 0.; Tr0 *= ratio; Tr1 *= ratio; }
    } else C.torque = V3();
    return true;
  }
  static Real pown_sqrt3(Real x) { Real s = std::sqrt(x); return (s * s) * s; }  // Hertz.hpp:70 pow_i_2(x,3)=pown(sqrt(x),3)=(s*s)*s (Math.hpp pown: d=1,r=1 → pown(s*s,1)*s)
  bool lawHertzDMT(Contact& C) {  // Law2_L6Geom_HertzPhys_DMT::go, DMT/Hertz branch (Hertz.cpp:85-114,170-210)
    Real& Fn = C.force[0]; Real& Ft0 = C.force[1]; Real& Ft1 = C.force[2];
    C.torque = V3();
    const Real dt = P.dt; const Real velN = C.gvel[0]; const Real vT0 = C.gvel[1], vT1 = C.gvel[2];
    Real kn0 = C.K * std::sqrt(C.R);
    C.kn = (3 / 2.) * kn0 * std::sqrt(-C.uN);
    Real Fne, Fna = 0.;
    if (C.alpha == 0. || C.gamma == 0.) {
      if (C.uN > 0) return false;
      Fne = -kn0 * pown_sqrt3(-C.uN);
      Fna = 4 * M_PI * C.R * C.gamma;
    } else throw std::runtime_error("Schwarz model (alpha!=0 && gamma!=0) is out of scope (Hertz.cpp:115-168)");
    Real eta = (C.alpha_sqrtMK > 0 ? C.alpha_sqrtMK * std::sqrt(std::sqrt(std::max(-C.uN, 0.))) : 0.);
    Real Fnv = eta * velN;
    if (C.alpha == 0. && P.noAttraction && Fne + Fnv > 0) Fnv = -Fne;
    Fn = Fne + Fna + Fnv;
    if (trackEnergy) eadd(EN_VISCN, Fnv * velN * dt);   // Hertz.cpp:181
    C.kt = C.kt0 * std::sqrt(C.uN < 0 ? -C.uN : 0);
    Ft0 += (dt * C.kt) * vT0; Ft1 += (dt * C.kt) * vT1;
    Real maxFt = std::abs(std::min(0., Fn) * C.tanPhi);
    Real FtSq = Ft0 * Ft0 + Ft1 * Ft1;
    if (FtSq > pow2(maxFt)) {
      Real FtNorm = std::sqrt(FtSq); Real ratio = maxFt / FtNorm;
      if (trackEnergy) eadd(EN_PLAST, (.5 * (FtNorm - maxFt) + maxFt) * (FtNorm - maxFt) / C.kt);   // Hertz.cpp:193
      Ft0 *= ratio; Ft1 *= ratio;
    } else {
      Ft0 += eta * vT0; Ft1 += eta * vT1;
      if (trackEnergy) eadd(EN_VISCT, eta * (vT0 * vT0 + vT1 * vT1) * dt);   // Hertz.cpp:199
    }
    if (trackEnergy && C.uN < 0) {   // Hertz.cpp:203-207; normalElasticEnergy = kn0*(2/5)*pown(sqrt(delta),5) (Hertz.hpp:70,76)
      const Real sq = std::sqrt(-C.uN);
      eadd(EN_ELAST, kn0 * (2 / 5.) * (((sq * sq) * (sq * sq)) * sq) + 0.5 * (Ft0 * Ft0 + Ft1 * Ft1) / C.kt);
    }
    return true;
  }

  // ------------------------------------------------------------------------------------------
  // ContactLoop::run (ContactLoop.cpp:93-268) + applyForceUninodal (ContactLoop.cpp:270-276, Contact.cpp:24-35)
  // ------------------------------------------------------------------------------------------
  void rebuildLin() { lin.clear(); lin.reserve(con.size()); for (auto& kv : con) lin.push_back(&kv.second); linDirty = false; }
  void forceTorqueBranch(const Contact& C, bool isPA, const V3& nodePos, V3& F, V3& T, V3& xc) const {
    int sign = isPA ? 1 : -1;
    F = quatRotate(C.ori, C.force) * (Real)sign;
    T = (C.torque == V3(0, 0, 0) ? V3() : quatRotate(C.ori, C.torque)) * (Real)sign;
    xc = C.contPt - (nodePos + ((!isPA && P.periodic) ? cell.intrShiftPos(C.cellDist) : V3()));
  }
  void contactLoopRun() {
    if (linDirty) rebuildLin();
    const long n = (long)lin.size();
    std::vector<char> toRemove(n, 0);
    std::string errLocal;
    #pragma omp parallel for schedule(guided)
    for (long i = 0; i < n; i++) {
      Contact& C = *lin[i];
      try {
        const PRef& pa = par[C.idA]; const PRef& pb = par[C.idB];
        V3 shift2 = P.periodic ? cell.intrShiftPos(C.cellDist) : V3();
        const Sphere& sB = sph[pb.idx];
        bool geomCreated;
        if (pa.kind == SHAPE_SPHERE) {
          const Sphere& sA = sph[pa.idx];
          // minDist00Sq shortcut (ContactLoop.cpp:162-165, Sphere.cpp:86-89)
          if (!C.isReal() && C.stepCreated != step) {
            Real minDist00Sq = pow2(std::abs(P.distFactor * (sA.radius + sB.radius)));
            if (sqNorm(sA.pos - (sB.pos + shift2)) > minDist00Sq) continue;
          }
          geomCreated = cg2SphereSphere(C, sA, sB, shift2);
        } else if (pa.kind == SHAPE_WALL) geomCreated = cg2WallSphere(C, walls[pa.idx], sB, shift2);
        else geomCreated = cg2FacetSphere(C, facets[pa.idx], sB, shift2);
        if (!geomCreated) continue;
        if (!C.hasPhys) {
          C.stepCreated = (int)step;
          if (P.law == LAW_IDEALELPL) cp2Frict(C); else cp2Hertz(C);
        }
        bool keep = (P.law == LAW_IDEALELPL) ? lawIdealElPl(C) : lawHertzDMT(C);
        if (!keep) { C.reset(); toRemove[i] = 1; continue; }
        if (P.applyForces && !P.deterministic) applyForcesAtomic(C);
      } catch (std::exception& e) {
        #pragma omp critical
        errLocal = e.what();
      }
    }
    if (!errLocal.empty()) throw std::runtime_error(errLocal);
    for (long i = 0; i < n; i++) if (toRemove[i]) pending.push_back(pairKey(lin[i]->idA, lin[i]->idB));
    // deterministic pass.  The reference (ContactLoop.cpp:257-264) walks contacts in container order (creation order —
    // not reproducible); the oracle's canonical order is, per particle, ascending id of the OTHER particle.
    if (P.deterministic && P.applyForces) applyForcesDeterministic();
  }
  void applyForcesAtomic(const Contact& C) {
    for (int side = 0; side < 2; side++) {
      int id = side == 0 ? C.idA : C.idB; const PRef& p = par[id];
      if (p.kind == SHAPE_FACET) continue;  // multinodal: nothing applied (ContactLoop.cpp:272)
      V3 F, T, xc;
      V3& pf = p.kind == SHAPE_SPHERE ? sph[p.idx].force : walls[p.idx].force;
      V3& pt = p.kind == SHAPE_SPHERE ? sph[p.idx].torque : walls[p.idx].torque;
      const V3& pos = p.kind == SHAPE_SPHERE ? sph[p.idx].pos : walls[p.idx].pos;
      forceTorqueBranch(C, side == 0, pos, F, T, xc);
      V3 TT = cross(xc, F) + T;
      for (int k = 0; k < 3; k++) {
        #pragma omp atomic
        pf[k] += F[k];
        #pragma omp atomic
        pt[k] += TT[k];
      }
    }
  }
  void applyForcesDeterministic() {
    // incidence lists: (particle id) -> [(other id, contact*, side)]
    struct Inc { int other; const Contact* C; bool isPA; };
    std::vector<std::vector<Inc>> inc(par.size());
    for (auto& kv : con) { const Contact& C = kv.second; if (!C.isReal()) continue; inc[C.idA].push_back({C.idB, &C, true}); inc[C.idB].push_back({C.idA, &C, false}); }
    #pragma omp parallel for schedule(guided)
    for (long id = 0; id < (long)par.size(); id++) {
      const PRef& p = par[id];
      if (p.kind == SHAPE_FACET || inc[id].empty()) continue;
      std::sort(inc[id].begin(), inc[id].end(), [](const Inc& a, const Inc& b) { return a.other < b.other; });
      V3& pf = p.kind == SHAPE_SPHERE ? sph[p.idx].force : walls[p.idx].force;
      V3& pt = p.kind == SHAPE_SPHERE ? sph[p.idx].torque : walls[p.idx].torque;
      const V3& pos = p.kind == SHAPE_SPHERE ? sph[p.idx].pos : walls[p.idx].pos;
      for (const Inc& e : inc[id]) {
        V3 F, T, xc; forceTorqueBranch(*e.C, e.isPA, pos, F, T, xc);
        pf += F; pt += cross(xc, F) + T;
      }
    }
  }

  // ------------------------------------------------------------------------------------------
  // Leapfrog::run, spherical branch (Leapfrog.cpp:104-255,33-48,258-297)
  // ------------------------------------------------------------------------------------------
  static Real Sign(Real f) { if (f < 0) return -1; if (f > 0) return 1; return 0; }  // Math.hpp:145
  // ---- Impose hooks; time-dependent values are evaluated once per step (they depend on Scene.time and Scene.dt only)
  struct ImpStep { double vMag = 0; double vImp[3] = {0, 0, 0}, wImp[3] = {0, 0, 0}; V3 nextPos; Quat nextOri; };
  std::vector<ImpStep> impStep;
  void imposePrepareStep() {
    impStep.assign(imposes.size(), ImpStep());
    for (size_t i = 0; i < imposes.size(); i++) {
      ImposeDesc& d = imposes[i]; ImpStep& r = impStep[i];
      switch (d.kind) {
        case IMP_HARMONIC: { const Real omega = 2 * M_PI * d.v[0]; r.vMag = d.v[1] * omega * std::cos(omega * (time - d.v[5])); break; }      // Impose.hpp:8-9
        case IMP_ALIGNED_HARMONIC: for (int ax = 0; ax < 3; ax++) { if (std::isnan(d.v[ax]) || std::isnan(d.v[3 + ax])) { r.vImp[ax] = NaN; continue; } const Real omega = 2 * M_PI * d.v[ax]; r.vImp[ax] = d.v[3 + ax] * omega * std::cos(omega * time); } break;      // Impose.hpp:53-57
        case IMP_VARIABLE_VEL3D: {      // Impose.cpp:196-205,216
          if (d.times.empty()) throw std::runtime_error("VariableVelocity3d: times must not be empty.");
          Real t = time + .5 * P.dt - d.v[0];
          if ((d.flags & 2) && (t < 0 || t > d.times.back())) { const Real norm = t / d.times.back(); t = (norm - std::floor(norm)) * d.times.back(); }
          linInterp(t, d.times, d.vals, 3, d.ipos, r.vImp);
          if (!d.vals2.empty()) linInterp(t, d.times, d.vals2, 3, d.ipos2, r.wImp);
          break;
        }
        case IMP_VARIABLE_ALIGNED_ROTATION: {      // Impose.cpp:95-100
          Real t = time; if ((d.flags & 1) && t > d.times.back()) t = std::fmod(t, d.times.back());
          linInterp(t, d.times, d.vals, 1, d.ipos, &r.vMag); break;
        }
        case IMP_VEL_READ_FORCE: if (d.nNodes > 0) { d.sumF = 0; d.dist += d.v[3] * P.dt; } break;
        case IMP_CIRCULAR_ORBIT: case IMP_STABLE_CIRCULAR_ORBIT: if (d.nNodes > 0) d.dist += d.v[3] * P.dt; break;
        case IMP_INTERPOLATED_MOTION: {      // Impose.cpp:121-125: v[0] = t0; vals = poss, vals2 = oris (x,y,z,w); ONE cookie for both lookups
          const Real nextT = time + P.dt - d.v[0];
          double np[3]; linInterp(nextT, d.times, d.vals, 3, d.ipos, np); r.nextPos = V3(np[0], np[1], np[2]);
          // linearInterpolateRel (LinearInterpolate.hpp:13-28) on the orientations
          size_t pos = d.ipos, iA, iB; const size_t n = d.times.size(); Real rel;
          if (nextT <= d.times[0]) { d.ipos = 0; iA = 0; iB = n > 1 ? 1 : 0; rel = 0; }
          else if (nextT >= d.times[n - 1]) { d.ipos = n - 2; iA = n > 1 ? pos : n - 1; iB = n - 1; rel = 1; }
          else { pos = std::min(pos, n - 2); while ((d.times[pos] > nextT) || (d.times[pos + 1] < nextT)) { if (d.times[pos] > nextT) pos--; else pos++; } d.ipos = pos; iA = pos; iB = pos + 1; rel = (nextT - d.times[pos]) / (d.times[pos + 1] - d.times[pos]); }
          // Quaternionr::slerp (Eigen Geometry/Quaternion.h): dot of the 4 coefficients (x,y,z,w) as Eigen's SSE2 reduction adds them: (xx'+zz') + (yy'+ww')
          const double* a = &d.vals2[4 * iA]; const double* b = &d.vals2[4 * iB];
          const Real one = 1. - std::numeric_limits<Real>::epsilon();
          const Real dd = (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + a[3] * b[3]), absD = std::abs(dd);
          Real s0, s1;
          if (absD >= one) { s0 = 1. - rel; s1 = rel; }
          else { const Real theta = std::acos(absD), sinTheta = std::sin(theta); s0 = std::sin((1. - rel) * theta) / sinTheta; s1 = std::sin(rel * theta) / sinTheta; }
          if (dd < 0) s1 = -s1;
          r.nextOri = Quat(s0 * a[3] + s1 * b[3], s0 * a[0] + s1 * b[0], s0 * a[1] + s1 * b[1], s0 * a[2] + s1 * b[2]);
          break;
        }      // .angle += omega*dt, once per step (Impose.cpp:36,47)      // the first node of the step resets (Impose.cpp:166-173)
        case IMP_READ_FORCE: if (d.nNodes > 0) { d.F = V3(); d.T = V3(); } break;
        default: break;
      }
    }
  }
  void imposeVelocity(int set, V3& vel, V3& angVel, const V3& pos, const Quat* nodeOri = nullptr) {
    for (int ix : imposeSets[set]) {
      const ImposeDesc& d = imposes[ix]; const ImpStep& r = impStep[ix];
      switch (d.kind) {
        case IMP_INTERPOLATED_MOTION: {      // Impose.cpp:120-133 (nextPos / nextOri of the step: imposePrepareStep)
          vel = (r.nextPos - pos) / P.dt;
          if (nodeOri) {
            const Quat rq = quatMul(r.nextOri, Quat(nodeOri->w, -nodeOri->x, -nodeOri->y, -nodeOri->z));      // q+ q0*
            // AngleAxisr(Quaternionr) of Eigen 3.3/3.4 (Geometry/AngleAxis.h): n = |vec|; angle = 2 atan2(n, |w|); axis = vec / (w < 0 ? -n : n); n == 0: angle 0, axis x
            Real n = norm(V3(rq.x, rq.y, rq.z)); V3 axis(1, 0, 0); Real angle = 0;
            if (n != 0) { angle = 2 * std::atan2(n, std::abs(rq.w)); if (rq.w < 0) n = -n; axis = V3(rq.x, rq.y, rq.z) / n; }
            angVel = (axis * angle) / P.dt;
          }
          break;
        }
        case IMP_CIRCULAR_ORBIT: {      // Impose.cpp:22-37; v = node.pos(3), omega; ori = node.ori; flags&1 rotate
          const Quat qc(d.ori.w, -d.ori.x, -d.ori.y, -d.ori.z);
          const V3 loc = quatRotate(qc, pos - V3(d.v[0], d.v[1], d.v[2]));      // local 2d position (head<2>)
          const Real theta = std::atan2(loc.y, loc.x), rho = std::sqrt(loc.x * loc.x + loc.y * loc.y);
          const Real v2t = d.v[3] * rho;
          const Real ttheta = theta + .5 * v2t * P.dt;      // sic: the tangential velocity, not omega
          const Quat qz(std::cos(.5 * ttheta), 0., 0., std::sin(.5 * ttheta));      // Quaternionr(AngleAxisr(ttheta, UnitZ))
          vel = quatRotate(d.ori, quatRotate(qz, V3(0., v2t, 0.)));
          if (d.flags & 1) angVel = quatRotate(d.ori, V3(0., 0., d.v[3]));
          break;
        }
        case IMP_STABLE_CIRCULAR_ORBIT: {      // Impose.cpp:38-48; v = node.pos(3), omega, radius
          const Quat qc(d.ori.w, -d.ori.x, -d.ori.y, -d.ori.z); const V3 c(d.v[0], d.v[1], d.v[2]);
          const V3 loc = quatRotate(qc, pos - c);                                  // Node::glob2loc (Field.hpp:92)
          const Real th1 = std::atan2(loc.y, loc.x) + d.v[3] * P.dt;              // CompUtils::cart2cyl (CompUtils.cpp:115-117), target angle
          const V3 pos1 = quatRotate(d.ori, V3(d.v[4] * std::cos(th1), d.v[4] * std::sin(th1), loc.z)) + c;      // cyl2cart (:120-122), loc2glob (Field.hpp:94)
          vel = (pos1 - pos) / P.dt;
          if (d.flags & 1) angVel = quatRotate(d.ori, V3(0., 0., d.v[3]));
          break;
        }
        case IMP_HARMONIC: case IMP_VEL_READ_FORCE: {
          const V3 dir(d.v[d.kind == IMP_HARMONIC ? 2 : 0], d.v[d.kind == IMP_HARMONIC ? 3 : 1], d.v[d.kind == IMP_HARMONIC ? 4 : 2]);
          const Real mag = d.kind == IMP_HARMONIC ? r.vMag : d.v[3];
          const bool lat = d.kind == IMP_HARMONIC ? !(d.flags & 1) : (d.flags & 1);      // HarmonicOscillation.perpFree (inverted) / VelocityAndReadForce.latBlock
          if (lat) vel = dir * mag;
          else { const V3 vLat = vel - dir * dot(vel, dir); vel = vLat + dir * mag; }      // Impose.hpp:11-12, Impose.cpp:158-161
          break;
        }
        case IMP_ALIGNED_HARMONIC: for (int ax = 0; ax < 3; ax++) if (!std::isnan(r.vImp[ax])) vel[ax] = r.vImp[ax]; break;
        case IMP_VARIABLE_VEL3D: {      // Impose.cpp:193-226
          const bool rot = !(d.ori.w == 1 && d.ori.x == 0 && d.ori.y == 0 && d.ori.z == 0), diff = d.flags & 1;
          const Quat qc(d.ori.w, -d.ori.x, -d.ori.y, -d.ori.z);
          V3 v = rot ? quatRotate(qc, vel) : vel;
          for (int ax = 0; ax < 3; ax++) v[ax] = (diff ? v[ax] : 0) + (std::isnan(r.vImp[ax]) ? (diff ? 0 : v[ax]) : r.vImp[ax]);
          vel = rot ? quatRotate(d.ori, v) : v;
          if (!d.vals2.empty()) {
            V3 av = rot ? quatRotate(qc, angVel) : angVel;
            for (int ax = 0; ax < 3; ax++) av[ax] = (diff ? av[ax] : 0) + (std::isnan(r.wImp[ax]) ? (diff ? 0 : av[ax]) : r.wImp[ax]);
            angVel = rot ? quatRotate(d.ori, av) : av;
          }
          break;
        }
        case IMP_VARIABLE_ALIGNED_ROTATION: angVel[(int)d.v[0]] = r.vMag; break;
        case IMP_LOCAL6: {      // Impose.cpp:72-82 (note: ori takes global to local here)
          const Quat qc(d.ori.w, -d.ori.x, -d.ori.y, -d.ori.z);
          for (int i = 0; i < 6; i++) {
            if (!((int)d.v[6 + i] & IMPW_VELOCITY)) continue;
            if (i < 3) { V3 loc = quatRotate(d.ori, vel); loc[i] = d.v[i]; vel = quatRotate(qc, loc); }
            else { V3 loc = quatRotate(d.ori, angVel); loc[i % 3] = d.v[i]; angVel = quatRotate(qc, loc); }
          }
          break;
        }
        default: break;
      }
    }
  }
  void imposeForce(int set, const V3& pos, V3& force, V3& torque) {
    for (int ix : imposeSets[set]) {
      const ImposeDesc& d = imposes[ix];
      switch (d.kind) {
        case IMP_CONSTANT_FORCE: force += V3(d.v[0], d.v[1], d.v[2]); break;
        case IMP_RADIAL_FORCE: {      // Impose.cpp:49-58
          const V3 A(d.v[0], d.v[1], d.v[2]), B(d.v[3], d.v[4], d.v[5]);
          const V3 ab = B - A; const V3 axis = ab / norm(ab);
          const V3 C = A + axis * dot(pos - A, axis);
          const V3 towards = C - pos;
          if (sqNorm(towards) == 0) break;
          force -= (towards / norm(towards)) * d.v[6];
          break;
        }
        case IMP_LOCAL6: {      // Impose.cpp:83-87: the value replaces whatever was there (gravity included)
          const Quat qc(d.ori.w, -d.ori.x, -d.ori.y, -d.ori.z);
          for (int i = 0; i < 6; i++) {
            if (!((int)d.v[6 + i] & IMPW_FORCE)) continue;
            V3 u; u[i % 3] = 1;
            if (i < 3) force = quatRotate(qc, d.v[i] * u); else torque = quatRotate(qc, d.v[i] * u);
          }
          break;
        }
        default: break;
      }
    }
  }
  void imposeReadForce(int set, const V3& pos, const V3& force, const V3& torque) {
    for (int ix : imposeSets[set]) {
      ImposeDesc& d = imposes[ix];
      if (d.kind == IMP_VEL_READ_FORCE) {      // Impose.cpp:175-177
        const Real f = dot(force, V3(d.v[0], d.v[1], d.v[2]));
        #pragma omp critical(impose_read)
        d.sumF += ((d.flags & 2) ? -1 : 1) * f;
      } else if (d.kind == IMP_READ_FORCE) {   // Impose.cpp:144-155
        #pragma omp critical(impose_read)
        {
          if (!(d.flags & 1)) { d.F += force; d.T += torque; }
          else { const Quat qc(d.ori.w, -d.ori.x, -d.ori.y, -d.ori.z); d.F += quatRotate(qc, force); d.T += quatRotate(qc, torque + cross(pos - V3(d.v[0], d.v[1], d.v[2]), force)); }
        }
      }
    }
  }
  void leapfrogRun() {
    const Real dt = P.dt; const Real damping = P.damping;
    imposePrepareStep();
    const int homoDeform = P.periodic ? cell.homoDeform : -1;
    M3 dGradV = subMM(cell.nextGradV, cell.gradV);
    M3 midGradV = scaleM(addMM(cell.gradV, cell.nextGradV), .5);
    M3 ImLL4hInv, IpLL4h, LmL; V3 deltaSpinVec;
    if (homoDeform == HOMO_GRADV2) {
      const M3& pprevL = cell.gradV; const M3& nnextL = cell.nextGradV;
      ImLL4hInv = inverseM(subMM(M3::Identity(), scaleM(scaleM(addMM(nnextL, pprevL), dt), 1 / 4.)));
      IpLL4h = addMM(M3::Identity(), scaleM(scaleM(addMM(nnextL, pprevL), dt), 1 / 4.));
      LmL = subMM(nnextL, pprevL);
      deltaSpinVec = -.5 * leviCivita(scaleM(subMM(pprevL, transposeM(pprevL)), .5)) + .5 * leviCivita(scaleM(subMM(nnextL, transposeM(nnextL)), .5));
    }
    const V3 g(P.gravity[0], P.gravity[1], P.gravity[2]);
    const bool hasGravity = (g != V3(0, 0, 0));
    const long n = (long)sph.size();
    // Leapfrog.cpp:130-132: no energy in the first step of a periodic scene with a velocity gradient (the mean-field velocity is not applied yet)
    bool gradVIsIdentity = true;
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) if (cell.gradV(a, b) != (a == b ? 1. : 0.)) gradVIsIdentity = false;
    const bool reallyTrackEnergy = trackEnergy && (!P.periodic || step > 0 || gradVIsIdentity);
    #pragma omp parallel for schedule(guided)
    for (long i = 0; i < n; i++) {
      Sphere& s = sph[i];
      if (s.clump >= 0) continue;      // integrated via the clump's master node (Leapfrog.cpp:150)
      const unsigned fl = s.flags;
      const bool blockedAll = (fl & DOF_ALL) == DOF_ALL, blockedNone = (fl & DOF_ALL) == 0;
      const bool damp = (damping != 0. && !(fl & DAMPING_SKIP));
      V3& f = s.force; V3& t = s.torque;
      if (reallyTrackEnergy) {
        // Leapfrog::doDampingDissipation (Leapfrog.cpp:53-65) and doGravityWork (:67-74), with the values at the start of the step
        if (damp && !(fl & ENERGY_SKIP)) {
          const V3 av(std::abs(s.vel.x), std::abs(s.vel.y), std::abs(s.vel.z)), af(std::abs(f.x), std::abs(f.y), std::abs(f.z));
          const V3 aw(std::abs(s.angVel.x), std::abs(s.angVel.y), std::abs(s.angVel.z)), at(std::abs(t.x), std::abs(t.y), std::abs(t.z));
          eadd(EN_DAMP, dot(av, af) * damping * dt + dot(aw, at) * damping * dt);
        }
        if (hasGravity && !(fl & GRAVITY_SKIP)) {
          Real gr = 0;
          if (blockedNone) gr = -dot(g, s.vel) * s.mass * dt;
          else { for (int k = 0; k < 3; k++) if (!(fl & (1u << k))) gr -= g[k] * s.vel[k] * s.mass * dt; }
          eadd(EN_GRAV, gr);
        }
      }
      V3 pprevFluctVel, pprevFluctAngVel, linAccel, angAccel;
      if (!blockedAll) {
        // computeAccel (Leapfrog.cpp:37-42)
        if (blockedNone) linAccel = f / s.mass;
        else { for (int k = 0; k < 3; k++) if (!(fl & (1u << k))) linAccel[k] += f[k] / s.mass; }
        if (P.periodic) { pprevFluctVel = cell.pprevFluctVel(s.pos, s.vel, dt); pprevFluctAngVel = cell.pprevFluctAngVel(s.angVel); }
        else { pprevFluctVel = s.vel; pprevFluctAngVel = s.angVel; }
        // nonviscDamp2nd (Leapfrog.cpp:33-35)
        if (damp) for (int k = 0; k < 3; k++) linAccel[k] *= 1 - damping * Sign(f[k] * (pprevFluctVel[k] + 0.5 * dt * linAccel[k]));
        if (homoDeform == HOMO_GRADV2) s.vel = mulMV(ImLL4hInv, (mulMV(LmL, s.pos) + mulMV(IpLL4h, s.vel)) + linAccel * dt);
        else s.vel += dt * linAccel;
        if (s.inertia != 0.) {
          if (blockedNone) angAccel = V3(t.x / s.inertia, t.y / s.inertia, t.z / s.inertia);
          else { for (int k = 0; k < 3; k++) if (!(fl & (1u << (k + 3)))) angAccel[k] += t[k] / s.inertia; }
          if (damp) for (int k = 0; k < 3; k++) angAccel[k] *= 1 - damping * Sign(t[k] * (pprevFluctAngVel[k] + 0.5 * dt * angAccel[k]));
          s.angVel += dt * angAccel;
          if (homoDeform == HOMO_GRADV2) s.angVel -= deltaSpinVec;
        }
        if (s.impose >= 0) imposeVelocity(s.impose, s.vel, s.angVel, s.pos);      // re-applied to cancel the effect of forces (Leapfrog.cpp:203-205)
      } else {
        if (homoDeform == HOMO_GRADV2) s.vel = mulMV(ImLL4hInv, mulMV(LmL, s.pos) + mulMV(IpLL4h, s.vel));
      }
      // applyPeriodicCorrections (Leapfrog.cpp:258-273)
      if (homoDeform >= 0 && homoDeform != HOMO_GRADV2) {
        if (homoDeform == HOMO_VEL || homoDeform == HOMO_VEL_2ND) {
          if (homoDeform == HOMO_VEL_2ND) s.vel += mulMV(midGradV, s.vel - (dt / 2.) * linAccel) * dt;
          s.vel += mulMV(dGradV, s.pos);
        } else if (homoDeform == HOMO_POS) s.pos += mulMV(cell.nextGradV, s.pos) * dt;
      }
      if (reallyTrackEnergy && !(fl & ENERGY_SKIP)) {
        // Leapfrog::doKineticEnergy (Leapfrog.cpp:76-98): on-step value from the mid-step fluctuation velocities; a fully blocked node passes zeros (:198)
        const V3 cfv = pprevFluctVel + (.5 * dt) * linAccel, cfw = pprevFluctAngVel + (.5 * dt) * angAccel;
        const Real Etrans = .5 * s.mass * sqNorm(cfv);
        Real Erot = 0.5 * dot(cfw, V3(s.inertia * cfw.x, s.inertia * cfw.y, s.inertia * cfw.z));
        if (std::isnan(Erot) || std::isinf(s.inertia)) Erot = 0;
        eadd(EN_KINETIC, Etrans + Erot);
      }
      s.pos += s.vel * dt;  // leapfrogTranslate (Leapfrog.cpp:275-278)
      // leapfrogSphericalRotate (Leapfrog.cpp:280-297)
      V3 axis = s.angVel;
      if (axis != V3(0, 0, 0)) {
        Real angle = norm(axis); axis = axis / angle;
        Real ha = 0.5 * (angle * dt);  // Eigen Quaternion(AngleAxis): ha=0.5*angle; w=cos(ha); vec=sin(ha)*axis
        Real sn = std::sin(ha);
        Quat q(std::cos(ha), sn * axis.x, sn * axis.y, sn * axis.z);
        s.ori = quatMul(q, s.ori);
      }
      quatNormalize(s.ori);
      if (s.impose >= 0) imposeReadForce(s.impose, s.pos, s.force, s.torque);      // Leapfrog.cpp:238
      if (P.reset) {
        s.force = (hasGravity && !(fl & GRAVITY_SKIP)) ? (s.mass * g) : V3();
        s.torque = V3();
        if (s.impose >= 0) imposeForce(s.impose, s.pos, s.force, s.torque);        // Leapfrog.cpp:244
      }
      if (s.impose >= 0) imposeVelocity(s.impose, s.vel, s.angVel, s.pos);         // Leapfrog.cpp:248
    }
    // clump master nodes (Leapfrog.cpp:148-255 with isClump): forces from the members, spherical or aspherical rotation, then the members follow
    if (!clumps.empty()) {
      if (trackEnergy) throw std::runtime_error("Scene.trackEnergy with clumps is not supported");
      #pragma omp parallel for schedule(guided)
      for (long ci = 0; ci < (long)clumps.size(); ci++) {
        Clump& c = clumps[ci];
        const unsigned fl = c.flags;
        const bool blockedAll = (fl & DOF_ALL) == DOF_ALL, blockedNone = (fl & DOF_ALL) == 0, blockedAllRot = (fl & 56u) == 56u;
        const bool damp = (damping != 0. && !(fl & DAMPING_SKIP));
        if (!blockedAll) {      // ClumpData::forceTorqueFromMembers (Clump.cpp:272-280): accumulates to the node's own force/torque (gravity from the last reset)
          for (int m : c.members) { const Sphere& ms = sph[m]; c.force += ms.force; c.torque += ms.torque + cross(ms.pos - c.pos, ms.force); }
        }
        V3& f = c.force; V3& t = c.torque;
        const bool aspherical = !(c.inertia.x == c.inertia.y && c.inertia.y == c.inertia.z), useAspherical = aspherical && !blockedAllRot;      // Particle.hpp:275-277
        V3 linAccel, angAccel;
        if (!blockedAll) {
          if (blockedNone) linAccel = f / c.mass;
          else { for (int k = 0; k < 3; k++) if (!(fl & (1u << k))) linAccel[k] += f[k] / c.mass; }
          // fluctuation velocities (Leapfrog.cpp:181-184): the master node of a clump is a node like any other, in a periodic cell too
          V3 pfv = c.vel, pfw = c.angVel;
          if (P.periodic) { pfv = cell.pprevFluctVel(c.pos, c.vel, dt); pfw = cell.pprevFluctAngVel(c.angVel); }
          if (damp) for (int k = 0; k < 3; k++) linAccel[k] *= 1 - damping * Sign(f[k] * (pfv[k] + 0.5 * dt * linAccel[k]));
          if (homoDeform == HOMO_GRADV2) c.vel = mulMV(ImLL4hInv, (mulMV(LmL, c.pos) + mulMV(IpLL4h, c.vel)) + linAccel * dt);
          else c.vel += dt * linAccel;
          if (c.inertia != V3(0, 0, 0)) {
            if (!useAspherical) {
              if (blockedNone) angAccel = V3(t.x / c.inertia.x, t.y / c.inertia.y, t.z / c.inertia.z);
              else { for (int k = 0; k < 3; k++) if (!(fl & (1u << (k + 3)))) angAccel[k] += t[k] / c.inertia[k]; }
              if (damp) for (int k = 0; k < 3; k++) angAccel[k] *= 1 - damping * Sign(t[k] * (pfw[k] + 0.5 * dt * angAccel[k]));
              c.angVel += dt * angAccel;
              if (homoDeform == HOMO_GRADV2) c.angVel -= deltaSpinVec;
            } else {
              for (int k = 0; k < 3; k++) if (fl & (1u << (k + 3))) t[k] = 0;                                    // block DOFs here (Leapfrog.cpp:188)
              if (damp) for (int k = 0; k < 3; k++) t[k] *= 1 - damping * Sign(t[k] * pfw[k]);                   // nonviscDamp1st (Leapfrog.cpp:28-30)
            }
          }
          if (c.impose >= 0) imposeVelocity(c.impose, c.vel, c.angVel, c.pos, &c.ori);      // Leapfrog.cpp:203-205
        } else if (homoDeform == HOMO_GRADV2) c.vel = mulMV(ImLL4hInv, mulMV(LmL, c.pos) + mulMV(IpLL4h, c.vel));      // fixed node (Leapfrog.cpp:206-210)
        if (homoDeform >= 0 && homoDeform != HOMO_GRADV2) {      // applyPeriodicCorrections (Leapfrog.cpp:258-273)
          if (homoDeform == HOMO_VEL || homoDeform == HOMO_VEL_2ND) {
            if (homoDeform == HOMO_VEL_2ND) c.vel += mulMV(midGradV, c.vel - (dt / 2.) * linAccel) * dt;
            c.vel += mulMV(dGradV, c.pos);
          } else if (homoDeform == HOMO_POS) c.pos += mulMV(cell.nextGradV, c.pos) * dt;
        }
        c.pos += c.vel * dt;
        if (!useAspherical) {      // leapfrogSphericalRotate (Leapfrog.cpp:280-297)
          V3 axis = c.angVel;
          if (axis != V3(0, 0, 0)) {
            Real angle = norm(axis); axis = axis / angle;
            Real ha = 0.5 * (angle * dt); Real sn = std::sin(ha);
            c.ori = quatMul(Quat(std::cos(ha), sn * axis.x, sn * axis.y, sn * axis.z), c.ori);
          }
          quatNormalize(c.ori);
        } else {      // leapfrogAsphericalRotate (Leapfrog.cpp:299-321)
          if (c.inertia == V3(0, 0, 0)) throw std::runtime_error("Leapfrog::run: inertia==(0,0,0), but the node wants to use aspherical integrator");
          const Quat qc(c.ori.w, -c.ori.x, -c.ori.y, -c.ori.z);
          if (std::isnan(std::min(c.angMom.x, std::min(c.angMom.y, c.angMom.z))) || c.angMom.x != c.angMom.x || c.angMom.y != c.angMom.y || c.angMom.z != c.angMom.z) {
            const V3 wl = quatRotate(qc, c.angVel); c.angMom = V3(c.inertia.x * wl.x, c.inertia.y * wl.y, c.inertia.z * wl.z);
          }
          const M3 A = quatToMat(qc);                                             // rotation matrix from global to local r.f.
          const V3 l_n = c.angMom + dt / 2 * t;
          const V3 l_b_n = mulMV(A, l_n);
          const V3 w_b_n(l_b_n.x / c.inertia.x, l_b_n.y / c.inertia.y, l_b_n.z / c.inertia.z);
          const Quat dotQ_n = dotQ(w_b_n, c.ori);
          const Quat Q_half(c.ori.w + dt / 2 * dotQ_n.w, c.ori.x + dt / 2 * dotQ_n.x, c.ori.y + dt / 2 * dotQ_n.y, c.ori.z + dt / 2 * dotQ_n.z);
          c.angMom += dt * t;
          const V3 l_b_half = mulMV(A, c.angMom);
          const V3 w_b_half(l_b_half.x / c.inertia.x, l_b_half.y / c.inertia.y, l_b_half.z / c.inertia.z);
          const Quat dotQ_half = dotQ(w_b_half, Q_half);
          c.ori = Quat(c.ori.w + dt * dotQ_half.w, c.ori.x + dt * dotQ_half.x, c.ori.y + dt * dotQ_half.y, c.ori.z + dt * dotQ_half.z);
          c.angVel = quatRotate(c.ori, w_b_half);
          quatNormalize(c.ori);
        }
        if (P.reset) {      // gravity on the clump itself, not on the members; Impose::force on top (Leapfrog.cpp:240-245)
          c.force = (hasGravity && !(fl & GRAVITY_SKIP)) ? (c.mass * g) : V3(); c.torque = V3();
          if (c.impose >= 0) imposeForce(c.impose, c.pos, c.force, c.torque);
        }
        if (c.impose >= 0) imposeVelocity(c.impose, c.vel, c.angVel, c.pos, &c.ori);      // Leapfrog.cpp:248, before applyToMembers (:252)
        // ClumpData::applyToMembers (Clump.cpp:293-307)
        for (size_t k = 0; k < c.members.size(); k++) {
          Sphere& ms = sph[c.members[k]];
          ms.pos = c.pos + quatRotate(c.ori, c.relPos[k]);
          ms.ori = quatMul(c.ori, c.relOri[k]);
          ms.vel = c.vel + cross(c.angVel, ms.pos - c.pos);
          ms.angVel = c.angVel;
          if (P.reset) { ms.force = V3(); ms.torque = V3(); }
        }
      }
    }
    // boundary nodes that are in DemField.nodes (walls added with a velocity): mass 0 and all DOFs blocked => Leapfrog's "fixed particle" branch
    // (Leapfrog.cpp:195-199): no acceleration, leapfrogTranslate pos += vel*dt (:275-278), force/torque reset to mass*gravity = 0 (:240-245)
    // clump master nodes of kinematic facet groups: all DOFs blocked => Leapfrog's fixed-particle branch: leapfrogTranslate + leapfrogSphericalRotate with the
    // prescribed velocities (Leapfrog.cpp:195-199,275-297), then ClumpData::applyToMembers (Clump.cpp:293-307, called from Leapfrog.cpp:252)
    for (FacetGroup& G : groups) {
      if (P.periodic) {      // a fully blocked node in a deforming cell (Leapfrog.cpp:206-215 fixed branch, applyPeriodicCorrections :258-273), no acceleration
        if (homoDeform == HOMO_GRADV2) G.vel = mulMV(ImLL4hInv, mulMV(LmL, G.pos) + mulMV(IpLL4h, G.vel));
        else if (homoDeform == HOMO_VEL || homoDeform == HOMO_VEL_2ND) {
          if (homoDeform == HOMO_VEL_2ND) G.vel += mulMV(midGradV, G.vel - (dt / 2.) * V3()) * dt;
          G.vel += mulMV(dGradV, G.pos);
        } else if (homoDeform == HOMO_POS) G.pos += mulMV(cell.nextGradV, G.pos) * dt;
      }
      G.pos += G.vel * dt;
      V3 axis = G.angVel;
      if (axis != V3(0, 0, 0)) {
        Real angle = norm(axis); axis = axis / angle;
        Real ha = 0.5 * (angle * dt);
        Real sn = std::sin(ha);
        Quat q(std::cos(ha), sn * axis.x, sn * axis.y, sn * axis.z);
        G.ori = quatMul(q, G.ori);
      }
      quatNormalize(G.ori);
      if (G.impose >= 0) imposeVelocity(G.impose, G.vel, G.angVel, G.pos, &G.ori);      // before applyToMembers (Leapfrog.cpp:248-252)
      for (size_t k = 0; k < G.facets.size(); k++) {
        Facet& f = facets[G.facets[k]];
        for (int i = 0; i < 3; i++) {
          f.v[i] = G.pos + quatRotate(G.ori, G.relPos[3 * k + i]);
          f.vvel[i] = G.vel + cross(G.angVel, f.v[i] - G.pos);
        }
      }
    }
    // facet nodes that carry an imposition are in DemField.nodes (guessMoving: || impose): fixed-node branch = leapfrogTranslate with the node's velocity, then
    // Impose::velocity at the end of the node's step sets the velocity of the next one (Leapfrog.cpp:195-199,248)
    for (Facet& f : facets) {
      if (f.impose < 0) continue;
      for (int i = 0; i < 3; i++) {
        if (P.periodic) {      // the cell's corrections of the fixed branch (Leapfrog.cpp:206-215,258-273)
          if (homoDeform == HOMO_GRADV2) f.vvel[i] = mulMV(ImLL4hInv, mulMV(LmL, f.v[i]) + mulMV(IpLL4h, f.vvel[i]));
          else if (homoDeform == HOMO_VEL || homoDeform == HOMO_VEL_2ND) {
            if (homoDeform == HOMO_VEL_2ND) f.vvel[i] += mulMV(midGradV, f.vvel[i] - (dt / 2.) * V3()) * dt;
            f.vvel[i] += mulMV(dGradV, f.v[i]);
          } else if (homoDeform == HOMO_POS) f.v[i] += mulMV(cell.nextGradV, f.v[i]) * dt;
        }
        f.v[i] += f.vvel[i] * dt; V3 dummyW; imposeVelocity(f.impose, f.vvel[i], dummyW, f.v[i]);
      }
    }
    for (Wall& w : walls) {
      if (!w.moving) continue;
      if (P.periodic) {      // a fully blocked node in a deforming cell (Leapfrog.cpp:206-215 fixed branch, applyPeriodicCorrections :258-273), no acceleration
        if (w.impose >= 0) throw std::runtime_error("impositions on walls in a periodic cell are not supported");
        if (homoDeform == HOMO_GRADV2) w.vel = mulMV(ImLL4hInv, mulMV(LmL, w.pos) + mulMV(IpLL4h, w.vel));
        else if (homoDeform == HOMO_VEL || homoDeform == HOMO_VEL_2ND) {
          if (homoDeform == HOMO_VEL_2ND) w.vel += mulMV(midGradV, w.vel - (dt / 2.) * V3()) * dt;
          w.vel += mulMV(dGradV, w.pos);
        } else if (homoDeform == HOMO_POS) w.pos += mulMV(cell.nextGradV, w.pos) * dt;
      }
      w.pos += w.vel * dt;
      if (P.reset) { w.force = V3(); w.torque = V3(); }
      if (w.impose >= 0) { V3 dummyW; imposeVelocity(w.impose, w.vel, dummyW, w.pos); }      // fully blocked node: only the call at the end of the node's step (Leapfrog.cpp:248)
    }
  }

  // ------------------------------------------------------------------------------------------
  // DynDt (DynDt.cpp:16-113) and pWaveDt (Funcs.cpp:151-175)
  // ------------------------------------------------------------------------------------------
  Real critDtStiffness() {
    // DynDt::nodalStiffAdd walks p->contacts, a std::map keyed by the OTHER particle's id (Particle.hpp:36) => ascending other id
    struct Inc { int other; const Contact* C; int side; };
    std::vector<std::vector<Inc>> inc(par.size());
    for (auto& kv : con) { const Contact& C = kv.second; if (!C.isReal()) continue; inc[C.idA].push_back({C.idB, &C, 0}); inc[C.idB].push_back({C.idA, &C, 1}); }
    Real ret = Inf;
    for (size_t id = 0; id < par.size(); id++) {
      const PRef& p = par[id];
      if (p.kind != SHAPE_SPHERE) continue;  // walls/facets are not in dem.nodes
      const Sphere& s = sph[p.idx];
      std::sort(inc[id].begin(), inc[id].end(), [](const Inc& a, const Inc& b) { return a.other < b.other; });
      if (s.clump >= 0) continue;      // a clumped node is not in DemField.nodes: its contacts count for the clump's master node (below)
      if ((s.flags & DOF_ALL) == DOF_ALL) continue;
      V3 ktrans, krot;
      for (const Inc& e : inc[id]) {
        const Contact& C = *e.C;
        V3 n = quatRotate(C.ori, V3(1, 0, 0));
        V3 n2(n.x * n.x, n.y * n.y, n.z * n.z);
        ktrans += n2 * (C.kn - C.kt) + V3(C.kt, C.kt, C.kt);
        krot += pow2(C.lens[e.side]) * C.kt * V3(n2[1] + n2[2], n2[2] + n2[0], n2[0] + n2[1]);
      }
      Real r = Inf;
      for (int k = 0; k < 3; k++) if (ktrans[k] != 0 && s.mass > 0. && !(s.flags & (1u << k))) r = std::min(r, s.mass / std::abs(ktrans[k]));
      for (int k = 0; k < 3; k++) if (krot[k] != 0 && s.inertia > 0. && !(s.flags & (1u << (k + 3)))) r = std::min(r, s.inertia / std::abs(krot[k]));
      ret = std::min(ret, 2 * r);
    }
    // clump master nodes: stiffnesses of the contacts of all member particles, member by member (DynDt.cpp:48-52), against the clump's mass and principal inertia
    for (const Clump& c : clumps) {
      if ((c.flags & DOF_ALL) == DOF_ALL) continue;
      V3 ktrans, krot;
      for (int m : c.members) {
        for (const Inc& e : inc[sph[m].id]) {
          const Contact& C = *e.C;
          V3 n = quatRotate(C.ori, V3(1, 0, 0));
          V3 n2(n.x * n.x, n.y * n.y, n.z * n.z);
          ktrans += n2 * (C.kn - C.kt) + V3(C.kt, C.kt, C.kt);
          krot += pow2(C.lens[e.side]) * C.kt * V3(n2[1] + n2[2], n2[2] + n2[0], n2[0] + n2[1]);
        }
      }
      Real r = Inf;
      for (int k = 0; k < 3; k++) if (ktrans[k] != 0 && c.mass > 0. && !(c.flags & (1u << k))) r = std::min(r, c.mass / std::abs(ktrans[k]));
      for (int k = 0; k < 3; k++) if (krot[k] != 0 && c.inertia[k] > 0. && !(c.flags & (1u << (k + 3)))) r = std::min(r, c.inertia[k] / std::abs(krot[k]));
      ret = std::min(ret, 2 * r);
    }
    return std::sqrt(ret);
  }
  static Real pown(Real x, unsigned n) {  // Math.hpp:161-176
    unsigned d = n >> 1, r = n & 1;
    Real x_2_d = d == 0 ? 1 : pown(x * x, d);
    Real x_r = r == 0 ? 1 : x;
    return x_2_d * x_r;
  }
  void dynDtRun() {  // PeriodicEngine::isActivated stepPeriod logic (Engine.cpp:170-195) + DynDt::run
    if (P.dynDtPeriod <= 0) return;
    // defaults stepModulo=true, initRun=true (Engine.hpp:150,153): runs at step 0 and whenever step%stepPeriod==0
    bool initNow = (dynDtStepLast < 0);
    if (initNow) dynDtStepLast = step;
    if (!((step - dynDtStepLast >= P.dynDtPeriod) || (step % P.dynDtPeriod == 0))) return;
    dynDtStepPrev = dynDtStepLast; dynDtStepLast = step;
    Real crDt = critDtStiffness();
    if (std::isinf(crDt)) return;
    int nSteps = (int)(step - dynDtStepPrev);
    Real maxNewDt = P.dt * pown(1. + P.maxRelInc, (unsigned)nSteps);
    Real nDt = std::min(crDt * P.dtSafety, maxNewDt);
    if (nDt != P.dt) nextDt = nDt;
  }
  Real pWaveDt() const {
    Real dt = Inf;
    for (auto& s : sph) { const Material& m = mats[s.mat]; dt = std::min(dt, s.radius / (1. * std::sqrt(m.young / m.density))); }
    return dt;
  }

  // ------------------------------------------------------------------------------------------
  // Scene::doOneStep (Scene.cpp:366-432)
  // ------------------------------------------------------------------------------------------
  void oneStep() {
    if (std::isnan(P.dt) || std::isinf(P.dt) || P.dt < 0) {
      Real crDt = pWaveDt();
      if (std::isinf(crDt)) throw std::runtime_error("Failed to obtain meaningful dt from engines and fields automatically");
      P.dt = P.dtSafety * crDt;
    }
    if (P.periodic) cell.integrateAndUpdate(P.dt);
    if (trackEnergy) { energy[EN_KINETIC] = 0; energy[EN_ELAST] = 0; }   // energy->resetResettables() (Scene.cpp:416)
    leapfrogRun();
    colliderRun();
    contactLoopRun();
    if (P.periodic && evalStress) {   // ContactLoop.cpp:243-254 (after the contact loop, before the deterministic force application: same state)
      const M3 st = stressNow(); const Real vol = detM(cell.hSize);
      if (trackEnergy) {
        const M3 mid = scaleM(addMM(st, prevStress), .5);
        const Real midVol = !std::isnan(prevVol) ? .5 * (prevVol + vol) : vol;
        const M3 gs = mulMM(cell.gradV, mid);
        gradVWork += -trace3(gs(0, 0), gs(1, 1), gs(2, 2)) * P.dt * midVol;
      }
      prevVol = vol; prevStress = st;
    }
    dynDtRun();
    if (P.periodic) cell.setNextGradV();
    step++; time += P.dt;
    if (!std::isnan(nextDt)) { P.dt = nextDt; nextDt = NaN; }
  }

  // ContactLoop.stress (ContactLoop.cpp:116,218-231,243-244): sum over real contacts of F (x) branch / V with F = node.ori * phys.force (global force as applied to
  // pA), branch = Contact::dPos (Contact.cpp:54-59), V = Cell::getVolume (Cell.hpp:121).  The reference adds the terms in its (thread-dependent) loop order; here:
  // canonical contact order.
  M3 stressNow() const {
    M3 st; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) st(i, j) = 0;
    for (auto& kv : con) {
      const Contact& C = kv.second; if (!C.isReal()) continue;
      const PRef& a = par[C.idA]; const PRef& b = par[C.idB];
      if (a.kind == SHAPE_FACET || b.kind == SHAPE_FACET) throw std::runtime_error("ContactLoop.trackWork not allowed with multi-nodal particles in contact");
      const V3& pa = a.kind == SHAPE_SPHERE ? sph[a.idx].pos : walls[a.idx].pos;
      const V3& pb = b.kind == SHAPE_SPHERE ? sph[b.idx].pos : walls[b.idx].pos;
      V3 branch = pb - pa;
      if (C.cellDist[0] != 0 || C.cellDist[1] != 0 || C.cellDist[2] != 0) branch = branch + cell.intrShiftPos(C.cellDist);
      const V3 F = quatRotate(C.ori, C.force);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) st(i, j) += F[i] * branch[j];
    }
    const Real vol = detM(cell.hSize);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) st(i, j) = st(i, j) / vol;
    return st;
  }

  // observables (Particle.cpp:247-265, Funcs.cpp:84-113)
  Real kineticEnergy() const {
    Real ek = 0;
    for (auto& s : sph) {
      V3 fv = P.periodic ? cell.pprevFluctVel(s.pos, s.vel, P.dt) : s.vel;
      V3 fw = P.periodic ? cell.pprevFluctAngVel(s.angVel) : s.angVel;
      ek += .5 * (s.mass * dot(fv, fv));
      if (!std::isinf(s.inertia)) ek += .5 * s.inertia * dot(fw, fw);
    }
    return ek;
  }
};
}  // namespace orc

// ===============================================================================================
// C interface (ctypes)
// ===============================================================================================
using namespace orc;
extern "C" {
struct orc_params {  // keep in sync with tests/_harness.py OrcParams
  double dt, dtSafety, gravity[3], distFactor, damping, verletDist, wallHalfThick, relRollStiff, relTwistStiff, rollTanPhi, poisson, en, maxRelInc;
  unsigned loneMask;
  int deterministic, reset, noRatch, iniLensTouch, trsfRenorm, applyForces, law, noSlip, noBreak, noFrict, noAttraction, dynDtPeriod, periodic;
};
#define ORC_TRY(h) try {
#define ORC_CATCH(h) } catch (std::exception & e) { ((Scene*)h)->err = e.what(); return -1; } return 0;

void* orc_create() { return new Scene(); }
void orc_destroy(void* h) { delete (Scene*)h; }
const char* orc_last_error(void* h) { return ((Scene*)h)->err.c_str(); }
void orc_default_params(orc_params* p) {
  Params d; memset(p, 0, sizeof(*p));
  p->dt = d.dt; p->dtSafety = d.dtSafety; p->distFactor = d.distFactor; p->damping = d.damping; p->verletDist = d.verletDist; p->wallHalfThick = 0;
  p->relRollStiff = 0; p->relTwistStiff = 0; p->rollTanPhi = 0; p->poisson = d.poisson; p->en = d.en; p->maxRelInc = d.maxRelInc; p->loneMask = d.loneMask;
  p->deterministic = 1; p->reset = 1; p->noRatch = 0; p->iniLensTouch = 1; p->trsfRenorm = 100; p->applyForces = 1; p->law = 0; p->noSlip = 0; p->noBreak = 0; p->noFrict = 0;
  p->noAttraction = 1; p->dynDtPeriod = 100; p->periodic = 0;
}
void orc_set_params(void* h, const orc_params* p) {
  Params& P = ((Scene*)h)->P;
  P.dt = p->dt; P.dtSafety = p->dtSafety; for (int k = 0; k < 3; k++) P.gravity[k] = p->gravity[k]; P.distFactor = p->distFactor; P.damping = p->damping;
  P.verletDist = p->verletDist; P.wallHalfThick = p->wallHalfThick; P.relRollStiff = p->relRollStiff; P.relTwistStiff = p->relTwistStiff; P.rollTanPhi = p->rollTanPhi;
  P.poisson = p->poisson; P.en = p->en; P.maxRelInc = p->maxRelInc; P.loneMask = p->loneMask; P.deterministic = p->deterministic; P.reset = p->reset; P.noRatch = p->noRatch;
  P.iniLensTouch = p->iniLensTouch; P.trsfRenorm = p->trsfRenorm; P.applyForces = p->applyForces; P.law = p->law; P.noSlip = p->noSlip; P.noBreak = p->noBreak; P.noFrict = p->noFrict;
  P.noAttraction = p->noAttraction; P.dynDtPeriod = p->dynDtPeriod; P.periodic = p->periodic;
}
int orc_add_material(void* h, int kind, double density, double young, double ktDivKn, double tanPhi, double surfEnergy, double alpha) {
  Scene* S = (Scene*)h; S->mats.push_back(Material{kind, density, young, ktDivKn, tanPhi, surfEnergy, alpha}); return (int)S->mats.size() - 1;
}
// bulk add; ids are assigned sequentially in call order (like S.dem.par.add). vel/angVel/ori may be NULL.
int orc_add_spheres(void* h, int n, const double* pos, const double* radius, const double* mass, const double* inertia, const int* mat,
                    const unsigned* mask, const unsigned* flags, const double* vel, const double* angVel, const double* ori /*xyzw*/) {
  Scene* S = (Scene*)h; int first = -1; S->lastAdded.clear();
  for (int i = 0; i < n; i++) {
    // ParticleContainer::findFreeId (ParticleContainer.cpp:23-36): ids freed by remove() are handed out again, oldest first; ids beyond the (shrunk, :152-155)
    // container are dropped.  The restatement never shrinks its arrays, so the container size is 1 + the largest id still alive.
    int size = 0;
    for (int q = (int)S->par.size() - 1; q >= 0; q--) { const PRef& pr = S->par[q]; if (pr.kind != SHAPE_SPHERE || !S->sph[pr.idx].dead) { size = q + 1; break; } }
    int id = size;
    while (!S->freeIds.empty()) { const int f = S->freeIds.front(); S->freeIds.erase(S->freeIds.begin()); if (f <= size) { id = f; break; } }
    if (first < 0) first = id;
    S->lastAdded.push_back(id);
    Sphere s; s.id = id; s.pos = V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]); s.radius = radius[i]; s.mass = mass[i]; s.inertia = inertia[i];
    s.mat = mat[i]; s.mask = mask ? mask[i] : 5u; s.flags = flags ? flags[i] : 0u;
    if (vel) s.vel = V3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
    if (angVel) s.angVel = V3(angVel[3 * i], angVel[3 * i + 1], angVel[3 * i + 2]);
    if (ori) s.ori = Quat(ori[4 * i + 3], ori[4 * i], ori[4 * i + 1], ori[4 * i + 2]);
    if (id < (int)S->par.size()) S->sph[S->par[id].idx] = s;      // a removed particle's id: the slot is taken over (nothing of the old particle survives)
    else { S->par.push_back(PRef{SHAPE_SPHERE, (int)S->sph.size()}); S->sph.push_back(s); }
  }
  S->conDirty = true; return first;
}
int orc_add_wall(void* h, const double* pos, int axis, int sense, int mat, unsigned mask, const double* vel) {
  Scene* S = (Scene*)h; Wall w; w.id = (int)S->par.size(); w.pos = V3(pos[0], pos[1], pos[2]); w.axis = axis; w.sense = sense; w.mat = mat; w.mask = mask;
  if (vel) w.vel = V3(vel[0], vel[1], vel[2]);
  w.moving = (w.vel != V3(0, 0, 0));
  S->par.push_back(PRef{SHAPE_WALL, (int)S->walls.size()}); S->walls.push_back(w); S->conDirty = true; return w.id;
}
int orc_add_facet(void* h, const double* v /*9*/, double halfThick, const double* fakeVel, const double* n21lim, const double* vvel /*9 or NULL*/, int mat, unsigned mask) {
  Scene* S = (Scene*)h; Facet f; f.id = (int)S->par.size();
  for (int i = 0; i < 3; i++) { f.v[i] = V3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); if (vvel) f.vvel[i] = V3(vvel[3 * i], vvel[3 * i + 1], vvel[3 * i + 2]); f.n21lim[i] = n21lim ? n21lim[i] : NaN; }
  f.halfThick = halfThick; if (fakeVel) f.fakeVel = V3(fakeVel[0], fakeVel[1], fakeVel[2]); f.mat = mat; f.mask = mask;
  S->par.push_back(PRef{SHAPE_FACET, (int)S->facets.size()}); S->facets.push_back(f); S->conDirty = true; return f.id;
}
// facetIdx: indices into the facets in the order they were added
int orc_add_facet_group(void* h, const double* pos, const double* ori /*x,y,z,w*/, const double* vel, const double* angVel, int n, const int* facetIdx) {
  Scene* S = (Scene*)h;
  ORC_TRY(h)
  FacetGroup G; G.pos = V3(pos[0], pos[1], pos[2]); G.ori = Quat(ori[3], ori[0], ori[1], ori[2]); G.vel = V3(vel[0], vel[1], vel[2]); G.angVel = V3(angVel[0], angVel[1], angVel[2]);
  const Quat qc(G.ori.w, -G.ori.x, -G.ori.y, -G.ori.z);   // conjugate
  for (int k = 0; k < n; k++) {
    if (facetIdx[k] < 0 || facetIdx[k] >= (int)S->facets.size()) throw std::runtime_error("facet group: bad facet index");
    G.facets.push_back(facetIdx[k]);
    for (int i = 0; i < 3; i++) G.relPos.push_back(quatRotate(qc, S->facets[facetIdx[k]].v[i] - G.pos));   // Clump.cpp:252
  }
  // members follow the master from the start (ClumpData::applyToMembers runs in every Leapfrog): positions as given, velocities from the master
  for (size_t k = 0; k < G.facets.size(); k++) { Facet& f = S->facets[G.facets[k]]; for (int i = 0; i < 3; i++) f.vvel[i] = G.vel + cross(G.angVel, f.v[i] - G.pos); }
  S->groups.push_back(G);
  ORC_CATCH(h)
}
// ParticleContainer.addClumped(spheres) with the clump's mass properties given by the caller (ClumpData::makeClump computes them from the members: Clump.cpp:187-261;
// here they are inputs, so that oracle and engine start from the same numbers): members become slaves (DemData::setClumped), relPos / relOri in the clump's frame
int orc_add_clump(void* h, int n, const int* sphereIds, const double* pos, const double* ori /*x,y,z,w*/, const double* vel, const double* angVel, double mass, const double* inertia,
                  unsigned flags) {
  Scene* S = (Scene*)h;
  ORC_TRY(h)
  Clump c; c.pos = V3(pos[0], pos[1], pos[2]); c.ori = Quat(ori[3], ori[0], ori[1], ori[2]); c.vel = V3(vel[0], vel[1], vel[2]); c.angVel = V3(angVel[0], angVel[1], angVel[2]);
  c.mass = mass; c.inertia = V3(inertia[0], inertia[1], inertia[2]); c.flags = flags;
  const Quat qc(c.ori.w, -c.ori.x, -c.ori.y, -c.ori.z);
  for (int k = 0; k < n; k++) {
    const int id = sphereIds[k];
    if (id < 0 || id >= (int)S->par.size() || S->par[id].kind != SHAPE_SPHERE) throw std::runtime_error("clump: member #" + std::to_string(id) + " is not a sphere");
    Sphere& ms = S->sph[S->par[id].idx];
    if (ms.clump >= 0) throw std::runtime_error("Node of particle #" + std::to_string(id) + ": already clumped.");
    ms.clump = (int)S->clumps.size();
    c.members.push_back(S->par[id].idx);
    c.relPos.push_back(quatRotate(qc, ms.pos - c.pos));      // Clump.cpp:252-253
    c.relOri.push_back(quatMul(qc, ms.ori));
  }
  // the members follow the master from the start
  for (size_t k = 0; k < c.members.size(); k++) { Sphere& ms = S->sph[c.members[k]]; ms.vel = c.vel + cross(c.angVel, ms.pos - c.pos); ms.angVel = c.angVel; }
  S->clumps.push_back(c); S->conDirty = true;
  ORC_CATCH(h)
}
// Impose: descriptors, sets of descriptors (CombinedImpose), attachment to nodes.  v[16]: HARMONIC freq, amp, dir(3, normalized), t0 [flags&1: perpFree] |
// ALIGNED_HARMONIC freqs(3), amps(3) | CONSTANT_FORCE F(3) | RADIAL_FORCE A(3), B(3), F | LOCAL6 values(6), whats(6) + ori | VEL_READ_FORCE dir(3), vel [flags&1 latBlock, &2 invF] |
// READ_FORCE node pos(3) + ori [flags&1: has node] | VARIABLE_VEL3D t0 + ori + times/vals(/vals2 = angVels) [flags&1 diff, &2 wrap] | VARIABLE_ALIGNED_ROTATION axis + times/vals(1 per time) [flags&1 wrap]
int orc_add_impose(void* h, int kind, int flags, const double* v, const double* ori /*x,y,z,w or NULL*/, int nTimes, const double* times, const double* vals, const double* vals2) {
  Scene* S = (Scene*)h;
  ImposeDesc d; d.kind = kind; d.flags = flags; for (int k = 0; k < 16; k++) d.v[k] = v ? v[k] : 0.;
  if (ori) d.ori = Quat(ori[3], ori[0], ori[1], ori[2]);
  if (kind == IMP_VEL_READ_FORCE) d.dist = d.v[4];      // VelocityAndReadForce.dist is an attribute: a scene may start with a value
  const int w = kind == IMP_VARIABLE_ALIGNED_ROTATION ? 1 : 3;
  if (nTimes > 0) { d.times.assign(times, times + nTimes); d.vals.assign(vals, vals + (size_t)w * nTimes); if (vals2) d.vals2.assign(vals2, vals2 + (size_t)(kind == IMP_INTERPOLATED_MOTION ? 4 : w) * nTimes); }
  S->imposes.push_back(d);
  return (int)S->imposes.size() - 1;
}
int orc_add_impose_set(void* h, int n, const int* ops) { Scene* S = (Scene*)h; S->imposeSets.push_back(std::vector<int>(ops, ops + n)); return (int)S->imposeSets.size() - 1; }
// kind of node: 0 sphere (particle id), 1 wall (particle id), 2 facet group (index), 3 the three nodes of a facet (particle id), 4 master node of a clump of spheres (index)
int orc_attach_impose(void* h, int nodeKind, int which, int set) {
  Scene* S = (Scene*)h;
  ORC_TRY(h)
  if (set < 0 || set >= (int)S->imposeSets.size()) throw std::runtime_error("impose set out of range");
  bool readF = false; for (int ix : S->imposeSets[set]) { S->imposes[ix].nNodes++; if (S->imposes[ix].what() & IMPW_READ_FORCE) readF = true; }
  if (nodeKind == 0) { if (which < 0 || which >= (int)S->par.size() || S->par[which].kind != SHAPE_SPHERE) throw std::runtime_error("impose: not a sphere"); S->sph[S->par[which].idx].impose = set; }
  else if (nodeKind == 1) { if (which < 0 || which >= (int)S->par.size() || S->par[which].kind != SHAPE_WALL) throw std::runtime_error("impose: not a wall"); if (readF) throw std::runtime_error("readForce on walls is not part of this path"); Wall& w = S->walls[S->par[which].idx]; w.impose = set; w.moving = true; /* DemData::guessMoving: || impose (Particle.cpp:639) */ }
  else if (nodeKind == 3) { if (which < 0 || which >= (int)S->par.size() || S->par[which].kind != SHAPE_FACET) throw std::runtime_error("impose: not a facet"); if (readF) throw std::runtime_error("readForce on facet nodes is not part of this path"); for (int ix : S->imposeSets[set]) S->imposes[ix].nNodes += 2; S->facets[S->par[which].idx].impose = set; }
  else if (nodeKind == 4) { if (which < 0 || which >= (int)S->clumps.size()) throw std::runtime_error("impose: no such clump"); if (readF) throw std::runtime_error("readForce on clump master nodes is not part of this path"); S->clumps[which].impose = set; }
  else { if (which < 0 || which >= (int)S->groups.size()) throw std::runtime_error("impose: no such facet group"); if (readF) throw std::runtime_error("readForce on facet groups is not part of this path"); S->groups[which].impose = set; }
  ORC_CATCH(h)
}
void orc_get_impose_state(void* h, int ix, double* out /*sumF, dist, F(3), T(3)*/) {
  const ImposeDesc& d = ((Scene*)h)->imposes[ix];
  out[0] = d.sumF; out[1] = d.dist; for (int k = 0; k < 3; k++) { out[2 + k] = d.F[k]; out[5 + k] = d.T[k]; }
}
void orc_get_clumps(void* h, double* pos, double* ori, double* vel, double* angVel) {
  Scene* S = (Scene*)h;
  for (size_t i = 0; i < S->clumps.size(); i++) {
    const Clump& c = S->clumps[i];
    for (int k = 0; k < 3; k++) { pos[3 * i + k] = c.pos[k]; vel[3 * i + k] = c.vel[k]; angVel[3 * i + k] = c.angVel[k]; }
    ori[4 * i] = c.ori.x; ori[4 * i + 1] = c.ori.y; ori[4 * i + 2] = c.ori.z; ori[4 * i + 3] = c.ori.w;
  }
}
void orc_get_facets(void* h, double* verts /*9 per facet*/) {
  Scene* S = (Scene*)h;
  for (size_t f = 0; f < S->facets.size(); f++) for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) verts[9 * f + 3 * i + k] = S->facets[f].v[i][k];
}
void orc_set_cell(void* h, const double* hSize /*row-major 9*/, const double* gradV, const double* nextGradV, int homoDeform) {
  Scene* S = (Scene*)h; Cell& c = S->cell;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { c.hSize(i, j) = hSize[3 * i + j]; c.gradV(i, j) = gradV ? gradV[3 * i + j] : 0; c.nextGradV(i, j) = nextGradV ? nextGradV[3 * i + j] : 0; }
  c.pprevHsize = c.hSize; c.trsf = M3::Identity(); c.homoDeform = homoDeform; c.integrateAndUpdate(0);  // Cell::setHSize + postLoad (Cell.hpp:107,121)
}
void orc_set_next_gradv(void* h, const double* g) { Cell& c = ((Scene*)h)->cell; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c.nextGradV(i, j) = g[3 * i + j]; }
void orc_get_cell(void* h, double* hSize) { Cell& c = ((Scene*)h)->cell; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) hSize[3 * i + j] = c.hSize(i, j); }
int orc_step(void* h, int n) { ORC_TRY(h) for (int i = 0; i < n; i++) ((Scene*)h)->oneStep(); ORC_CATCH(h) }
// run only parts (for per-engine tests)
int orc_collider_run(void* h) { ORC_TRY(h)((Scene*)h)->colliderRun(); ORC_CATCH(h) }
long orc_get_step(void* h) { return ((Scene*)h)->step; }
double orc_get_dt(void* h) { return ((Scene*)h)->P.dt; }
double orc_get_time(void* h) { return ((Scene*)h)->time; }
double orc_get_verlet_dist(void* h) { return ((Scene*)h)->P.verletDist; }
int orc_get_nfullruns(void* h) { return ((Scene*)h)->nFullRuns; }
double orc_pwave_dt(void* h) { return ((Scene*)h)->pWaveDt(); }
double orc_kinetic_energy(void* h) { return ((Scene*)h)->kineticEnergy(); }
// ParticleContainer::remove (ParticleContainer.cpp:121-157): the particle's contacts are deleted at once, the particle leaves the container (its id goes to
// freeIds and is handed out again by the next add) and the contacts are marked dirty, so the collider re-initialises at the next step.  The restatement keeps the array slot and takes the
// particle out of the game: no mask (never a collision candidate again), no motion, no forces; the getters skip it.
int orc_remove_spheres(void* h, int n, const int* ids) {
  Scene* S = (Scene*)h;
  ORC_TRY(h)
  for (int k = 0; k < n; k++) {
    if (ids[k] < 0 || ids[k] >= (int)S->par.size() || S->par[ids[k]].kind != SHAPE_SPHERE || S->sph[S->par[ids[k]].idx].dead) throw std::runtime_error("remove: no such sphere #" + std::to_string(ids[k]));
    Sphere& sp = S->sph[S->par[ids[k]].idx];
    S->freeIds.push_back(ids[k]);
    sp.dead = true; sp.mask = 0; sp.vel = V3(); sp.angVel = V3(); sp.force = V3(); sp.torque = V3(); sp.flags = DOF_ALL | GRAVITY_SKIP | ENERGY_SKIP | DAMPING_SKIP;
  }
  for (auto it = S->con.begin(); it != S->con.end();) {
    const Contact& C = it->second;
    auto deadId = [&](int id) { const PRef& p = S->par[id]; return p.kind == SHAPE_SPHERE && S->sph[p.idx].dead; };
    if (deadId(C.idA) || deadId(C.idB)) it = S->con.erase(it); else ++it;
  }
  S->linDirty = true; S->conDirty = true;   // queued removal requests of erased contacts are skipped by removePending (key not found)
  ORC_CATCH(h)
}
// ids handed out by the last orc_add_spheres / ids of the live spheres in the order of orc_get_spheres
int orc_last_added_ids(void* h, int* out) { Scene* S = (Scene*)h; for (size_t i = 0; i < S->lastAdded.size(); i++) out[i] = S->lastAdded[i]; return (int)S->lastAdded.size(); }
void orc_get_sphere_ids(void* h, int* out) { Scene* S = (Scene*)h; size_t i = 0; for (auto& s : S->sph) if (!s.dead) out[i++] = s.id; }
int orc_num_spheres(void* h) { int n = 0; for (auto& s : ((Scene*)h)->sph) if (!s.dead) n++; return n; }
void orc_get_spheres(void* h, double* pos, double* ori, double* vel, double* angVel, double* force, double* torque) {
  Scene* S = (Scene*)h;
  size_t i = 0;
  for (size_t q = 0; q < S->sph.size(); q++) {
    const Sphere& s = S->sph[q];
    if (s.dead) continue;
    for (int k = 0; k < 3; k++) { if (pos) pos[3 * i + k] = s.pos[k]; if (vel) vel[3 * i + k] = s.vel[k]; if (angVel) angVel[3 * i + k] = s.angVel[k]; if (force) force[3 * i + k] = s.force[k]; if (torque) torque[3 * i + k] = s.torque[k]; }
    if (ori) { ori[4 * i] = s.ori.x; ori[4 * i + 1] = s.ori.y; ori[4 * i + 2] = s.ori.z; ori[4 * i + 3] = s.ori.w; }
    i++;
  }
}
void orc_set_sphere_state(void* h, int idx, const double* pos, const double* vel, const double* angVel) {
  Sphere& s = ((Scene*)h)->sph[idx];
  if (pos) s.pos = V3(pos[0], pos[1], pos[2]); if (vel) s.vel = V3(vel[0], vel[1], vel[2]); if (angVel) s.angVel = V3(angVel[0], angVel[1], angVel[2]);
}
void orc_get_walls(void* h, double* pos, double* force, double* torque) {
  Scene* S = (Scene*)h;
  for (size_t i = 0; i < S->walls.size(); i++) for (int k = 0; k < 3; k++) { if (pos) pos[3 * i + k] = S->walls[i].pos[k]; if (force) force[3 * i + k] = S->walls[i].force[k]; if (torque) torque[3 * i + k] = S->walls[i].torque[k]; }
}
void orc_get_wall_forces(void* h, double* force) { Scene* S = (Scene*)h; for (size_t i = 0; i < S->walls.size(); i++) for (int k = 0; k < 3; k++) force[3 * i + k] = S->walls[i].force[k]; }
long orc_num_contacts(void* h, int realOnly) { Scene* S = (Scene*)h; long n = 0; for (auto& kv : S->con) if (!realOnly || kv.second.isReal()) n++; return n; }
// contacts in canonical key order. Arrays may be NULL. Returns count written.
long orc_get_contacts(void* h, int realOnly, int* ids /*2n*/, int* cellDist /*3n*/, int* isReal, double* uN, double* forceLocal /*3n*/, double* torqueLocal, double* trsf /*9n col-major*/,
                      double* contPt /*3n*/, double* phys /*3n: kn,kt,tanPhi*/, double* gvel /*3n*/, int* stepCreated) {
  Scene* S = (Scene*)h; long n = 0;
  for (auto& kv : S->con) {
    const Contact& C = kv.second; if (realOnly && !C.isReal()) continue;
    if (ids) { ids[2 * n] = C.idA; ids[2 * n + 1] = C.idB; }
    if (cellDist) for (int k = 0; k < 3; k++) cellDist[3 * n + k] = C.cellDist[k];
    if (isReal) isReal[n] = C.isReal();
    if (uN) uN[n] = C.uN;
    for (int k = 0; k < 3; k++) { if (forceLocal) forceLocal[3 * n + k] = C.force[k]; if (torqueLocal) torqueLocal[3 * n + k] = C.torque[k]; if (contPt) contPt[3 * n + k] = C.contPt[k]; if (gvel) gvel[3 * n + k] = C.gvel[k]; }
    if (trsf) for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) trsf[9 * n + 3 * j + i] = C.trsf(i, j);
    if (phys) { phys[3 * n] = C.kn; phys[3 * n + 1] = C.kt; phys[3 * n + 2] = C.tanPhi; }
    if (stepCreated) stepCreated[n] = C.stepCreated;
    n++;
  }
  return n;
}
// ContactLoop.stress as the last step left it (Scene::stressNow), row-major 3x3.  Only defined in periodic scenes (doStress, ContactLoop.cpp:123).
int orc_get_stress(void* h, double* out9) {
  Scene* S = (Scene*)h;
  ORC_TRY(h)
  if (!S->P.periodic) throw std::runtime_error("ContactLoop.evalStress needs a periodic scene");
  const M3 st = S->stressNow();
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out9[3 * i + j] = st(i, j);
  ORC_CATCH(h)
}
void orc_set_eval_stress(void* h, int on) { ((Scene*)h)->evalStress = on != 0; }
double orc_get_gradv_work(void* h) { return ((Scene*)h)->gradVWork; }
void orc_set_track_energy(void* h, int on) { ((Scene*)h)->trackEnergy = on != 0; }
// kinetic, grav, nonviscDamp, elast, plast, broken, viscN, viscT
void orc_get_energy(void* h, double* out8) { for (int k = 0; k < Scene::EN_COUNT; k++) out8[k] = ((Scene*)h)->energy[k]; }
int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#endif
}
}
