// wdem_kernels.cuh — hand-written sm_100a kernels of the DEM hot path (float64, HBM/L2-bound; no tensor cores).
// Reference semantics are cited per function as file:line relative to the woodem/woo tree.
#pragma once
#include <stdint.h>
#include "wdem_math.cuh"

namespace wd {

// DemData flags (src/dem/Particle.hpp:198-203)
enum { F_DOF_ALL = 63, F_CLUMPED = 64, F_CLUMP = 128, F_ENERGY_SKIP = 256, F_GRAVITY_SKIP = 512, F_DAMPING_SKIP = 2048 };
// engine-internal bits kept in the same word (DemData.flags uses bits 0..11): ghost copy of a sphere owned by another subdomain;
// sphere that migrated into this subdomain at the current collider full run
// slot whose sphere left the arrays in place (a ghost of the previous generation): sorted behind the last cell and cut off by the collider full run
enum : uint32_t { F_GHOST = 1u << 30, F_ARRIVED = 1u << 29, F_DEADSLOT = 1u << 28 };
// contact flags
enum { CF_REAL = 1, CF_ORPHAN = 2, CF_DEAD = 4, CF_CLAIMED = 8, CF_KIND_SHIFT = 4, CF_KIND_MASK = 3u << 4, CF_DORMANT = 64 };
__host__ __device__ __forceinline__ int cf_kind(uint32_t f) { return (int)((f >> 4) & 3u); }
enum { KIND_SS = 0, KIND_WS = 1, KIND_FS = 2 };

// ------------------------------------------------------------------------------------------------
// Device views
// ------------------------------------------------------------------------------------------------
struct Spheres {          // one slot per sphere; slots are re-ordered by grid cell at every collider full run
  double4* posr;          // x,y,z,radius      (32 B: one sector per gather)
  double4* velm;          // vx,vy,vz,mass
  double4* angi;          // wx,wy,wz,inertia
  double4* ori;           // quaternion x,y,z,w
  double *fx, *fy, *fz, *tx, *ty, *tz;   // force / torque (streamed only)
  double *lx, *ly, *lz;   // position at the last collider full run (Aabb::nodeLastPos)
  uint32_t* flags;        // DemData.flags
  uint32_t* mask;         // Particle.mask
  int32_t* mat;           // material index
  int32_t* id;            // Particle.id
  int32_t* up;            // upload index (for downloads)
  const int32_t* clumpOfId;   // by Particle.id: index of the clump the sphere is a member of, -1 free; NULL when the scene has no clumps
  const int32_t* impOfId;     // by Particle.id: index of the node's set of impositions (DemData.impose), -1 none; NULL when no sphere carries one
};
// Impose (Particle.hpp:130-160, Impose.hpp, Impose.cpp): one resolved imposition.  Everything that depends on Scene.time / Scene.dt only (harmonic
// factors, interpolated tables) is evaluated by the host once per step with libm, so host, device and oracle see the same numbers.
#define WD_MAX_IMPOSE 16
enum { IOP_DIR = 1, IOP_AXES_SET = 2, IOP_VAR3D = 3, IOP_ANGAXIS = 4, IOP_LOCAL6 = 5, IOP_FADD = 6, IOP_FRADIAL = 7, IOP_READFT = 8, IOP_ORBIT = 9, IOP_SORBIT = 10, IOP_INTERP = 11 };
struct ImpOp {
  int kind, flags;      // DIR: flags&1 zero the lateral velocity, &2 read force (VelocityAndReadForce), &4 invF | VAR3D: &1 diff, &2 rotated frame, &4 angular velocity too | READFT: &1 node frame
  double ori[4];        // x,y,z,w
  double a[7];          // INTERP: position to reach at the end of the coming step (3), dt; ori = orientation to reach | ORBIT / SORBIT: node pos(3), omega, radius (SORBIT), dt; ori = node.ori; flags&1 rotate | DIR dir(3), magnitude | AXES_SET values(3; NaN: untouched) | VAR3D vImp(3), wImp(3) | ANGAXIS value | LOCAL6 values(6) | FADD F(3) | FRADIAL A(3), unit axis(3), F | READFT node pos(3)
  int m[6];             // ANGAXIS m[0] = axis | LOCAL6 whats(6)
};
struct ImpSlot {        // what one step of a batch needs (ring indexed by the step's position in the batch, like Counters::gflag)
  ImpOp op[WD_MAX_IMPOSE];
  double wallVel[32 * 3]; int wallHas[32];      // new velocity of walls that carry a velocity imposition (computed by the host: all DOFs blocked, nothing else changes it)
  double acc[WD_MAX_IMPOSE * 8];                // per op: sumF | F(3), T(3) of the read-force kinds; zeroed when the batch is enqueued
};
struct ImposeDev { ImpSlot* slot; const int* setStart; const int* setOps; };
__host__ __device__ __forceinline__ q4 impOri(const ImpOp& o) { q4 q; q.x = o.ori[0]; q.y = o.ori[1]; q.z = o.ori[2]; q.w = o.ori[3]; return q; }
// Impose::velocity of every imposition of the set, in order (CombinedImpose::velocity, Impose.cpp:288-292)
// `pos`: the node's position at the time of the call (only the orbit kinds read it: atan2 / sin / cos of the device's math library there, so nodes that carry them
// follow the oracle to rounding, not bit for bit).
// `nodeOri`: the node's orientation (InterpolatedMotion only: master nodes of facet groups, evaluated by the host with libm; NULL elsewhere).
__host__ __device__ __forceinline__ void impose_velocity(const ImposeDev& I, int set, d3& vel, d3& w, const d3& pos, const q4* nodeOri = nullptr) {
  for (int e = I.setStart[set]; e < I.setStart[set + 1]; e++) {
    const ImpOp& o = I.slot->op[I.setOps[e]];
    switch (o.kind) {
      case IOP_INTERP: {   // InterpolatedMotion (Impose.cpp:120-133): the velocities that take the node to (nextPos, nextOri) in one step
        vel = (mk3(o.a[0], o.a[1], o.a[2]) - pos) / o.a[3];
        if (nodeOri) {
          const q4 r = quatMul(impOri(o), quatConj(*nodeOri));      // Leapfrog::leapfrogSphericalRotate does q+ = r q0, so r = q+ q0*
          // AngleAxisr(r) (Eigen/src/Geometry/AngleAxis.h, operator=(QuaternionBase)): angle = 2 atan2(|vec|, |w|), axis = vec / (+-|vec|)
          double n = nrm(mk3(r.x, r.y, r.z)); d3 axis = mk3(1., 0., 0.); double angle = 0.;
          if (n != 0.) { angle = 2. * atan2(n, fabs(r.w)); if (r.w < 0.) n = -n; axis = mk3(r.x, r.y, r.z) / n; }
          w = (axis * angle) / o.a[3];
        }
        break;
      }
      case IOP_ORBIT: {    // CircularOrbit (Impose.cpp:22-37): tangential velocity at the approximate mid-step polar angle, around the local z axis of `node`
        const q4 q = impOri(o), qc = quatConj(q);
        const d3 loc = quatRotate(qc, pos - mk3(o.a[0], o.a[1], o.a[2]));
        const double theta = atan2(loc.y, loc.x), rho = sqrt(loc.x * loc.x + loc.y * loc.y);
        const double v2t = o.a[3] * rho, tth = theta + .5 * v2t * o.a[5];
        q4 qz; qz.x = 0.; qz.y = 0.; qz.z = sin(.5 * tth); qz.w = cos(.5 * tth);      // Quaternionr(AngleAxisr(ttheta, UnitZ))
        vel = quatRotate(q, quatRotate(qz, mk3(0., v2t, 0.)));
        if (o.flags & 1) w = quatRotate(q, mk3(0., 0., o.a[3]));
        break;
      }
      case IOP_SORBIT: {   // StableCircularOrbit (Impose.cpp:38-48): the velocity that takes the node to (radius, theta + omega dt, z) in one step
        const q4 q = impOri(o), qc = quatConj(q);
        const d3 c = mk3(o.a[0], o.a[1], o.a[2]);
        const d3 loc = quatRotate(qc, pos - c);
        const double th1 = atan2(loc.y, loc.x) + o.a[3] * o.a[5];
        const d3 pos1 = quatRotate(q, mk3(o.a[4] * cos(th1), o.a[4] * sin(th1), loc.z)) + c;
        vel = (pos1 - pos) / o.a[5];
        if (o.flags & 1) w = quatRotate(q, mk3(0., 0., o.a[3]));
        break;
      }
      case IOP_DIR: {      // HarmonicOscillation (Impose.hpp:6-13), VelocityAndReadForce (Impose.cpp:156-162)
        const d3 dir = mk3(o.a[0], o.a[1], o.a[2]);
        if (o.flags & 1) vel = dir * o.a[3];
        else { const d3 vLat = vel - dir * dot(vel, dir); vel = vLat + dir * o.a[3]; }
        break;
      }
      case IOP_AXES_SET:   // AlignedHarmonicOscillations (Impose.hpp:50-58)
        if (o.a[0] == o.a[0]) vel.x = o.a[0];
        if (o.a[1] == o.a[1]) vel.y = o.a[1];
        if (o.a[2] == o.a[2]) vel.z = o.a[2];
        break;
      case IOP_VAR3D: {    // VariableVelocity3d (Impose.cpp:193-226)
        const bool diff = o.flags & 1, rot = o.flags & 2;
        const q4 q = impOri(o), qc = quatConj(q);
        d3 v = rot ? quatRotate(qc, vel) : vel;
        v.x = (diff ? v.x : 0) + (!(o.a[0] == o.a[0]) ? (diff ? 0 : v.x) : o.a[0]);
        v.y = (diff ? v.y : 0) + (!(o.a[1] == o.a[1]) ? (diff ? 0 : v.y) : o.a[1]);
        v.z = (diff ? v.z : 0) + (!(o.a[2] == o.a[2]) ? (diff ? 0 : v.z) : o.a[2]);
        vel = rot ? quatRotate(q, v) : v;
        if (o.flags & 4) {
          d3 av = rot ? quatRotate(qc, w) : w;
          av.x = (diff ? av.x : 0) + (!(o.a[3] == o.a[3]) ? (diff ? 0 : av.x) : o.a[3]);
          av.y = (diff ? av.y : 0) + (!(o.a[4] == o.a[4]) ? (diff ? 0 : av.y) : o.a[4]);
          av.z = (diff ? av.z : 0) + (!(o.a[5] == o.a[5]) ? (diff ? 0 : av.z) : o.a[5]);
          w = rot ? quatRotate(q, av) : av;
        }
        break;
      }
      case IOP_ANGAXIS: setc(w, o.m[0], o.a[0]); break;      // VariableAlignedRotation (Impose.cpp:95-100)
      case IOP_LOCAL6: {   // Local6Dofs (Impose.cpp:72-82): ori takes global to local
        const q4 q = impOri(o), qc = quatConj(q);
        for (int i = 0; i < 6; i++) {
          if (!(o.m[i] & 1)) continue;
          if (i < 3) { d3 loc = quatRotate(q, vel); setc(loc, i, o.a[i]); vel = quatRotate(qc, loc); }
          else { d3 loc = quatRotate(q, w); setc(loc, i - 3, o.a[i]); w = quatRotate(qc, loc); }
        }
        break;
      }
      default: break;
    }
  }
}
// Impose::force at Leapfrog's reset (Leapfrog.cpp:244): on top of mass*gravity, before the contact forces of the step are added
__host__ __device__ __forceinline__ void impose_force(const ImposeDev& I, int set, d3 pos, d3& F, d3& T) {
  for (int e = I.setStart[set]; e < I.setStart[set + 1]; e++) {
    const ImpOp& o = I.slot->op[I.setOps[e]];
    switch (o.kind) {
      case IOP_FADD: F = F + mk3(o.a[0], o.a[1], o.a[2]); break;      // ConstantForce (Impose.cpp:45-47)
      case IOP_FRADIAL: {                                             // RadialForce (Impose.cpp:49-58)
        const d3 A = mk3(o.a[0], o.a[1], o.a[2]), axis = mk3(o.a[3], o.a[4], o.a[5]);
        const d3 C = A + axis * dot(pos - A, axis);
        const d3 towards = C - pos;
        if (sqn(towards) == 0) break;
        F = F - (towards / nrm(towards)) * o.a[6];
        break;
      }
      case IOP_LOCAL6: {                                              // Impose.cpp:83-87: replaces what was there
        const q4 qc = quatConj(impOri(o));
        for (int i = 0; i < 6; i++) {
          if (!(o.m[i] & 2)) continue;
          d3 u = mk3(0, 0, 0); setc(u, i % 3, 1.);
          if (i < 3) F = quatRotate(qc, o.a[i] * u); else T = quatRotate(qc, o.a[i] * u);
        }
        break;
      }
      default: break;
    }
  }
}
#ifdef __CUDACC__
// Impose::readForce before the reset (Leapfrog.cpp:238); sums over the nodes of an imposition (OpenMPAccumulator in the reference: unordered there as well)
__device__ __forceinline__ void impose_read_force(const ImposeDev& I, int set, d3 pos, d3 F, d3 T) {
  for (int e = I.setStart[set]; e < I.setStart[set + 1]; e++) {
    const int ix = I.setOps[e]; const ImpOp& o = I.slot->op[ix]; double* acc = I.slot->acc + 8 * ix;
    if (o.kind == IOP_DIR && (o.flags & 2)) { const double f = dot(F, mk3(o.a[0], o.a[1], o.a[2])); atomicAdd(acc, ((o.flags & 4) ? -1. : 1.) * f); }      // Impose.cpp:175-177
    else if (o.kind == IOP_READFT) {                                                                                                                      // Impose.cpp:144-155
      d3 f = F, t = T;
      if (o.flags & 1) { const q4 qc = quatConj(impOri(o)); f = quatRotate(qc, F); t = quatRotate(qc, T + cross(pos - mk3(o.a[0], o.a[1], o.a[2]), F)); }
      atomicAdd(acc + 1, f.x); atomicAdd(acc + 2, f.y); atomicAdd(acc + 3, f.z); atomicAdd(acc + 4, t.x); atomicAdd(acc + 5, t.y); atomicAdd(acc + 6, t.z);
    }
  }
}
#endif

// Clumps of spheres with mass (ClumpData, src/dem/Clump.hpp): one entry per master node; members are addressed by particle id (their slots move at collider full runs)
struct Clumps {
  int n;
  double4 *pos /*x,y,z,mass*/, *ori /*x,y,z,w*/, *vel, *angVel, *angMom, *force, *torque;
  const double4* inertia; const uint32_t* flags;
  const int* memStart; const int32_t* memId; const double4* relPos; const double4* relOri;
  const int* impSet;      // per clump: set of impositions on the master node, -1 none; NULL when no clump carries one
};

// moving: the wall's node is in DemField.nodes (added with a nonzero velocity: Particle::guessMoving, Particle.cpp:638-640) and is translated by Leapfrog
struct Wall { double pos[3]; double vel[3]; double lastPos[3]; int32_t axis, sense; uint32_t mask; int32_t mat; int32_t id; int32_t moving; };
// group1: 1 + index of the kinematic group (clump of boundary particles with a prescribed rigid motion) the facet belongs to, 0 = static;
// rel: node positions in the group's frame (ClumpData.relPos, Clump.cpp:252); lastv: Aabb::nodeLastPos
struct Facet { double v[9]; double vvel[9]; double fakeVel[3]; double halfThick; double n21lim[3]; double bmin[3], bmax[3]; uint32_t mask; int32_t mat; int32_t id; int32_t group1;
               double rel[9]; double lastv[9]; int32_t nodeImp1, pad1; };      // nodeImp1: 1 + the set of impositions its three nodes carry (0: none)
#define WD_MAX_GROUPS 8
struct GroupU { d3 pos; q4 ori; d3 vel, angVel; };
struct GroupsU { int n; GroupU g[WD_MAX_GROUPS]; };
struct FacetGrid { double ox, oy, oz, inv; int nx, ny, nz; const int* cellStart; const int* cellFacet; };
struct Material { int32_t kind; int32_t pad; double density, young, ktDivKn, tanPhi, surfEnergy, alpha; };

struct Contacts {         // one entry per candidate pair (real or potential)
  int2* pair;             // SS: (slotA, slotB) with id[slotA] < id[slotB]; WS: (wall index, slot); FS: (facet index, slot)
  uint32_t* flags;
  int32_t* stepCreated;
  int32_t *cdx, *cdy, *cdz;            // Contact.cellDist (periodic only)
  // The persistent state of a contact lives in four double4 planes (every per-thread access of a warp is one fully used 32-byte sector per lane, 1 KB contiguous per
  // warp; the force reduction gathers rC and rD; the history carry-over of a full run copies 4 sectors per contact instead of the 14 scalar planes of round 1):
  //   rA: L6Geom.trsf column 0 (xyz), w = CPhys.force[1] (ft0)        rB: trsf column 1 (xyz), w = CPhys.force[2] (ft1)
  //   rC: contact point CGeom.node.pos (xyz), w = uN                   rD: global force on pA (xyz), w = CPhys.force[0] (fn)
  // (trsf column 2 = c0 x c1 is recomputed bit-identically.)  An array-of-records layout (one 128-byte line per contact, staged through shared memory with cp.async)
  // was measured and is slower: 0.299 ms vs 0.227 ms for the contact kernel at 1M spheres — strided per-lane accesses cost 4x the L1 wavefronts.
  double4 *rA, *rB, *rC, *rD;
  double4* ph;            // Cp2 constants, written once at creation: linear (kn, kt, tanPhi, -); Hertz (kn0=K*sqrt(R), kt0, tanPhi, alpha_sqrtMK)
  double* ph4;            // Hertz only: Fna=4*pi*R*gamma (or -1: alpha!=0)
  double4* tq;            // CPhys.torque xyz (only with rolling/twisting resistance)
  double4* Tg;            // global contact torque on pA (rolling resistance only)
  uint32_t* realBits;     // 1 bit per contact: real (CF_REAL and not dead); set at full runs (k_realbits), flipped by the contact kernels when a contact is
                          // created / broken.  Half a megabyte for 4 M candidates, i.e. L2-resident: the force reduction tests it instead of gathering Fg.
};

struct StepU {            // per-step uniforms
  double dt, damping; d3 gravity; int hasGravity, reset; double maxD2;
  int periodic, homoDeform;
  m33 hSize;              // Cell.hSize (intrShiftPos)
  m33 shiftVelM;          // gradV*pprevHsize (GRADV2) or gradV*hSize (VEL, VEL_2ND) or 0  (Cell::intrShiftVel)
  m33 gradV, nextGradV, dGradV, midGradV, ImLL4hInv, IpLL4h, LmL; d3 spinVec, deltaSpinVec;
  long long step; int renormNow; int noRatch;
  const int* gIn; int* gOut;
  int impSlot;      // position of this step in the ring of Impose slots (ImpSlot)      // this step's slots of Counters::gflag
  double absDistFactor;
  int law; int noSlip, noBreak, noFrict, noAttraction; int rolling; double relRollStiff, relTwistStiff, rollTanPhi;
  double poisson, en, viscAlpha;  // viscAlpha = -sqrt(5) ln(en)/sqrt(ln(en)^2+pi^2), host-computed (Hertz.cpp:69)
  double* energy;                 // Scene.trackEnergy: device accumulators (EN_*), else nullptr; only the TRACK kernel variants look at it
  int trackLeapfrog;              // Leapfrog.cpp:130-132: reallyTrackEnergy (off in step 0 of a periodic scene with a velocity gradient)
};
// EnergyTracker entries of this path (core/EnergyTracker.hpp): "kinetic" and "elast" are resettable (zeroed in the step prologue, Scene.cpp:416)
enum { EN_KINETIC = 0, EN_GRAV, EN_DAMP, EN_ELAST, EN_PLAST, EN_BROKEN, EN_VISCN, EN_VISCT, EN_COUNT };
// EnergyTracker::add from a (possibly divergent) warp: the lowest active lane adds up the values of the active lanes in lane order and issues one atomic
__device__ __forceinline__ void energy_add(double* en, int ix, double v) {
  if (v != v) v = 0.;                          // NaN is ignored (EnergyTracker.cpp:30)
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  double sum = 0.;
  for (int src = 0; src < 32; src++) {
    if (!((m >> src) & 1u)) continue;            // uniform over the active lanes
    const double o = __shfl_sync(m, v, src);
    sum += o;
  }
  if (lane == leader && sum != 0.) atomicAdd(en + ix, sum);
}

// gflag: the Verlet-dirty flags of a BATCH of steps enqueued without a host wait in between.  Step k of a batch reads gflag[k] ("a full run was already due when I
// started: I am void, pass it on") and raises gflag[k+1] ("a sphere moved farther than verletDist in this step": contacts and force reduction of this step and every
// kernel of the later steps return at once).  gflag[0] is always 0.  The host reads the array once per batch, does the collider full run for the first raised step
// and goes on from there.  [error, gflag...] is one block of ints so that decomposed runs can max-reduce it in one call.
#define WD_KMAX 32
struct Counters { unsigned long long nReal, nPotEval, nCreated, nBroken; int error; int gflag[WD_KMAX + 1]; int errA, errB; };
enum { ERR_NONE = 0, ERR_NAN_FORCE = 1, ERR_NAN_POS = 2, ERR_NO_PHYS = 3, ERR_SCHWARZ = 4, ERR_FACET_IMPOSSIBLE = 5, ERR_AMBIGUOUS_PERI = 6, ERR_PEER_TIMEOUT = 7, ERR_CLUMP_INERTIA = 8 };

__device__ __forceinline__ double signd(double f) { return f < 0 ? -1. : (f > 0 ? 1. : 0.); }  // Mathr::Sign (Math.hpp:145)

// ------------------------------------------------------------------------------------------------
// Leapfrog::run, spherical branch (src/dem/Leapfrog.cpp:104-255; helpers :33-48, :258-297) fused with the
// Verlet-distance dirtiness test of AabbCollider::aabbIsDirty (src/dem/Collision.cpp:80-123).
// Streaming kernel: one thread per sphere.
// ------------------------------------------------------------------------------------------------
// host-state staging of the end-to-end path (wdem_step_e2e): pos/vel/angVel come from / go to buffers in upload order
struct LeapIO { const double* pos; const double* vel; const double* angVel; double* posOut; double* velOut; };
template <bool EXT, bool TRACK, bool IMP = false>
__device__ __forceinline__ void leapfrog_body(const Spheres& S, int i, const StepU& u, Counters* cnt, const LeapIO& io, long long ut, const ImposeDev* I = nullptr) {
  const uint32_t fl = S.flags[i];
  int impSet = -1;
  if (IMP) impSet = S.impOfId[S.id[i]];
  if (fl & (F_GHOST | F_CLUMPED)) return;   // ghost: integrated by its owner, refreshed by the halo exchange; clumped: follows its master node (k_clumps_step; Leapfrog.cpp:153)
  double4 pr = S.posr[i], vm = S.velm[i], ai = S.angi[i], qo = S.ori[i];
  if (EXT) {
    pr.x = io.pos[3 * ut]; pr.y = io.pos[3 * ut + 1]; pr.z = io.pos[3 * ut + 2];
    vm.x = io.vel[3 * ut]; vm.y = io.vel[3 * ut + 1]; vm.z = io.vel[3 * ut + 2];
    ai.x = io.angVel[3 * ut]; ai.y = io.angVel[3 * ut + 1]; ai.z = io.angVel[3 * ut + 2];
  }
  d3 f = mk3(S.fx[i], S.fy[i], S.fz[i]);
  d3 t = mk3(S.tx[i], S.ty[i], S.tz[i]);
  d3 pos = xyz(pr), vel = xyz(vm), w = xyz(ai);
  const double mass = vm.w, inertia = ai.w, dt = u.dt;
  const bool blockedAll = (fl & F_DOF_ALL) == F_DOF_ALL, blockedNone = (fl & F_DOF_ALL) == 0;
  const bool damp = (u.damping != 0. && !(fl & F_DAMPING_SKIP));
  d3 linAccel = mk3(0, 0, 0), angAccel = mk3(0, 0, 0);
  d3 trkV = mk3(0, 0, 0), trkW = mk3(0, 0, 0);                             // pprevFluctVel / pprevFluctAngVel for doKineticEnergy (zero for a fully blocked node, :198)
  if (TRACK && u.trackLeapfrog) {
    // Leapfrog::doDampingDissipation (:53-65) and doGravityWork (:67-74), with the values at the start of the step
    if (damp && !(fl & F_ENERGY_SKIP))
      energy_add(u.energy, EN_DAMP, dot(mk3(fabs(vel.x), fabs(vel.y), fabs(vel.z)), mk3(fabs(f.x), fabs(f.y), fabs(f.z))) * u.damping * dt +
                                        dot(mk3(fabs(w.x), fabs(w.y), fabs(w.z)), mk3(fabs(t.x), fabs(t.y), fabs(t.z))) * u.damping * dt);
    if (u.hasGravity && !(fl & F_GRAVITY_SKIP)) {
      double gr = 0;
      if (blockedNone) gr = -dot(u.gravity, vel) * mass * dt;
      else {
        if (!(fl & 1u)) gr -= u.gravity.x * vel.x * mass * dt;
        if (!(fl & 2u)) gr -= u.gravity.y * vel.y * mass * dt;
        if (!(fl & 4u)) gr -= u.gravity.z * vel.z * mass * dt;
      }
      energy_add(u.energy, EN_GRAV, gr);
    }
  }
  if (!blockedAll) {
    if (blockedNone) linAccel = f / mass;                                  // computeAccel :37-42
    else {
      if (!(fl & 1u)) linAccel.x += f.x / mass;
      if (!(fl & 2u)) linAccel.y += f.y / mass;
      if (!(fl & 4u)) linAccel.z += f.z / mass;
    }
    d3 pfv, pfw;
    if (u.periodic) {                                                      // Cell::pprevFluctVel / pprevFluctAngVel (Cell.cpp:36-57)
      if (u.homoDeform == 2 || u.homoDeform == 3) pfv = vel - mulMV(u.gradV, pos);
      else if (u.homoDeform == 4) pfv = vel - mulMV(u.gradV, pos - (dt / 2) * vel);
      else pfv = vel;
      pfw = (u.homoDeform == 4) ? w - u.spinVec : w;
    } else { pfv = vel; pfw = w; }
    if (TRACK) { trkV = pfv; trkW = pfw; }
    if (damp) {                                                            // nonviscDamp2nd :33-35
      linAccel.x *= 1 - u.damping * signd(f.x * (pfv.x + 0.5 * dt * linAccel.x));
      linAccel.y *= 1 - u.damping * signd(f.y * (pfv.y + 0.5 * dt * linAccel.y));
      linAccel.z *= 1 - u.damping * signd(f.z * (pfv.z + 0.5 * dt * linAccel.z));
    }
    if (u.periodic && u.homoDeform == 4) vel = mulMV(u.ImLL4hInv, (mulMV(u.LmL, pos) + mulMV(u.IpLL4h, vel)) + linAccel * dt);
    else vel = vel + dt * linAccel;
    if (inertia != 0.) {
      if (blockedNone) angAccel = mk3(t.x / inertia, t.y / inertia, t.z / inertia);
      else {
        if (!(fl & 8u)) angAccel.x += t.x / inertia;
        if (!(fl & 16u)) angAccel.y += t.y / inertia;
        if (!(fl & 32u)) angAccel.z += t.z / inertia;
      }
      if (damp) {
        angAccel.x *= 1 - u.damping * signd(t.x * (pfw.x + 0.5 * dt * angAccel.x));
        angAccel.y *= 1 - u.damping * signd(t.y * (pfw.y + 0.5 * dt * angAccel.y));
        angAccel.z *= 1 - u.damping * signd(t.z * (pfw.z + 0.5 * dt * angAccel.z));
      }
      w = w + dt * angAccel;
      if (u.periodic && u.homoDeform == 4) w = w - u.deltaSpinVec;
    }
    if (IMP && impSet >= 0) impose_velocity(*I, impSet, vel, w, pos);      // re-applied so that forces do not alter what is prescribed (Leapfrog.cpp:203-205)
  } else if (u.periodic && u.homoDeform == 4) vel = mulMV(u.ImLL4hInv, mulMV(u.LmL, pos) + mulMV(u.IpLL4h, vel));
  if (u.periodic && u.homoDeform != 4) {                                   // applyPeriodicCorrections :258-273
    if (u.homoDeform == 2 || u.homoDeform == 3) {
      if (u.homoDeform == 3) vel = vel + mulMV(u.midGradV, vel - (dt / 2.) * linAccel) * dt;
      vel = vel + mulMV(u.dGradV, pos);
    } else if (u.homoDeform == 1) pos = pos + mulMV(u.nextGradV, pos) * dt;
  }
  if (TRACK && u.trackLeapfrog && !(fl & F_ENERGY_SKIP)) {                  // doKineticEnergy :76-98
    const d3 cfv = trkV + (.5 * dt) * linAccel, cfw = trkW + (.5 * dt) * angAccel;
    const double Etrans = .5 * mass * sqn(cfv);
    double Erot = 0.5 * dot(cfw, mk3(inertia * cfw.x, inertia * cfw.y, inertia * cfw.z));
    if (!(Erot == Erot) || isinf(inertia)) Erot = 0;
    energy_add(u.energy, EN_KINETIC, Etrans + Erot);
  }
  pos = pos + vel * dt;                                                    // leapfrogTranslate :275-278
  // leapfrogSphericalRotate :280-297 (orientation of a sphere feeds no contact functor: tolerance-level parity)
  if (w.x != 0. || w.y != 0. || w.z != 0.) {
    double angle = nrm(w); d3 ax = w / angle;
    double ha = 0.5 * (angle * dt), sn, cs;
    sincos(ha, &sn, &cs);
    double qw = cs, qx = sn * ax.x, qy = sn * ax.y, qz = sn * ax.z;
    double ow = qo.w, ox = qo.x, oy = qo.y, oz = qo.z;
    qo.w = qw * ow - qx * ox - qy * oy - qz * oz;
    qo.x = qw * ox + qx * ow + qy * oz - qz * oy;
    qo.y = qw * oy + qy * ow + qz * ox - qx * oz;
    qo.z = qw * oz + qz * ow + qx * oy - qy * ox;
  }
  {
    double nn = sqrt(((qo.x * qo.x + qo.y * qo.y) + qo.z * qo.z) + qo.w * qo.w);
    qo.x /= nn; qo.y /= nn; qo.z /= nn; qo.w /= nn;
  }
  if (IMP && impSet >= 0) {
    impose_read_force(*I, impSet, pos, f, t);      // the forces of the previous step, before they are reset (Leapfrog.cpp:238); the force impositions act in k_reduce
    impose_velocity(*I, impSet, vel, w, pos);      // Leapfrog.cpp:248
  }
  S.posr[i] = make_double4(pos.x, pos.y, pos.z, pr.w);
  S.velm[i] = make_double4(vel.x, vel.y, vel.z, mass);
  S.angi[i] = make_double4(w.x, w.y, w.z, inertia);
  S.ori[i] = qo;
  // Verlet test (Collision.cpp:113-121): d2 > maxD2 strictly
  d3 dl = mk3(S.lx[i], S.ly[i], S.lz[i]) - pos;
  double d2 = fmax(0., sqn(dl));
  if (d2 > u.maxD2) *u.gOut = 1;
  if (!(d2 == d2)) { *u.gOut = 1; cnt->error = ERR_NAN_POS; cnt->errA = S.id[i]; }
  if (EXT) {
    io.posOut[3 * ut] = pos.x; io.posOut[3 * ut + 1] = pos.y; io.posOut[3 * ut + 2] = pos.z;
    io.velOut[3 * ut] = vel.x; io.velOut[3 * ut + 1] = vel.y; io.velOut[3 * ut + 2] = vel.z;
  }
}
// Leapfrog::run for clump master nodes (Leapfrog.cpp:148-255 with isClump): ClumpData::forceTorqueFromMembers (Clump.cpp:272-280), spherical or aspherical
// integration (leapfrogAsphericalRotate, Leapfrog.cpp:299-331), reset, ClumpData::applyToMembers (Clump.cpp:293-307), and the members' Verlet test.
// One thread per clump: members are few, and their forces must be summed in member order.
__device__ __forceinline__ q4 dotQ(d3 w, q4 Q) {   // Leapfrog::DotQ (Leapfrog.cpp:324-331)
  q4 d;
  d.w = (-Q.x * w.x - Q.y * w.y - Q.z * w.z) / 2;
  d.x = (Q.w * w.x - Q.z * w.y + Q.y * w.z) / 2;
  d.y = (Q.z * w.x + Q.w * w.y - Q.x * w.z) / 2;
  d.z = (-Q.y * w.x + Q.x * w.y + Q.w * w.z) / 2;
  return d;
}
__global__ void __launch_bounds__(128) k_clumps_step(Clumps K, Spheres S, const int* __restrict__ id2slot, StepU u, Counters* cnt, ImposeDev I) {
  if (*u.gIn) return;      // void step (k_leapfrog passes the flag on)
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= K.n) return;
  const uint32_t fl = K.flags[ci];
  const double4 p4 = K.pos[ci], o4 = K.ori[ci], in4 = K.inertia[ci];
  d3 pos = xyz(p4), vel = xyz(K.vel[ci]), w = xyz(K.angVel[ci]), f = xyz(K.force[ci]), t = xyz(K.torque[ci]), angMom = xyz(K.angMom[ci]);
  const d3 inertia = xyz(in4); const double mass = p4.w, dt = u.dt;
  q4 ori; ori.x = o4.x; ori.y = o4.y; ori.z = o4.z; ori.w = o4.w;
  const int m0 = K.memStart[ci], m1 = K.memStart[ci + 1];
  const int impSet = K.impSet ? K.impSet[ci] : -1;      // DemData.impose of the master node
  const bool blockedAll = (fl & F_DOF_ALL) == F_DOF_ALL, blockedNone = (fl & F_DOF_ALL) == 0, blockedAllRot = (fl & 56u) == 56u;
  const bool damp = (u.damping != 0. && !(fl & F_DAMPING_SKIP));
  if (!blockedAll) {
    for (int m = m0; m < m1; m++) {
      const int s = id2slot[K.memId[m]];
      const d3 mf = mk3(S.fx[s], S.fy[s], S.fz[s]), mt = mk3(S.tx[s], S.ty[s], S.tz[s]);
      f = f + mf; t = t + (mt + cross(xyz(S.posr[s]) - pos, mf));
    }
  }
  const bool aspherical = !(inertia.x == inertia.y && inertia.y == inertia.z), useAspherical = aspherical && !blockedAllRot;      // Particle.hpp:275-277
  const bool hasInertia = !(inertia.x == 0. && inertia.y == 0. && inertia.z == 0.);
  d3 linAccel = mk3(0, 0, 0), angAccel = mk3(0, 0, 0);
  if (!blockedAll) {
    if (blockedNone) linAccel = f / mass;
    else {
      if (!(fl & 1u)) linAccel.x += f.x / mass;
      if (!(fl & 2u)) linAccel.y += f.y / mass;
      if (!(fl & 4u)) linAccel.z += f.z / mass;
    }
    d3 pfv = vel, pfw = w;      // Cell::pprevFluctVel / pprevFluctAngVel (Cell.cpp:36-57): the clump's master node is a node like any other
    if (u.periodic) {
      if (u.homoDeform == 2 || u.homoDeform == 3) pfv = vel - mulMV(u.gradV, pos);
      else if (u.homoDeform == 4) pfv = vel - mulMV(u.gradV, pos - (dt / 2) * vel);
      if (u.homoDeform == 4) pfw = w - u.spinVec;
    }
    if (damp) {
      linAccel.x *= 1 - u.damping * signd(f.x * (pfv.x + 0.5 * dt * linAccel.x));
      linAccel.y *= 1 - u.damping * signd(f.y * (pfv.y + 0.5 * dt * linAccel.y));
      linAccel.z *= 1 - u.damping * signd(f.z * (pfv.z + 0.5 * dt * linAccel.z));
    }
    if (u.periodic && u.homoDeform == 4) vel = mulMV(u.ImLL4hInv, (mulMV(u.LmL, pos) + mulMV(u.IpLL4h, vel)) + linAccel * dt);
    else vel = vel + dt * linAccel;
    if (hasInertia) {
      if (!useAspherical) {
        if (blockedNone) angAccel = mk3(t.x / inertia.x, t.y / inertia.y, t.z / inertia.z);
        else {
          if (!(fl & 8u)) angAccel.x += t.x / inertia.x;
          if (!(fl & 16u)) angAccel.y += t.y / inertia.y;
          if (!(fl & 32u)) angAccel.z += t.z / inertia.z;
        }
        if (damp) {
          angAccel.x *= 1 - u.damping * signd(t.x * (pfw.x + 0.5 * dt * angAccel.x));
          angAccel.y *= 1 - u.damping * signd(t.y * (pfw.y + 0.5 * dt * angAccel.y));
          angAccel.z *= 1 - u.damping * signd(t.z * (pfw.z + 0.5 * dt * angAccel.z));
        }
        w = w + dt * angAccel;
        if (u.periodic && u.homoDeform == 4) w = w - u.deltaSpinVec;
      } else {
        if (fl & 8u) t.x = 0; if (fl & 16u) t.y = 0; if (fl & 32u) t.z = 0;      // Leapfrog.cpp:188
        if (damp) {                                                           // nonviscDamp1st (Leapfrog.cpp:28-30) with the fluctuation angular velocity
          t.x *= 1 - u.damping * signd(t.x * pfw.x); t.y *= 1 - u.damping * signd(t.y * pfw.y); t.z *= 1 - u.damping * signd(t.z * pfw.z);
        }
      }
    }
    if (impSet >= 0) impose_velocity(I, impSet, vel, w, pos, &ori);      // Leapfrog.cpp:203-205
  } else if (u.periodic && u.homoDeform == 4) vel = mulMV(u.ImLL4hInv, mulMV(u.LmL, pos) + mulMV(u.IpLL4h, vel));      // fixed node, gradV2 (Leapfrog.cpp:206-210)
  if (u.periodic && u.homoDeform != 4) {      // applyPeriodicCorrections (Leapfrog.cpp:258-273)
    if (u.homoDeform == 2 || u.homoDeform == 3) {
      if (u.homoDeform == 3) vel = vel + mulMV(u.midGradV, vel - (dt / 2.) * linAccel) * dt;
      vel = vel + mulMV(u.dGradV, pos);
    } else if (u.homoDeform == 1) pos = pos + mulMV(u.nextGradV, pos) * dt;
  }
  pos = pos + vel * dt;
  if (!useAspherical) {      // leapfrogSphericalRotate (Leapfrog.cpp:280-297)
    if (w.x != 0. || w.y != 0. || w.z != 0.) {
      const double angle = nrm(w); const d3 ax = w / angle;
      double ha = 0.5 * (angle * dt), sn, cs;
      sincos(ha, &sn, &cs);
      q4 q; q.w = cs; q.x = sn * ax.x; q.y = sn * ax.y; q.z = sn * ax.z;
      ori = quatMul(q, ori);
    }
    ori = quatNormalized(ori);
  } else {                   // leapfrogAsphericalRotate (Leapfrog.cpp:299-321)
    if (!hasInertia) { cnt->error = ERR_CLUMP_INERTIA; cnt->errA = ci; return; }
    const q4 qc = quatConj(ori);
    if (!(angMom.x == angMom.x) || !(angMom.y == angMom.y) || !(angMom.z == angMom.z)) {
      const d3 wl = quatRotate(qc, w); angMom = mk3(inertia.x * wl.x, inertia.y * wl.y, inertia.z * wl.z);
    }
    const m33 A = quatToMat(qc);
    const d3 l_n = angMom + (dt / 2) * t;
    const d3 l_b_n = mulMV(A, l_n);
    const d3 w_b_n = mk3(l_b_n.x / inertia.x, l_b_n.y / inertia.y, l_b_n.z / inertia.z);
    const q4 dq_n = dotQ(w_b_n, ori);
    q4 Qh; Qh.w = ori.w + (dt / 2) * dq_n.w; Qh.x = ori.x + (dt / 2) * dq_n.x; Qh.y = ori.y + (dt / 2) * dq_n.y; Qh.z = ori.z + (dt / 2) * dq_n.z;
    angMom = angMom + dt * t;
    const d3 l_b_h = mulMV(A, angMom);
    const d3 w_b_h = mk3(l_b_h.x / inertia.x, l_b_h.y / inertia.y, l_b_h.z / inertia.z);
    const q4 dq_h = dotQ(w_b_h, Qh);
    ori.w = ori.w + dt * dq_h.w; ori.x = ori.x + dt * dq_h.x; ori.y = ori.y + dt * dq_h.y; ori.z = ori.z + dt * dq_h.z;
    w = quatRotate(ori, w_b_h);
    ori = quatNormalized(ori);
  }
  if (u.reset) {      // gravity acts on the master node, not on the members; Impose::force on top (Leapfrog.cpp:240-245)
    f = (u.hasGravity && !(fl & F_GRAVITY_SKIP)) ? (mass * u.gravity) : mk3(0, 0, 0); t = mk3(0, 0, 0);
    if (impSet >= 0) impose_force(I, impSet, pos, f, t);
  }
  if (impSet >= 0) impose_velocity(I, impSet, vel, w, pos, &ori);      // Leapfrog.cpp:248, before applyToMembers (:252)
  K.pos[ci] = make_double4(pos.x, pos.y, pos.z, mass); K.ori[ci] = make_double4(ori.x, ori.y, ori.z, ori.w);
  K.vel[ci] = make_double4(vel.x, vel.y, vel.z, 0); K.angVel[ci] = make_double4(w.x, w.y, w.z, 0); K.angMom[ci] = make_double4(angMom.x, angMom.y, angMom.z, 0);
  K.force[ci] = make_double4(f.x, f.y, f.z, 0); K.torque[ci] = make_double4(t.x, t.y, t.z, 0);
  bool dirty = false, nan = false;
  for (int m = m0; m < m1; m++) {
    const int s = id2slot[K.memId[m]];
    const double4 rp = K.relPos[m], ro = K.relOri[m];
    const d3 mp = pos + quatRotate(ori, xyz(rp));
    q4 rq; rq.x = ro.x; rq.y = ro.y; rq.z = ro.z; rq.w = ro.w;
    const q4 mo = quatMul(ori, rq);
    const d3 mv = vel + cross(w, mp - pos);
    S.posr[s] = make_double4(mp.x, mp.y, mp.z, S.posr[s].w);
    S.velm[s] = make_double4(mv.x, mv.y, mv.z, S.velm[s].w);
    S.angi[s] = make_double4(w.x, w.y, w.z, S.angi[s].w);
    S.ori[s] = make_double4(mo.x, mo.y, mo.z, mo.w);
    if (u.reset) { S.fx[s] = 0; S.fy[s] = 0; S.fz[s] = 0; S.tx[s] = 0; S.ty[s] = 0; S.tz[s] = 0; }
    const d3 dl = mk3(S.lx[s], S.ly[s], S.lz[s]) - mp;      // Verlet test (Collision.cpp:113-121)
    const double d2 = fmax(0., sqn(dl));
    if (d2 > u.maxD2) dirty = true;
    if (!(d2 == d2)) nan = true;
  }
  if (dirty || nan) *u.gOut = 1;
  if (nan) { cnt->error = ERR_NAN_POS; cnt->errA = K.memId[m0]; }
}

template <bool TRACK>
__global__ void __launch_bounds__(256, 4) k_leapfrog(Spheres S, int n, StepU u, Counters* cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (*u.gIn) { if (i == 0) *u.gOut = 1; return; }      // an earlier step of this batch is waiting for its collider full run: void, and say so to the next step
  if (i >= n) return;
  leapfrog_body<false, TRACK>(S, i, u, cnt, LeapIO{}, 0);
}
// the same with the Impose hooks (scenes in which some sphere carries an imposition)
template <bool TRACK>
__global__ void __launch_bounds__(256, 4) k_leapfrog_imp(Spheres S, int n, StepU u, Counters* cnt, ImposeDev I) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (*u.gIn) { if (i == 0) *u.gOut = 1; return; }
  if (i >= n) return;
  leapfrog_body<false, TRACK, true>(S, i, u, cnt, LeapIO{}, 0, &I);
}
// the same integrator over a chunk [first, first+cnt) of the UPLOAD order: inputs are read from, and the new pos/vel written to, staging
// buffers (coalesced), the rest of the state is gathered by slot.  Lets host<->device copies of other chunks overlap with the integration.
template <bool TRACK>
__global__ void __launch_bounds__(256, 4) k_leapfrog_io(Spheres S, StepU u, Counters* cnt, const int* __restrict__ up2slot, long long first, long long cnt_, LeapIO io) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (*u.gIn) { if (first + t == 0) *u.gOut = 1; return; }
  if (t >= cnt_) return;
  const int slot = up2slot[first + t];
  if (slot < 0) return;      // an id the subdomain does not own (wdem_step_e2e_owned counts and reports it)
  leapfrog_body<true, TRACK>(S, slot, u, cnt, io, first + t);
}

// ClumpData::applyToMembers (Clump.cpp:293-307) for facets of kinematic groups, after the host has integrated the groups' master nodes: node position
// = clumpPos + clumpOri*relPos, node velocity = clump.vel + clump.angVel x (pos - clumpPos); plus the Verlet test of the facet's Aabb (every node,
// Collision.cpp:113-121; rotation is ignored for facets: maxRot=-1, Facet.cpp:277).  mode 1 (at a full run): refresh Aabb (Bo1_Facet_Aabb::go, Facet.cpp:275-298)
// and Aabb::nodeLastPos of ALL facets.
__global__ void __launch_bounds__(128) k_facets_step(Facet* facets, int n, GroupsU G, StepU u, Counters* cnt, int mode, ImposeDev I) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  Facet* F = facets + f;
  if (mode == 1) {
    const double ht = F->halfThick;
    for (int k = 0; k < 3; k++) {
      double mn = F->v[k] - ht, mx = F->v[k] + ht;
      for (int j = 1; j < 3; j++) { mn = fmin(mn, F->v[3 * j + k] - ht); mx = fmax(mx, F->v[3 * j + k] + ht); }
      F->bmin[k] = mn; F->bmax[k] = mx;
    }
    for (int k = 0; k < 9; k++) F->lastv[k] = F->v[k];
    return;
  }
  if (*u.gIn) return;
  if (F->nodeImp1 > 0) {
    // nodes that carry an imposition are in DemField.nodes: Leapfrog's fixed-node branch translates them with their velocity (leapfrogTranslate, Leapfrog.cpp:195-199)
    // and Impose::velocity at the end of the node's step (:248) sets the velocity of the next one (py/triangulated.py:59-64: a drum that really rotates)
    double d2 = 0;
    for (int i = 0; i < 3; i++) {
      d3 vv = mk3(F->vvel[3 * i], F->vvel[3 * i + 1], F->vvel[3 * i + 2]), wdum = mk3(0, 0, 0);
      d3 p0 = mk3(F->v[3 * i], F->v[3 * i + 1], F->v[3 * i + 2]);
      if (u.periodic) {      // the cell's corrections of the fixed branch (Leapfrog.cpp:206-215,258-273), as for walls
        if (u.homoDeform == 4) vv = mulMV(u.ImLL4hInv, mulMV(u.LmL, p0) + mulMV(u.IpLL4h, vv));
        else if (u.homoDeform == 2 || u.homoDeform == 3) {
          if (u.homoDeform == 3) vv = vv + mulMV(u.midGradV, vv - (u.dt / 2.) * mk3(0, 0, 0)) * u.dt;
          vv = vv + mulMV(u.dGradV, p0);
        } else if (u.homoDeform == 1) p0 = p0 + mulMV(u.nextGradV, p0) * u.dt;
      }
      const d3 p = p0 + vv * u.dt;
      impose_velocity(I, F->nodeImp1 - 1, vv, wdum, p);
      F->v[3 * i] = p.x; F->v[3 * i + 1] = p.y; F->v[3 * i + 2] = p.z;
      F->vvel[3 * i] = vv.x; F->vvel[3 * i + 1] = vv.y; F->vvel[3 * i + 2] = vv.z;
      d2 = fmax(d2, sqn(mk3(F->lastv[3 * i], F->lastv[3 * i + 1], F->lastv[3 * i + 2]) - p));
    }
    if (d2 > u.maxD2 || !(d2 == d2)) *u.gOut = 1;
    return;
  }
  const int gi = F->group1 - 1;
  if (gi < 0) return;
  const GroupU g = G.g[gi];
  double d2 = 0;
  for (int i = 0; i < 3; i++) {
    const d3 p = g.pos + quatRotate(g.ori, mk3(F->rel[3 * i], F->rel[3 * i + 1], F->rel[3 * i + 2]));
    const d3 vv = g.vel + cross(g.angVel, p - g.pos);
    F->v[3 * i] = p.x; F->v[3 * i + 1] = p.y; F->v[3 * i + 2] = p.z;
    F->vvel[3 * i] = vv.x; F->vvel[3 * i + 1] = vv.y; F->vvel[3 * i + 2] = vv.z;
    d2 = fmax(d2, sqn(mk3(F->lastv[3 * i], F->lastv[3 * i + 1], F->lastv[3 * i + 2]) - p));
  }
  if (d2 > u.maxD2 || !(d2 == d2)) *u.gOut = 1;
}
// Leapfrog for boundary nodes that are in DemField.nodes (moving walls: mass 0, all DOFs blocked => the "fixed particle" branch, Leapfrog.cpp:195-199):
// leapfrogTranslate only (:275-278), plus the Verlet test of their Aabb (Collision.cpp:113-121).  setLast: record Aabb::nodeLastPos at a full run.
__global__ void k_walls_step(Wall* walls, int n, StepU u, Counters* cnt, int setLast, const ImpSlot* imp) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  Wall W = walls[w];
  if (setLast) { for (int k = 0; k < 3; k++) walls[w].lastPos[k] = W.pos[k]; return; }
  if (!W.moving || *u.gIn) return;
  d3 pos = mk3(W.pos[0], W.pos[1], W.pos[2]), vel = mk3(W.vel[0], W.vel[1], W.vel[2]);
  if (u.periodic) {      // a fully blocked node in a deforming cell: Leapfrog's fixed branch corrects its velocity / position like any node's (Leapfrog.cpp:206-215,258-273)
    if (u.homoDeform == 4) vel = mulMV(u.ImLL4hInv, mulMV(u.LmL, pos) + mulMV(u.IpLL4h, vel));
    else if (u.homoDeform == 2 || u.homoDeform == 3) {
      if (u.homoDeform == 3) vel = vel + mulMV(u.midGradV, vel - (u.dt / 2.) * mk3(0, 0, 0)) * u.dt;
      vel = vel + mulMV(u.dGradV, pos);
    } else if (u.homoDeform == 1) pos = pos + mulMV(u.nextGradV, pos) * u.dt;
    walls[w].vel[0] = vel.x; walls[w].vel[1] = vel.y; walls[w].vel[2] = vel.z;
  }
  pos = pos + vel * u.dt;
  walls[w].pos[0] = pos.x; walls[w].pos[1] = pos.y; walls[w].pos[2] = pos.z;
  if (imp && w < 32 && imp->wallHas[w]) for (int k = 0; k < 3; k++) walls[w].vel[k] = imp->wallVel[3 * w + k];      // Impose::velocity at the end of the node's step (Leapfrog.cpp:248)
  const d3 dl = mk3(W.lastPos[0], W.lastPos[1], W.lastPos[2]) - pos;
  const double d2 = fmax(0., sqn(dl));
  if (d2 > u.maxD2 || !(d2 == d2)) *u.gOut = 1;
}

// ------------------------------------------------------------------------------------------------
// Shared tail of the three Cg2 functors: Cg2_Any_Any_L6Geom__Base::handleSpheresLikeContact
// (src/dem/L6Geom.cpp:49-216, approxMask==0) -> Cp2 (once) -> Law2 -> global force output.
// ------------------------------------------------------------------------------------------------
// who to name in a device error report; ids are looked up only on the (cold) error path
struct ErrCtx { const int32_t* sid; int slotA; int idA; int slotB; };
__device__ __noinline__ void report_error(Counters* cnt, int code, const ErrCtx& ec) {
  cnt->error = code; cnt->errA = ec.slotA >= 0 ? ec.sid[ec.slotA] : ec.idA; cnt->errB = ec.sid[ec.slotB];
}
// persistent state of a real contact, loaded up front so that all of a thread's loads are in flight together
struct Hist { d3 t0, t1; double4 cp; double ft0, ft1, ph0, ph1, ph2, ph3, ph4; };
template <int LAW>
__device__ __forceinline__ void load_hist(const Contacts& C, long long c, Hist& H) {
  const double4 a = C.rA[c], b = C.rB[c], p = C.ph[c];
  H.cp = C.rC[c];
  H.t0 = xyz(a); H.ft0 = a.w; H.t1 = xyz(b); H.ft1 = b.w;
  H.ph0 = p.x; H.ph1 = p.y; H.ph2 = p.z;
  if ((LAW & 1) == 1) { H.ph3 = p.w; H.ph4 = C.ph4[c]; }
}

template <int LAW>
__device__ __forceinline__ void contact_tail(const Contacts& C, long long c, uint32_t fl, const StepU& u, d3 pos1, d3 vel1, d3 w1, d3 pos2,
                                             d3 vel2, d3 w2, d3 shiftVel2, d3 normal, d3 contPt, double uN, double r1, double r2,
                                             const Material* mats, const int32_t* matArr, double massA, double massB, Counters* cnt,
                                             const ErrCtx& ec, int matBoundary, const Hist& H, unsigned& nRealOut, const bool mine = true) {
  // mine (energy tracking in decomposed runs): a cross-slab pair is evaluated on both sides; the side owning pA books its energy
  const bool wasReal = fl & CF_REAL;
  d3 c0, c1, c2, gvel = mk3(0, 0, 0), gang = mk3(0, 0, 0);
  double kn, kt, tanPhi, p3 = 0, p4 = 0, ftA = 0, ftB = 0, tq0 = 0, tq1 = 0, tq2 = 0;
  if (!wasReal) {
    // ---- create geometry (L6Geom.cpp:51-70) + L6Geom::setInitialLocalCoords (L6Geom.cpp:25-36)
    c0 = normal;
    d3 ly = cross(c0, fabs(c0.y) < fabs(c0.z) ? mk3(0, 1, 0) : mk3(0, 0, 1));
    ly = ly - c0 * dot(ly, c0);
    ly = ly / nrm(ly);
    c1 = ly;
    c2 = cross(c0, c1);
    // iniLensTouch=true: lens=(|r1|,|r2|); contA = pi*min(r)^2 (or the positive one)
    double l0 = fabs(r1), l1 = fabs(r2);
    double rr = (r1 > 0 && r2 > 0) ? fmin(r1, r2) : (r1 > 0 ? r1 : r2);
    double A = M_PI * (rr * rr);
    // ---- Cp2 (ContactLoop.cpp:179-181)
    const Material m1 = mats[ec.slotA >= 0 ? matArr[ec.slotA] : matBoundary], m2 = mats[matArr[ec.slotB]];
    if ((LAW & 1) == 0) {  // Cp2_FrictMat_FrictPhys::updateFrictPhys (FrictMat.cpp:20-44)
      kn = 1 / (1 / (m1.young * A / l0) + 1 / (m2.young * A / l1));
      kt = .5 * (m1.ktDivKn + m2.ktDivKn) * kn;
      tanPhi = fmin(m1.tanPhi, m2.tanPhi);
      C.ph[c] = make_double4(kn, kt, tanPhi, 0.);
    } else {         // Cp2_HertzMat_HertzPhys::go (Hertz.cpp:34-77)
      if (m1.kind != 1 || m2.kind != 1) { report_error(cnt, ERR_NO_PHYS, ec); return; }
      const double nu1 = u.poisson, nu2 = u.poisson, E1 = m1.young, E2 = m2.young;
      double K = (4 / 3.) * 1. / ((1 - nu1 * nu1) / E1 + (1 - nu2 * nu2) / E2);
      double R = 1. / (1. / l0 + 1. / l1);
      double kn0 = K * sqrt(R);
      double gamma = fmin(m1.surfEnergy, m2.surfEnergy);
      double alpha = .5 * (m1.alpha + m2.alpha);
      if (!(alpha == 0. || gamma == 0.)) { report_error(cnt, ERR_SCHWARZ, ec); return; }
      double G1 = E1 / (2 * (1 + nu1)), G2 = E2 / (2 * (1 + nu2));
      double G = .5 * (G1 + G2);
      double nu = .5 * (nu1 + nu2);
      double kt0 = 2 * sqrt(4 * R) * G / (2 - nu);
      tanPhi = fmin(m1.tanPhi, m2.tanPhi);
      double asmk = 0.0;
      if (u.en > 0 && u.en < 1.) {
        double mbar = (massA <= 0 && massB > 0) ? massB : ((massA > 0 && massB <= 0) ? massA : (massA * massB) / (massA + massB));
        asmk = fmax(0., u.viscAlpha * sqrt(mbar * kn0));
      }
      // ph[4] = Fna = 4*pi*R*gamma (Hertz.cpp:113) when alpha==0; alpha!=0 implies gamma==0 here (Fna=0) and is flagged by -1:
      // the noAttraction clamp applies only when alpha==0 (Hertz.cpp:177)
      kn = kn0; kt = kt0; p3 = asmk; p4 = (alpha == 0.) ? (4 * M_PI * R * gamma) : -1.0;
      C.ph[c] = make_double4(kn0, kt0, tanPhi, asmk); C.ph4[c] = p4;
    }
    C.stepCreated[c] = (int)u.step;
    atomicAdd(&cnt->nCreated, 1ull);
  } else {
    // ---- update geometry (L6Geom.cpp:72-216)
    const d3 prevN = H.t0;
    const d3 prev1 = H.t1;
    const d3 prevContPt = xyz(H.cp);
    const double dt = u.dt;
    d3 normRotVec = cross(prevN, normal);
    d3 midN = .5 * (prevN + normal);
    double nm = nrm(midN);
    if (nm == 0) midN = normal; else midN = midN / nm;
    d3 normTwistVec = midN * dt * .5 * dot(midN, w1 + w2);
    d3 rot = normRotVec + normTwistVec;
    d3 mid1 = prev1 - cross(prev1, rot) / 2.;
    d3 mid2 = cross(midN, mid1);
    c0 = normal;
    c1 = prev1 - cross(mid1, rot);
    c2 = cross(c0, c1);
    if (u.renormNow) {                                                     // L6Geom.cpp:124-129
      c0 = c0 / nrm(c0);
      c1 = c1 - c0 * dot(c1, c0);
      c1 = c1 / nrm(c1);
      c2 = cross(c0, c1);
    }
    d3 midContPt = .5 * (prevContPt + contPt);
    d3 midPos1 = pos1 - (dt / 2.) * vel1;
    d3 v2s = vel2 + shiftVel2;
    d3 midPos2 = pos2 - (dt / 2.) * v2s;
    d3 c1x = (u.noRatch && r1 > 0) ? (r1 * midN) : (midContPt - midPos1);
    d3 c2x = (u.noRatch && r2 > 0) ? ((-r2) * midN) : (midContPt - midPos2);
    d3 relVel = (v2s + cross(w2, c2x)) - (vel1 + cross(w1, c1x));
    gvel = mk3(dot(midN, relVel), dot(mid1, relVel), dot(mid2, relVel));
    if (u.rolling) { d3 dw = w2 - w1; gang = mk3(dot(midN, dw), dot(mid1, dw), dot(mid2, dw)); }
    kn = H.ph0; kt = H.ph1; tanPhi = H.ph2;
    if ((LAW & 1) == 1) { p3 = H.ph3; p4 = H.ph4; }
    ftA = H.ft0; ftB = H.ft1;
    // CPhys.torque is zeroed at the top of the law every step (IdealElPl.cpp:43), so it carries no history
  }
  // ---- Law2
  double Fn; bool keep = true;
  if ((LAW & 1) == 0) {  // Law2_L6Geom_FrictPhys_IdealElPl::go (IdealElPl.cpp:15-125), iniEqlb/trackEnergy unsupported
    if (uN > 0 && !u.noBreak) {
      keep = false;
      if (LAW >= 2 && mine && u.noSlip) {   // IdealElPl.cpp:28-40: the elastic potential of the breaking contact is lost
        const double FnB = kn * uN, t0 = ftA + (u.dt * kt) * gvel.y, t1 = ftB + (u.dt * kt) * gvel.z;
        energy_add(u.energy, EN_BROKEN, .5 * ((FnB * FnB) / kn + (t0 * t0 + t1 * t1) / kt));
      }
    } else {
      Fn = kn * uN;
      if (u.noFrict || tanPhi == 0.) { ftA = 0; ftB = 0; }
      else {
        ftA += (u.dt * kt) * gvel.y; ftB += (u.dt * kt) * gvel.z;
        double maxFt = fabs(Fn) * tanPhi;
        double sq = ftA * ftA + ftB * ftB;
        if (sq > maxFt * maxFt && !u.noSlip) {
          double FtNorm = sqrt(sq);
          double ratio = maxFt / FtNorm;
          if (FtNorm == 0.) ratio = 0.;
          if (LAW >= 2 && mine) energy_add(u.energy, EN_PLAST, kt != 0. ? (maxFt) * (FtNorm - maxFt) / kt : 0.);   // IdealElPl.cpp:64-75
          ftA *= ratio; ftB *= ratio;
        }
        if (!(Fn == Fn) || !(ftA == ftA) || !(ftB == ftB)) report_error(cnt, ERR_NAN_FORCE, ec);
      }
      if (LAW >= 2 && mine) {                                              // IdealElPl.cpp:85-92
        double elast = 0.5 * ((Fn * Fn) / kn + (kt != 0. ? (ftA * ftA + ftB * ftB) / kt : 0.));
        energy_add(u.energy, EN_ELAST, elast);
      }
      if (u.rolling) {                                                     // IdealElPl.cpp:94-123
        double charLen = fabs(r1) + fabs(r2);
        double kr = kn * charLen;
        if (u.relTwistStiff > 0) {
          tq0 += u.dt * u.relTwistStiff * kr * gang.x;
          double maxTt = fabs(Fn * u.rollTanPhi * charLen);
          if (fabs(tq0) > maxTt) tq0 = copysign(maxTt, tq0);
        }
        tq1 += (u.dt * kr) * gang.y; tq2 += (u.dt * kr) * gang.z;
        double maxTr = fmax(0., fabs(Fn) * u.rollTanPhi * charLen);
        double sq = tq1 * tq1 + tq2 * tq2;
        if (sq > maxTr * maxTr) { double TrNorm = sqrt(sq); double ratio = maxTr / TrNorm; if (TrNorm == 0.) ratio = 0.; tq1 *= ratio; tq2 *= ratio; }
      }
    }
  } else {         // Law2_L6Geom_HertzPhys_DMT::go, Hertz/DMT branch (Hertz.cpp:85-114,170-210)
    const double kn0 = kn, kt0 = kt, asmk = p3;
    const bool alphaZero = !(p4 < 0);
    const double Fna = alphaZero ? p4 : 0.;
    if (uN > 0) keep = false;
    else {
      double s = sqrt(-uN);
      double Fne = -kn0 * ((s * s) * s);
      double eta = (asmk > 0 ? asmk * sqrt(sqrt(fmax(-uN, 0.))) : 0.);
      double Fnv = eta * gvel.x;
      if (alphaZero && u.noAttraction && Fne + Fnv > 0) Fnv = -Fne;
      Fn = Fne + Fna + Fnv;
      if (LAW >= 2 && mine) energy_add(u.energy, EN_VISCN, Fnv * gvel.x * u.dt);                      // Hertz.cpp:181
      double ktc = kt0 * sqrt(uN < 0 ? -uN : 0);
      ftA += (u.dt * ktc) * gvel.y; ftB += (u.dt * ktc) * gvel.z;
      double maxFt = fabs(fmin(0., Fn) * tanPhi);
      double sq = ftA * ftA + ftB * ftB;
      if (sq > maxFt * maxFt) {
        double FtNorm = sqrt(sq); double ratio = maxFt / FtNorm;
        if (LAW >= 2 && mine) energy_add(u.energy, EN_PLAST, (.5 * (FtNorm - maxFt) + maxFt) * (FtNorm - maxFt) / ktc);   // Hertz.cpp:193
        ftA *= ratio; ftB *= ratio;
      } else {
        ftA += eta * gvel.y; ftB += eta * gvel.z;
        if (LAW >= 2 && mine) energy_add(u.energy, EN_VISCT, eta * (gvel.y * gvel.y + gvel.z * gvel.z) * u.dt);           // Hertz.cpp:199
      }
      if (LAW >= 2 && mine && uN < 0) {                                                               // Hertz.cpp:203-207, Hertz.hpp:70,76
        const double sq5 = sqrt(-uN);
        energy_add(u.energy, EN_ELAST, kn0 * (2 / 5.) * (((sq5 * sq5) * (sq5 * sq5)) * sq5) + 0.5 * (ftA * ftA + ftB * ftB) / ktc);
      }
      if (!(Fn == Fn) || !(ftA == ftA) || !(ftB == ftB)) report_error(cnt, ERR_NAN_FORCE, ec);
    }
  }
  if (!keep) {
    // ContactContainer::requestRemoval -> Contact::reset (ContactContainer.cpp:164-174): geom/phys dropped; the pair stays a potential
    // contact while its (frozen) AABBs overlap, i.e. unless it is an orphan (InsertionSortCollider.hpp:213-218)
    uint32_t nf = (fl & ~CF_REAL) | ((fl & CF_ORPHAN) ? CF_DEAD : 0);
    C.flags[c] = nf;
    C.rD[c] = make_double4(0, 0, 0, 0);
    if (wasReal) { atomicAnd(&C.realBits[c >> 5], ~(1u << (c & 31))); atomicAdd(&cnt->nBroken, 1ull); }
    return;
  }
  // ---- store state
  C.rA[c] = make_double4(c0.x, c0.y, c0.z, ftA);
  C.rB[c] = make_double4(c1.x, c1.y, c1.z, ftB);
  C.rC[c] = make_double4(contPt.x, contPt.y, contPt.z, uN);
  if (!wasReal) { C.flags[c] = fl | CF_REAL; atomicOr(&C.realBits[c >> 5], 1u << (c & 31)); }
  // ---- Contact::getForceTorqueBranch (Contact.cpp:24-35): F = ori * phys.force, ori = Quaternionr(trsf)
  q4 q = quatFromCols(c0, c1, c2);
  d3 Fglob = quatRotate(q, mk3(Fn, ftA, ftB));
  C.rD[c] = make_double4(Fglob.x, Fglob.y, Fglob.z, Fn);
  if (u.rolling) {
    C.tq[c] = make_double4(tq0, tq1, tq2, 0.);
    d3 Tg = (tq0 == 0. && tq1 == 0. && tq2 == 0.) ? mk3(0, 0, 0) : quatRotate(q, mk3(tq0, tq1, tq2));
    C.Tg[c] = make_double4(Tg.x, Tg.y, Tg.z, 0);
  }
  nRealOut = 1;
}

__device__ __forceinline__ void block_count(unsigned isReal, unsigned evald, Counters* cnt) {
  unsigned r = __reduce_add_sync(0xffffffffu, isReal), e = __reduce_add_sync(0xffffffffu, evald);
  if ((threadIdx.x & 31) == 0) { if (r) atomicAdd(&cnt->nReal, (unsigned long long)r); if (e) atomicAdd(&cnt->nPotEval, (unsigned long long)e); }
}

// ------------------------------------------------------------------------------------------------
// ContactLoop::run body for Sphere+Sphere (src/dem/ContactLoop.cpp:134-216) with Cg2_Sphere_Sphere_L6Geom::go
// (src/dem/Sphere.cpp:92-117) fused with Cp2 and Law2.  One thread per candidate pair.
// ------------------------------------------------------------------------------------------------
template <int LAW>
__device__ __forceinline__ void contact_ss(const Contacts& C, long long c, uint32_t fl, int2 pr, const Spheres& S, const Material* mats, const StepU& u, Counters* cnt, unsigned& isReal,
                                           const Hist& H) {
  const bool real = fl & CF_REAL;
  // every load of a real contact is issued up front (the history by the caller: straight from global memory or from the staged tile), before the first use
  const double4 pa = S.posr[pr.x], pb = S.posr[pr.y];
  double4 va = make_double4(0, 0, 0, 0), vb = va, wa = va, wb = va;
  if (real) { va = S.velm[pr.x]; vb = S.velm[pr.y]; wa = S.angi[pr.x]; wb = S.angi[pr.y]; }
  d3 shift2 = mk3(0, 0, 0), shiftVel2 = mk3(0, 0, 0);
  if (u.periodic) {
    d3 cd = mk3((double)C.cdx[c], (double)C.cdy[c], (double)C.cdz[c]);
    shift2 = mulMV(u.hSize, cd); shiftVel2 = mulMV(u.shiftVelM, cd);
  }
  const d3 p1 = xyz(pa), p2 = xyz(pb) + shift2;
  const double r1 = pa.w, r2 = pb.w;
  d3 relPos = p2 - p1;
  double d2 = sqn(relPos);
  double rs = u.absDistFactor * (r1 + r2);
  double unDistSq = d2 - rs * rs;
  if (unDistSq > 0 && !real) return;
  double dist = sqrt(d2);
  double uN = dist - (r1 + r2);
  d3 normal = relPos / dist;
  d3 contPt = p1 + (r1 + 0.5 * uN) * normal;
  double mA = 0, mB = 0;
  if (!real && (LAW & 1) == 1) { mA = S.velm[pr.x].w; mB = S.velm[pr.y].w; }
  const ErrCtx ec{S.id, pr.x, 0, pr.y};
  const bool mine = (LAW >= 2) ? !(S.flags[pr.x] & F_GHOST) : true;
  contact_tail<LAW>(C, c, fl, u, p1, xyz(va), xyz(wa), p2, xyz(vb), xyz(wb), shiftVel2, normal, contPt, uN, r1, r2, mats, S.mat, mA, mB, cnt, ec, 0, H, isReal, mine);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2 prefetch of everything contact_ss<LAW> will load for contact c (its table entry is already in registers)
template <int LAW>
__device__ __forceinline__ void prefetch_contact_ss(const Contacts& C, long long c, uint32_t fl, int2 pr, const Spheres& S) {
  if (fl & CF_DEAD) return;
  prefetch_l2(S.posr + pr.x); prefetch_l2(S.posr + pr.y);
  if (!(fl & CF_REAL)) return;
  prefetch_l2(S.velm + pr.x); prefetch_l2(S.velm + pr.y); prefetch_l2(S.angi + pr.x); prefetch_l2(S.angi + pr.y);
  prefetch_l2(C.rA + c); prefetch_l2(C.rB + c); prefetch_l2(C.rC + c); prefetch_l2(C.ph + c);
  if ((LAW & 1) == 1) prefetch_l2(C.ph4 + c);
}

template <int LAW>
__device__ __forceinline__ void prefetch_spheres_ss(uint32_t fl, int2 pr, const Spheres& S) {
  if (fl & CF_DEAD) return;
  prefetch_l2(S.posr + pr.x); prefetch_l2(S.posr + pr.y);
  if (!(fl & CF_REAL)) return;
  prefetch_l2(S.velm + pr.x); prefetch_l2(S.velm + pr.y); prefetch_l2(S.angi + pr.x); prefetch_l2(S.angi + pr.y);
}

// Persistent software-pipelined grid (one CTA per resident slot, tiles of 128 candidates, stride gridDim): while tile k is computed, the table
// entries (flags, pair) of tile k+2 are in flight into registers and the sphere/history lines of tile k+1 are being pulled into L2, so the
// dependent chain flags+pair -> gathers -> math of a tile only ever waits for L2, not for HBM.  Real-first order (k_carry) + round-robin tiles
// balance the long (real) and short (potential) tiles over the CTAs.
// 3 CTAs of 128 threads per SM = 168 registers per thread, no spills: measured 0.187 ms per step at 1M spheres (2.07M real + 0.35M active potential contacts)
// against 0.211 ms at 4 CTAs (128 registers, 230 bytes of spills per thread); round 1 measured 5 / 6 CTAs (96 / 80 registers) at 0.264 / 0.336 ms
#ifndef WD_CONTACT_CTAS_PER_SM
#define WD_CONTACT_CTAS_PER_SM 3
#endif
#ifndef WD_CONTACT_BLOCK
#define WD_CONTACT_BLOCK 128
#endif
template <int LAW>
__global__ void __launch_bounds__(WD_CONTACT_BLOCK, WD_CONTACT_CTAS_PER_SM) k_contact_ss(Contacts C, long long first, long long n, Spheres S, const Material* mats, StepU u, Counters* cnt, const int* __restrict__ guard) {
  const int stale = *guard;   // issued together with the loads below; a set flag means a collider full run is due and this speculative launch is void
  const long long end = first + n, stride = (long long)gridDim.x * blockDim.x;
  long long c0 = first + (long long)blockIdx.x * blockDim.x + threadIdx.x, c1 = c0 + stride;
  uint32_t fl0 = CF_DEAD, fl1 = CF_DEAD; int2 pr0 = make_int2(0, 0), pr1 = pr0;
  if (c0 < end) { fl0 = C.flags[c0]; pr0 = C.pair[c0]; }
  if (c1 < end) { fl1 = C.flags[c1]; pr1 = C.pair[c1]; }
  if (stale) return;          // uniform over the grid
  unsigned isReal = 0, evald = 0;
  // whole tiles: the loop condition is uniform over the CTA (tile base), lanes beyond the end carry CF_DEAD
  for (long long base = first + (long long)blockIdx.x * blockDim.x; base < end; base += stride) {
    const long long c2 = c1 + stride;
    uint32_t fl2 = CF_DEAD; int2 pr2 = make_int2(0, 0);
    if (c2 < end) { fl2 = C.flags[c2]; pr2 = C.pair[c2]; }
    prefetch_contact_ss<LAW>(C, c1, fl1, pr1, S);
    if (!(fl0 & CF_DEAD)) {
      evald++; unsigned r = 0; Hist H;
      if (fl0 & CF_REAL) load_hist<LAW>(C, c0, H);
      contact_ss<LAW>(C, c0, fl0, pr0, S, mats, u, cnt, r, H); isReal += r;
    }
    c0 = c1; fl0 = fl1; pr0 = pr1; c1 = c2; fl1 = fl2; pr1 = pr2;
  }
  block_count(isReal, evald, cnt);
}

// ------------------------------------------------------------------------------------------------
// Force / torque on walls that are NOT in DemField.nodes (at rest when added: ParticleContainer.cpp:193-198).  Leapfrog never visits them, so nothing ever resets
// their DemData.force: ContactLoop::applyForceUninodal keeps adding, step after step (ContactLoop.cpp:270-276) — scripts read wall.f as an accumulated
// value.  The wall+sphere kernels (k_contact_ws, k_contact_orphans) do it on the fly: every warp adds up the contacts of each wall it holds (butterfly shuffle,
// fixed order), the block leaves one partial sum per wall, and the last block to finish (ticket) adds the partials in block order into the accumulator.  The
// reference's own order of these additions is its contact-loop order; agreement is to rounding (tests: 1e-9); the walls do not move, so nothing feeds back.
// At most WD_ACC_WALLS walls take part (the rest report the last step's contacts like a moving wall).
// ------------------------------------------------------------------------------------------------
#define WD_ACC_WALLS 32
struct WallAcc {   // nWalls == 0: off
  int nWalls; int nBlocksTotal;      // blocks of all wall-contact launches of this step (the ticket counts them)
  int blockBase;                     // first partial row of this launch
  int foldInKernel;                  // few blocks: the last one to finish books the partial sums; many blocks: k_wall_fold does it after the force reduction
  double* partial /*[nBlocksTotal][nWalls][6]*/; double* acc /*[nWalls][6]*/; unsigned* ticket;
};
// called by ALL threads of a 128-thread block; (w, F, xc x F + T) of this thread's contact if it is a real contact on a wall at rest, w < 0 otherwise
__device__ __forceinline__ void wall_acc_block(const WallAcc& wa, int w, d3 F, d3 T) {
  __shared__ double sW[4][WD_ACC_WALLS][6];
  __shared__ unsigned last;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int q = lane; q < wa.nWalls * 6; q += 32) sW[wp][q / 6][q % 6] = 0.;
  __syncwarp();
  unsigned todo = __ballot_sync(0xffffffffu, w >= 0);
  while (todo) {      // one round per wall present in this warp
    const int cur = __shfl_sync(0xffffffffu, w, __ffs(todo) - 1);
    const bool mine = (w == cur);
    double v[6] = {mine ? F.x : 0., mine ? F.y : 0., mine ? F.z : 0., mine ? T.x : 0., mine ? T.y : 0., mine ? T.z : 0.};
#pragma unroll
    for (int k = 0; k < 6; k++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if (lane == 0) for (int k = 0; k < 6; k++) sW[wp][cur][k] = v[k];
    todo &= ~__ballot_sync(0xffffffffu, mine);
  }
  __syncthreads();
  double* row = wa.partial + (long long)(wa.blockBase + blockIdx.x) * wa.nWalls * 6;
  for (int q = threadIdx.x; q < wa.nWalls * 6; q += blockDim.x) row[q] = ((sW[0][q / 6][q % 6] + sW[1][q / 6][q % 6]) + sW[2][q / 6][q % 6]) + sW[3][q / 6][q % 6];
  if (!wa.foldInKernel) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(wa.ticket, 1u);
  __syncthreads();
  if (last != (unsigned)(wa.nBlocksTotal - 1)) return;
  __threadfence();
  // the last block: G groups of threads share the block range of every value, partial sums combined in group order (fixed order, a few hundred blocks)
  __shared__ double red[128];
  const int nv = wa.nWalls * 6, G = max(1, (int)blockDim.x / nv), nB = wa.nBlocksTotal, per = (nB + G - 1) / G;
  for (int q0 = 0; q0 < nv; q0 += blockDim.x) {      // one pass unless there are more than 21 walls
    const int q = q0 + (int)threadIdx.x % min(nv, (int)blockDim.x), g = (int)threadIdx.x / min(nv, (int)blockDim.x);
    double sum = 0;
    if (g < G && q < nv) {
      const int b0 = g * per, b1 = min(nB, b0 + per);
      int b = b0;
      for (; b + 4 <= b1; b += 4) {
        const double a0 = wa.partial[(long long)b * nv + q], a1 = wa.partial[(long long)(b + 1) * nv + q], a2 = wa.partial[(long long)(b + 2) * nv + q], a3 = wa.partial[(long long)(b + 3) * nv + q];
        sum = (((sum + a0) + a1) + a2) + a3;
      }
      for (; b < b1; b++) sum += wa.partial[(long long)b * nv + q];
    }
    __syncthreads();
    if (threadIdx.x < 128) red[threadIdx.x] = sum;
    __syncthreads();
    if (g == 0 && q < nv) {
      double tot = 0;
      for (int k = 0; k < G; k++) tot += red[k * min(nv, (int)blockDim.x) + (q - q0)];
      wa.acc[q] += tot;
    }
  }
  if (threadIdx.x == 0) *wa.ticket = 0;
}
// many partial rows (large scenes): one block per value adds its column up with a fixed tree and books it
__global__ void __launch_bounds__(256) k_wall_fold(const double* __restrict__ partial, int nB, int nv, double* acc, const int* __restrict__ guard) {
  if (*guard) return;
  const int q = blockIdx.x;
  double sum = 0;
  for (int b = threadIdx.x; b < nB; b += 256) sum += partial[(long long)b * nv + q];
  __shared__ double sm[256];
  sm[threadIdx.x] = sum; __syncthreads();
  for (int s_ = 128; s_ > 0; s_ >>= 1) { if ((int)threadIdx.x < s_) sm[threadIdx.x] += sm[threadIdx.x + s_]; __syncthreads(); }
  if (threadIdx.x == 0) acc[q] += sm[0];
}
// this thread's contribution after its wall+sphere contact was evaluated (reads back what the tail just stored)
__device__ __forceinline__ void wall_acc_contact(const WallAcc& wa, const Contacts& C, long long c, bool valid, int2 pr, const Wall* walls, const Spheres& S, const StepU& u, int kindIsWS) {
  int w = -1; d3 F = mk3(0, 0, 0), T = F;
  if (valid && kindIsWS && pr.x < wa.nWalls) {
    const uint32_t fl = C.flags[c];
    if ((fl & CF_REAL) && !(fl & CF_DEAD) && !walls[pr.x].moving && !(S.flags[pr.y] & F_GHOST)) {
      w = pr.x;
      F = xyz(C.rD[c]);                                                                 // wall = pA: sign +1 (Contact.cpp:28-31)
      const d3 xc = xyz(C.rC[c]) - mk3(walls[w].pos[0], walls[w].pos[1], walls[w].pos[2]);
      T = cross(xc, F);
      if (u.rolling) T = T + xyz(C.Tg[c]);
    }
  }
  wall_acc_block(wa, w, F, T);
}

// ------------------------------------------------------------------------------------------------
// Wall+Sphere: Cg2_Wall_Sphere_L6Geom::go (src/dem/Wall.cpp:50-84)
// ------------------------------------------------------------------------------------------------
template <int LAW>
__device__ __forceinline__ void contact_ws(const Contacts& C, long long c, uint32_t fl, int2 pr, const Spheres& S, const Wall* walls, const Material* mats, const StepU& u, Counters* cnt, unsigned& isReal) {
  const bool real = fl & CF_REAL;
  const Wall W = walls[pr.x];
  const double4 pb = S.posr[pr.y];
  double4 vb = make_double4(0, 0, 0, 0), wb = vb; Hist H;
  if (real) { vb = S.velm[pr.y]; wb = S.angi[pr.y]; load_hist<LAW>(C, c, H); }
  d3 shift2 = mk3(0, 0, 0), shiftVel2 = mk3(0, 0, 0);
  if (u.periodic) {
    d3 cd = mk3((double)C.cdx[c], (double)C.cdy[c], (double)C.cdz[c]);
    shift2 = mulMV(u.hSize, cd); shiftVel2 = mulMV(u.shiftVelM, cd);
  }
  const d3 wallPos = mk3(W.pos[0], W.pos[1], W.pos[2]);
  const d3 spherePos = xyz(pb) + shift2;
  const double radius = pb.w; const int ax = W.axis;
  double dist = comp(spherePos, ax) - comp(wallPos, ax);
  if (!real && fabs(dist) > radius) return;
  d3 contPt = spherePos; setc(contPt, ax, comp(wallPos, ax));
  d3 normal = mk3(0, 0, 0);
  double nax;
  if (W.sense == 0) {
    if (!real) nax = dist > 0 ? 1. : -1.;
    else nax = comp(H.t0, ax);
  } else nax = (W.sense == 1 ? 1. : -1.);
  setc(normal, ax, nax);
  double uN = nax * dist - radius;
  double mB = 0;
  if (!real && (LAW & 1) == 1) mB = S.velm[pr.y].w;
  const ErrCtx ec{S.id, -1, W.id, pr.y};
  contact_tail<LAW>(C, c, fl, u, wallPos, mk3(W.vel[0], W.vel[1], W.vel[2]), mk3(0, 0, 0), spherePos, xyz(vb), xyz(wb), shiftVel2, normal, contPt, uN, -radius, radius,
                    mats, S.mat, 0., mB, cnt, ec, W.mat, H, isReal);
}

template <int LAW>
__global__ void __launch_bounds__(128) k_contact_ws(Contacts C, long long first, long long n, Spheres S, const Wall* walls, const Material* mats, StepU u, Counters* cnt, const int* __restrict__ guard, WallAcc wa) {
  long long c = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned isReal = 0, evald = 0;
  const int stale = *guard;   // issued together with the loads below; a set flag means a collider full run is due and this speculative launch is void
  uint32_t fl = CF_DEAD; int2 pr = make_int2(0, 0);
  if (c < first + n) { fl = C.flags[c]; pr = C.pair[c]; }
  if (stale) return;          // uniform over the grid
  {
    if (!(fl & CF_DEAD)) { evald = 1; contact_ws<LAW>(C, c, fl, pr, S, walls, mats, u, cnt, isReal); }
  }
  block_count(isReal, evald, cnt);
  if (wa.nWalls) wall_acc_contact(wa, C, c, isReal != 0, pr, walls, S, u, 1);
}

// ------------------------------------------------------------------------------------------------
// Facet+Sphere: Cg2_Facet_Sphere_L6Geom::go (src/dem/Facet.cpp:333-391) with Facet::getNormal/getOuterVectors (:74-77,:113-118),
// interpolatePtLinAngVel (:161-177), adjustBoundaryContactGeom(_impl) (:220-265), CompUtils::closestSegmentPt /
// triangleBarycentrics (src/supp/base/CompUtils.cpp:59-65,94-112).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ d3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ d3 closestSegmentPt(d3 P, d3 A, d3 B) {
  d3 BA = B - A;
  if (sqn(BA) == 0.) return A;
  double u = (dot(P, BA) - dot(A, BA)) / (sqn(BA));
  return A + fmin(1., fmax(0., u)) * BA;
}
__device__ __forceinline__ d3 triangleBarycentrics(d3 x, d3 A, d3 B, d3 Cc) {
  d3 vB = B - A, vC = Cc - A, vX = x - A;
  double m00 = dot(vB, vB), m01 = dot(vC, vB), m10 = dot(vB, vC), m11 = dot(vC, vC);
  double invdet = 1. / (m00 * m11 - m10 * m01);   // Eigen 2x2: inverse = adj * (1/det)
  double i00 = m11 * invdet, i10 = -m10 * invdet, i01 = -m01 * invdet, i11 = m00 * invdet;
  double b0 = dot(vX, vB), b1 = dot(vX, vC);
  double u = i00 * b0 + i01 * b1, v = i10 * b0 + i11 * b1;
  return mk3(1 - u - v, u, v);
}
__device__ __forceinline__ void facetAdjustImpl(d3 fN, d3 edgePt, d3 unitOutVec, d3 otherCenter, double nnLim, d3& normal, d3& contPt) {
  if (nnLim != nnLim) return;
  double nn = dot(normal, fN);
  if ((nn >= 0 && nnLim > 0 && nn >= nnLim) || (nn < 0 && nnLim < 0 && nn <= nnLim)) return;
  double noLim = (fabs(nnLim) >= 1. ? 1. : sqrt(1. - nnLim * nnLim));
  double nno0, nno1;
  if ((nnLim >= 0 && nn <= -noLim) || (nnLim <= 0 && nn >= noLim)) { nno0 = -nnLim; nno1 = 0; } else { nno0 = nnLim; nno1 = noLim; }
  normal = nno0 * fN + nno1 * unitOutVec;
  contPt = otherCenter - normal * (dot(otherCenter - edgePt, normal));
}

template <int LAW>
__device__ __forceinline__ void contact_fs(const Contacts& C, long long c, uint32_t fl, int2 pr, const Spheres& S, const Facet* facets, const Material* mats, const StepU& u, Counters* cnt, unsigned& isReal) {
  const Facet* F = facets + pr.x;
  const double4 pb = S.posr[pr.y];
  double4 vb = make_double4(0, 0, 0, 0), wb = vb; Hist H;
  if (fl & CF_REAL) { vb = S.velm[pr.y]; wb = S.angi[pr.y]; load_hist<LAW>(C, c, H); }
  d3 shift2 = mk3(0, 0, 0), shiftVel2 = mk3(0, 0, 0);
  if (u.periodic) {
    d3 cd = mk3((double)C.cdx[c], (double)C.cdy[c], (double)C.cdz[c]);
    shift2 = mulMV(u.hSize, cd); shiftVel2 = mulMV(u.shiftVelM, cd);
  }
  const d3 sC = xyz(pb) + shift2;
  const double radius = pb.w, ht = F->halfThick;
  const d3 v0 = ld3(F->v), v1 = ld3(F->v + 3), v2 = ld3(F->v + 6);
  const d3 nn = cross(v1 - v0, v2 - v0);
  const d3 fNormal = nn / nrm(nn);
  double planeDist = dot(sC - v0, fNormal);
  const bool real = fl & CF_REAL;
  if (fabs(planeDist) > (radius + ht) && !real) return;
  d3 fC = sC - planeDist * fNormal;
  d3 o0 = cross(v1 - v0, nn), o1 = cross(v2 - v1, nn), o2 = cross(v0 - v2, nn);
  double l0 = dot(o0, fC - v0), l1 = dot(o1, fC - v1), l2 = dot(o2, fC - v2);
  int w = (l0 > 0 ? 1 : 0) + (l1 > 0 ? 2 : 0) + (l2 > 0 ? 4 : 0);
  d3 contPt;
  switch (w) {
    case 0: contPt = fC; break;
    case 1: contPt = closestSegmentPt(fC, v0, v1); break;
    case 2: contPt = closestSegmentPt(fC, v1, v2); break;
    case 4: contPt = closestSegmentPt(fC, v2, v0); break;
    case 3: contPt = v1; break;
    case 5: contPt = v0; break;
    case 6: contPt = v2; break;
    default: { const ErrCtx ec{S.id, -1, F->id, pr.y}; report_error(cnt, ERR_FACET_IMPOSSIBLE, ec); return; }
  }
  d3 normal = sC - contPt;
  if (sqn(normal) > (radius + ht) * (radius + ht) && !real) return;
  double dist = nrm(normal);
  if (dist != 0) normal = normal / dist;
  else {
    if (real) normal = quatRotate(quatFromCols(H.t0, H.t1, cross(H.t0, H.t1)), mk3(1, 0, 0));
    else normal = mk3(1, 0, 0);
  }
  if (w > 0) {  // Facet::adjustBoundaryContactGeom (Facet.cpp:220-237); edgeNums (Facet.cpp:374)
    const int e0 = (w == 1 || w == 3) ? 0 : ((w == 2 || w == 6) ? 1 : 2);
    const int e1 = (w == 3) ? 1 : ((w == 5) ? 0 : ((w == 6) ? 2 : -1));
    const d3 oe0 = e0 == 0 ? o0 : (e0 == 1 ? o1 : o2);
    const d3 ve0 = e0 == 0 ? v0 : (e0 == 1 ? v1 : v2);
    const double n0 = F->n21lim[e0];
    if (!(n0 != n0)) {
      if (e1 < 0) facetAdjustImpl(fNormal, ve0, oe0 / nrm(oe0), sC, n0, normal, contPt);
      else {
        const double n1 = F->n21lim[e1];
        if (!(n1 != n1)) {
          const d3 oe1 = e1 == 0 ? o0 : (e1 == 1 ? o1 : o2);
          const d3 ve1 = e1 == 0 ? v0 : (e1 == 1 ? v1 : v2);
          d3 out0 = oe0 / nrm(oe0), out1 = oe1 / nrm(oe1);
          double p0 = dot(out0, normal), pMin = dot(out0, out1), pMax = 1.;
          double t01 = (p0 - pMin) / (pMax - pMin);
          facetAdjustImpl(fNormal, ve1, (t01 * out0 + (1 - t01) * out1), sC, (t01 * n0 + (1 - t01) * n1), normal, contPt);
        }
      }
    }
  }
  if (ht > 0) contPt = contPt + normal * ht;
  double uN = dist - radius - ht;
  // Facet::interpolatePtLinAngVel
  const d3 vv0 = ld3(F->vvel), vv1 = ld3(F->vvel + 3), vv2 = ld3(F->vvel + 6), fk = ld3(F->fakeVel);
  d3 a = triangleBarycentrics(contPt, v0, v1, v2);
  d3 linVel = (a.x * vv0 + a.y * vv1) + a.z * vv2;
  d3 angVel = (cross(v0 - contPt, vv0) + cross(v1 - contPt, vv1)) + cross(v2 - contPt, vv2);
  if (!(fk.x == 0. && fk.y == 0. && fk.z == 0.)) {
    if (!(fk.x != fk.x)) linVel = linVel + fk;
    else linVel = dot(linVel, fNormal) * fNormal;
  }
  double mB = 0;
  if (!real && (LAW & 1) == 1) mB = S.velm[pr.y].w;
  const ErrCtx ec{S.id, -1, F->id, pr.y};
  contact_tail<LAW>(C, c, fl, u, contPt, linVel, angVel, sC, xyz(vb), xyz(wb), shiftVel2, normal, contPt, uN, fmax(ht, radius), radius, mats, S.mat, 0., mB, cnt, ec, F->mat, H, isReal);
}

template <int LAW>
__global__ void __launch_bounds__(128) k_contact_fs(Contacts C, long long first, long long n, Spheres S, const Facet* facets, const Material* mats, StepU u, Counters* cnt, const int* __restrict__ guard) {
  long long c = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned isReal = 0, evald = 0;
  const int stale = *guard;   // issued together with the loads below; a set flag means a collider full run is due and this speculative launch is void
  uint32_t fl = CF_DEAD; int2 pr = make_int2(0, 0);
  if (c < first + n) { fl = C.flags[c]; pr = C.pair[c]; }
  if (stale) return;          // uniform over the grid
  {
    if (!(fl & CF_DEAD)) { evald = 1; contact_fs<LAW>(C, c, fl, pr, S, facets, mats, u, cnt, isReal); }
  }
  block_count(isReal, evald, cnt);
}

// Orphans: contacts that are real although their AABBs no longer overlap (the collider never removes a real contact,
// InsertionSortCollider.cpp:78-113); rare, kept in a mixed-kind tail segment.
template <int LAW>
__global__ void __launch_bounds__(128) k_contact_orphans(Contacts C, long long first, long long n, Spheres S, const Wall* walls, const Facet* facets, const Material* mats, StepU u, Counters* cnt, const int* __restrict__ guard, WallAcc wa) {
  long long c = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned isReal = 0, evald = 0;
  const int stale = *guard;   // issued together with the loads below; a set flag means a collider full run is due and this speculative launch is void
  uint32_t fl = CF_DEAD; int2 pr = make_int2(0, 0);
  if (c < first + n) { fl = C.flags[c]; pr = C.pair[c]; }
  if (stale) return;          // uniform over the grid
  {
    if (!(fl & CF_DEAD)) {
      evald = 1;
      const int kind = cf_kind(fl);
      if (kind == KIND_SS) { Hist H; if (fl & CF_REAL) load_hist<LAW>(C, c, H); contact_ss<LAW>(C, c, fl, pr, S, mats, u, cnt, isReal, H); }
      else if (kind == KIND_WS) contact_ws<LAW>(C, c, fl, pr, S, walls, mats, u, cnt, isReal);
      else contact_fs<LAW>(C, c, fl, pr, S, facets, mats, u, cnt, isReal);
    }
  }
  block_count(isReal, evald, cnt);
  if (wa.nWalls) wall_acc_contact(wa, C, c, isReal != 0, pr, walls, S, u, cf_kind(fl) == KIND_WS);
}

// ------------------------------------------------------------------------------------------------
// Per-particle force/torque: ContactLoop::applyForceUninodal + Contact::getForceTorqueBranch + DemData::addForceTorque
// (ContactLoop.cpp:270-276, Contact.cpp:24-35, Particle.hpp:249) as a deterministic segmented reduction over the
// particle's incidence row (ascending id of the other particle), fused with Leapfrog's force reset (Leapfrog.cpp:240-245).
// ------------------------------------------------------------------------------------------------
// One thread per sphere walks its row in order; the loads of RB entries (entry word, global force, contact point) are issued together before the
// sequential adds, so a row costs ~3 dependent round trips per RB entries instead of 3 per entry.
#ifndef WD_REDUCE_RB
#define WD_REDUCE_RB 4
#endif
// 128 threads x 6 CTAs/SM (80 registers): measured best of {256x2 (94 regs) 0.149 ms, 128x5 0.145, 256x3 0.131, 128x6 0.127, 128x7 0.136, 128x8 0.162 (spills)}
#ifndef WD_REDUCE_BLOCK
#define WD_REDUCE_BLOCK 128
#endif
#ifndef WD_REDUCE_MINB
#define WD_REDUCE_MINB 6
#endif
#define RB WD_REDUCE_RB
// Measured alternatives on the 1 M-sphere bed (this kernel: 0.103-0.106 ms): gathers issued without waiting for the real bit 0.112 ms; the rows of a warp walked flat
// (every lane gathers RBK entries of the warp's contiguous entry range into a shared-memory tile, then every lane adds up its own row in order) 0.139 ms (RBK=2) ...
// 0.183 ms (RBK=6); the same block-wide with cp.async gathering 8/16-byte pieces straight into shared memory 0.185-0.206 ms.  The sequential row walk with RB
// independent gathers per pass is the fastest of the four.
__global__ void __launch_bounds__(WD_REDUCE_BLOCK, WD_REDUCE_MINB) k_reduce(Spheres S, int n, Contacts C, const int* __restrict__ rowStart, const uint32_t* __restrict__ rowEntry, StepU u,
                                               const int* __restrict__ guard, ImposeDev I) {
  if (guard && *guard) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t fl = S.flags[i];
  if (fl & F_GHOST) return;   // forces are reduced by the owner only
  const int e0 = rowStart ? rowStart[i] : 0, e1 = rowStart ? rowStart[i + 1] : 0;   // no rows: ContactLoop.applyForces=False, only Leapfrog's reset
  const double4 pr = S.posr[i];
  const d3 pos = xyz(pr);
  d3 F, T;
  if (u.reset) {
    F = (u.hasGravity && !(fl & (F_GRAVITY_SKIP | F_CLUMPED))) ? (S.velm[i].w * u.gravity) : mk3(0, 0, 0);      // clumped: gravity acts on the master node (Leapfrog.cpp:240-252)
    T = mk3(0, 0, 0);
    if (S.impOfId) { const int set = S.impOfId[S.id[i]]; if (set >= 0) impose_force(I, set, pos, F, T); }      // Impose::force at the reset (Leapfrog.cpp:244)
  } else { F = mk3(S.fx[i], S.fy[i], S.fz[i]); T = mk3(S.tx[i], S.ty[i], S.tz[i]); }
  for (int e = e0; e < e1; e += RB) {
    uint32_t ent[RB]; double4 fg[RB], cp[RB];
#pragma unroll
    for (int k = 0; k < RB; k++) ent[k] = (e + k < e1) ? rowEntry[e + k] : 0xffffffffu;
    // real? one L2-resident bit per contact decides, so the force and the contact point of a real contact are gathered together (one round trip)
    // and potential contacts (half of the candidates of a dense bed) cost no gather at all
    bool real[RB];
#pragma unroll
    for (int k = 0; k < RB; k++) { const uint32_t c = ent[k] >> 1; real[k] = (ent[k] != 0xffffffffu) && ((C.realBits[c >> 5] >> (c & 31)) & 1u); }
#pragma unroll
    for (int k = 0; k < RB; k++) {
      fg[k] = real[k] ? C.rD[ent[k] >> 1] : make_double4(0, 0, 0, 0);
      cp[k] = real[k] ? C.rC[ent[k] >> 1] : make_double4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < RB; k++) {
      if (!real[k]) continue;
      const uint32_t c = ent[k] >> 1;
      d3 Fc = xyz(fg[k]), xc;
      if (ent[k] & 1u) {  // this particle is pB: sign=-1, branch from pos+shift2
        Fc = Fc * (-1.);
        d3 p = pos;
        if (u.periodic) p = pos + mulMV(u.hSize, mk3((double)C.cdx[c], (double)C.cdy[c], (double)C.cdz[c]));
        xc = xyz(cp[k]) - p;
      } else xc = xyz(cp[k]) - pos;
      d3 Tc = cross(xc, Fc);
      if (u.rolling) { double4 tg = C.Tg[c]; d3 tt = xyz(tg); if (ent[k] & 1u) tt = tt * (-1.); Tc = Tc + tt; }
      F = F + Fc; T = T + Tc;
    }
  }
  S.fx[i] = F.x; S.fy[i] = F.y; S.fz[i] = F.z;
  S.tx[i] = T.x; S.ty[i] = T.y; S.tz[i] = T.z;
}

// ------------------------------------------------------------------------------------------------
// DynDt::nodalCritDtSq (src/dem/DynDt.cpp:16-58): per-sphere stiffness sums over real contacts -> min of 2*m/k.
// ------------------------------------------------------------------------------------------------
// DynDt::nodalStiffAdd (DynDt.cpp:16-36) for the sphere in slot i: its real contacts in the order of its incidence row
__device__ __forceinline__ void nodal_stiff_add(const Spheres& S, const Contacts& C, const int* __restrict__ rowStart, const uint32_t* __restrict__ rowEntry, int law, int i, d3& ktr, d3& krot) {
      const double myR = S.posr[i].w;
      for (int e = rowStart[i]; e < rowStart[i + 1]; e++) {
        const uint32_t ent = rowEntry[e]; const uint32_t c = ent >> 1;
        const uint32_t cfl = C.flags[c];
        if (!(cfl & CF_REAL)) continue;
        d3 c0 = xyz(C.rA[c]), c1 = xyz(C.rB[c]);
        d3 c2 = cross(c0, c1);
        d3 nrmv = quatRotate(quatFromCols(c0, c1, c2), mk3(1, 0, 0));
        d3 n2 = mk3(nrmv.x * nrmv.x, nrmv.y * nrmv.y, nrmv.z * nrmv.z);
        double kn, kt;
        const double4 php = C.ph[c];
        if (law == 0) { kn = php.x; kt = php.y; }
        else { double uN = C.rC[c].w; kn = (3 / 2.) * php.x * sqrt(-uN); kt = php.y * sqrt(uN < 0 ? -uN : 0); }
        // lens[ix] (iniLensTouch): SS: own radius; WS/FS: the sphere's radius on both sides (L6Geom.cpp:61, Wall.cpp:82)
        double len = myR;
        if (cf_kind(cfl) == KIND_SS) { int2 pr = C.pair[c]; len = (ent & 1u) ? S.posr[pr.y].w : S.posr[pr.x].w; }
        ktr = ktr + (n2 * (kn - kt) + mk3(kt, kt, kt));
        krot = krot + (len * len) * kt * mk3(n2.y + n2.z, n2.z + n2.x, n2.x + n2.y);
      }
}
__device__ __forceinline__ void block_min_store(double ret, double* blockMin) {
  __shared__ double sm[256];
  sm[threadIdx.x] = ret; __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if ((int)threadIdx.x < s) sm[threadIdx.x] = fmin(sm[threadIdx.x], sm[threadIdx.x + s]); __syncthreads(); }
  if (threadIdx.x == 0) blockMin[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) k_dyndt(Spheres S, int n, Contacts C, const int* __restrict__ rowStart, const uint32_t* __restrict__ rowEntry, int law,
                                              double* blockMin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ret = INFINITY;
  if (i < n) {
    const uint32_t fl = S.flags[i];
    if ((fl & F_DOF_ALL) != F_DOF_ALL && !(fl & (F_GHOST | F_CLUMPED))) {      // a clumped node is not in DemField.nodes: its contacts count for the master node (k_clumps_dyndt)
      d3 ktr = mk3(0, 0, 0), krot = mk3(0, 0, 0);
      nodal_stiff_add(S, C, rowStart, rowEntry, law, i, ktr, krot);
      const double mass = S.velm[i].w, inertia = S.angi[i].w;
      if (ktr.x != 0 && mass > 0. && !(fl & 1u)) ret = fmin(ret, mass / fabs(ktr.x));
      if (ktr.y != 0 && mass > 0. && !(fl & 2u)) ret = fmin(ret, mass / fabs(ktr.y));
      if (ktr.z != 0 && mass > 0. && !(fl & 4u)) ret = fmin(ret, mass / fabs(ktr.z));
      if (krot.x != 0 && inertia > 0. && !(fl & 8u)) ret = fmin(ret, inertia / fabs(krot.x));
      if (krot.y != 0 && inertia > 0. && !(fl & 16u)) ret = fmin(ret, inertia / fabs(krot.y));
      if (krot.z != 0 && inertia > 0. && !(fl & 32u)) ret = fmin(ret, inertia / fabs(krot.z));
      ret = 2 * ret;
    }
  }
  block_min_store(ret, blockMin);
}
// DynDt::nodalCritDtSq for clump master nodes (DynDt.cpp:48-52): contacts of all members, member by member, against the clump's mass and principal inertia
__global__ void __launch_bounds__(256) k_clumps_dyndt(Clumps K, Spheres S, const int* __restrict__ id2slot, Contacts C, const int* __restrict__ rowStart, const uint32_t* __restrict__ rowEntry,
                                                     int law, double* blockMin) {
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  double ret = INFINITY;
  if (ci < K.n) {
    const uint32_t fl = K.flags[ci];
    if ((fl & F_DOF_ALL) != F_DOF_ALL) {
      d3 ktr = mk3(0, 0, 0), krot = mk3(0, 0, 0);
      for (int m = K.memStart[ci]; m < K.memStart[ci + 1]; m++) nodal_stiff_add(S, C, rowStart, rowEntry, law, id2slot[K.memId[m]], ktr, krot);
      const double mass = K.pos[ci].w; const d3 inertia = xyz(K.inertia[ci]);
      if (ktr.x != 0 && mass > 0. && !(fl & 1u)) ret = fmin(ret, mass / fabs(ktr.x));
      if (ktr.y != 0 && mass > 0. && !(fl & 2u)) ret = fmin(ret, mass / fabs(ktr.y));
      if (ktr.z != 0 && mass > 0. && !(fl & 4u)) ret = fmin(ret, mass / fabs(ktr.z));
      if (krot.x != 0 && inertia.x > 0. && !(fl & 8u)) ret = fmin(ret, inertia.x / fabs(krot.x));
      if (krot.y != 0 && inertia.y > 0. && !(fl & 16u)) ret = fmin(ret, inertia.y / fabs(krot.y));
      if (krot.z != 0 && inertia.z > 0. && !(fl & 32u)) ret = fmin(ret, inertia.z / fabs(krot.z));
      ret = 2 * ret;
    }
  }
  block_min_store(ret, blockMin);
}
__global__ void __launch_bounds__(256) k_id2slot_fill(const int32_t* __restrict__ id, int n, int* id2slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) id2slot[id[i]] = i;
}

// kinetic energy Σ ½ m v·v + ½ I ω·ω (aperiodic: DemData::getEk_any, Particle.cpp:247-265); per-block partial sums, fixed order
__global__ void __launch_bounds__(256) k_ekin(Spheres S, int n, StepU u, double* blockSum) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0;
  if (i < n && !(S.flags[i] & F_GHOST)) {
    double4 v = S.velm[i], w = S.angi[i];
    d3 fv = xyz(v), fw = xyz(w);
    if (u.periodic) {   // Cell::pprevFluctVel / pprevFluctAngVel (Cell.cpp:36-57)
      const d3 pos = xyz(S.posr[i]);
      if (u.homoDeform == 2 || u.homoDeform == 3) fv = fv - mulMV(u.gradV, pos);
      else if (u.homoDeform == 4) { fv = fv - mulMV(u.gradV, pos - (u.dt / 2) * fv); fw = fw - u.spinVec; }
    }
    e = .5 * (v.w * sqn(fv)); if (!isinf(w.w)) e += .5 * w.w * sqn(fw);
  }
  __shared__ double sm[256];
  sm[threadIdx.x] = e; __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) blockSum[blockIdx.x] = sm[0];
}

// ContactLoop.stress (ContactLoop.cpp:116,218-231): sum over real contacts of F (x) branch, F = global force on pA (C.Fg), branch = Contact::dPos
// (Contact.cpp:54-59).  Evaluated on demand from the state the last step left; per-block partial sums in a fixed tree order, summed by the host in
// block order (the reference's own order depends on its thread schedule).  Facets are multi-nodal: the reference throws (ContactLoop.cpp:220).
__global__ void __launch_bounds__(256) k_stress(Contacts C, long long n, Spheres S, const Wall* __restrict__ walls, StepU u, double* blockSum /*9 per block*/, int* hasFacet) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (c < n) {
    const uint32_t fl = C.flags[c];
    if ((fl & CF_REAL) && !(fl & CF_DEAD)) {
      const int kind = cf_kind(fl); const int2 pr = C.pair[c];
      if (kind == KIND_FS) *hasFacet = 1;
      else if (!(S.flags[pr.y] & F_GHOST) || (kind == KIND_SS && !(S.flags[pr.x] & F_GHOST))) {
        // decomposed runs: a cross-slab pair lives on both sides; the side owning pA (sphere-sphere) / the sphere (wall) reports it
        const bool mineToReport = kind == KIND_SS ? !(S.flags[pr.x] & F_GHOST) : !(S.flags[pr.y] & F_GHOST);
        if (mineToReport) {
          const double4 F = C.rD[c];
          const d3 pb = xyz(S.posr[pr.y]);
          const d3 pa = kind == KIND_SS ? xyz(S.posr[pr.x]) : mk3(walls[pr.x].pos[0], walls[pr.x].pos[1], walls[pr.x].pos[2]);
          d3 br = pb - pa;
          if (u.periodic) { const int cx = C.cdx[c], cy = C.cdy[c], cz = C.cdz[c]; if (cx != 0 || cy != 0 || cz != 0) br = br + mulMV(u.hSize, mk3((double)cx, (double)cy, (double)cz)); }
          v[0] = F.x * br.x; v[1] = F.x * br.y; v[2] = F.x * br.z; v[3] = F.y * br.x; v[4] = F.y * br.y; v[5] = F.y * br.z; v[6] = F.z * br.x; v[7] = F.z * br.y; v[8] = F.z * br.z;
        }
      }
    }
  }
  __shared__ double sm[9][256];
  for (int k = 0; k < 9; k++) sm[k][threadIdx.x] = v[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if ((int)threadIdx.x < s) for (int k = 0; k < 9; k++) sm[k][threadIdx.x] += sm[k][threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) for (int k = 0; k < 9; k++) blockSum[9ll * blockIdx.x + k] = sm[k][0];
}

// ================================================================================================
// Broad phase (replaces InsertionSortCollider::run, src/dem/InsertionSortCollider.cpp:386-613): at every full run the set of
// potential contacts is rebuilt as { (i,j) : enlarged AABBs overlap (closed intervals, spatialOverlap :67-73) and
// Collider::mayCollide (Collision.cpp:21-43) } from a uniform grid over radix-sorted cell keys.
// ================================================================================================
// Size-class grids: level L holds the spheres with radius in [rmin*2^L, rmin*2^(L+1)); its cells are as large as the largest enlarged box of the
// level, so a sphere finds every partner of its own or a coarser level within 3x3x3 cells of that level's grid.  The SMALLER sphere (lower level;
// same level: lower slot) emits the pair, which keeps the work per thread bounded (<= 27 cells per level) for wide size distributions where a single
// grid sized for the smallest sphere makes the largest ones scan hundreds of cells.  Cells of all levels share one key space / cellStart array.
#define WD_MAX_LEVELS 8
struct HGrid {
  int nL, periodic;
  double rmin;
  double ox, oy, oz;                    // aperiodic: origin (lower corner of the bounding box of the centres)
  double inv[WD_MAX_LEVELS][3];         // cells per unit length
  int n[WD_MAX_LEVELS][3];              // cells per axis
  int base[WD_MAX_LEVELS + 1];          // first key of the level
  double reach[WD_MAX_LEVELS][3];       // largest half-extent of an enlarged box of the level, per axis (periodic: shear-aligned)
};
__host__ __device__ __forceinline__ int levelOf(const HGrid& g, double r) {
  int L = 0;
  while (L + 1 < g.nL && r >= g.rmin * (double)(2 << L)) L++;
  return L;
}

__global__ void __launch_bounds__(256) k_bounds(const double4* __restrict__ posr, int n, double* blk /*7 per block*/, double* out /*7: min xyz, max xyz, max radius*/, unsigned* ticket) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[7] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY, 0};
  if (i < n) { double4 p = posr[i]; v[0] = v[3] = p.x; v[1] = v[4] = p.y; v[2] = v[5] = p.z; v[6] = p.w; }
  __shared__ double sm[7][256];
  for (int k = 0; k < 7; k++) sm[k][threadIdx.x] = v[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      for (int k = 0; k < 3; k++) sm[k][threadIdx.x] = fmin(sm[k][threadIdx.x], sm[k][threadIdx.x + s]);
      for (int k = 3; k < 7; k++) sm[k][threadIdx.x] = fmax(sm[k][threadIdx.x], sm[k][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) for (int k = 0; k < 7; k++) blk[7 * blockIdx.x + k] = sm[k][0];
  // the last block to finish folds the block results into the 7 numbers
  __shared__ unsigned last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u);
  __syncthreads();
  if (last != gridDim.x - 1) return;
  __threadfence();
  double w[7] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY, 0};
  for (int b = threadIdx.x; b < (int)gridDim.x; b += 256) {
    for (int k = 0; k < 3; k++) w[k] = fmin(w[k], blk[7 * b + k]);
    for (int k = 3; k < 7; k++) w[k] = fmax(w[k], blk[7 * b + k]);
  }
  __syncthreads();
  for (int k = 0; k < 7; k++) sm[k][threadIdx.x] = w[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      for (int k = 0; k < 3; k++) sm[k][threadIdx.x] = fmin(sm[k][threadIdx.x], sm[k][threadIdx.x + s]);
      for (int k = 3; k < 7; k++) sm[k][threadIdx.x] = fmax(sm[k][threadIdx.x], sm[k][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { for (int k = 0; k < 7; k++) out[k] = sm[k][0]; *ticket = 0; }
}

// (kept for reference / tools) second stage as a kernel of its own: the block results -> 7 numbers (min xyz, max xyz, max radius)
__global__ void __launch_bounds__(256) k_bounds_final(const double* __restrict__ blk, int nb, double* out) {
  double v[7] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY, 0};
  for (int b = threadIdx.x; b < nb; b += 256) {
    for (int k = 0; k < 3; k++) v[k] = fmin(v[k], blk[7 * b + k]);
    for (int k = 3; k < 7; k++) v[k] = fmax(v[k], blk[7 * b + k]);
  }
  __shared__ double sm[7][256];
  for (int k = 0; k < 7; k++) sm[k][threadIdx.x] = v[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      for (int k = 0; k < 3; k++) sm[k][threadIdx.x] = fmin(sm[k][threadIdx.x], sm[k][threadIdx.x + s]);
      for (int k = 3; k < 7; k++) sm[k][threadIdx.x] = fmax(sm[k][threadIdx.x], sm[k][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) for (int k = 0; k < 7; k++) out[k] = sm[k][0];
}

__device__ __forceinline__ int cellCoord(double x, double o, double inv, int n) {
  int c = (int)floor((x - o) * inv);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void __launch_bounds__(256) k_cellkeys(const double4* __restrict__ posr, int n, HGrid g, uint32_t* keys, int* idx, const uint32_t* __restrict__ deadFlags, uint32_t deadKey) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (deadFlags && (deadFlags[i] & F_DEADSLOT)) { keys[i] = deadKey; idx[i] = i; return; }
  double4 p = posr[i];
  const int L = levelOf(g, p.w);
  int ix = cellCoord(p.x, g.ox, g.inv[L][0], g.n[L][0]), iy = cellCoord(p.y, g.oy, g.inv[L][1], g.n[L][1]), iz = cellCoord(p.z, g.oz, g.inv[L][2], g.n[L][2]);
  keys[i] = (uint32_t)(g.base[L] + (iz * g.n[L][1] + iy) * g.n[L][0] + ix);
  idx[i] = i;
}

// gather all per-sphere arrays into cell order; also refreshes lastPos (Aabb::nodeLastPos, Collision.cpp:138-141)
__global__ void __launch_bounds__(256) k_reorder(Spheres dst, Spheres src, const int* __restrict__ perm, int n, int* id2slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int o = perm[i];
  double4 p = src.posr[o];
  dst.posr[i] = p; dst.velm[i] = src.velm[o]; dst.angi[i] = src.angi[o]; dst.ori[i] = src.ori[o];
  dst.fx[i] = src.fx[o]; dst.fy[i] = src.fy[o]; dst.fz[i] = src.fz[o];
  dst.tx[i] = src.tx[o]; dst.ty[i] = src.ty[o]; dst.tz[i] = src.tz[o];
  dst.lx[i] = p.x; dst.ly[i] = p.y; dst.lz[i] = p.z;
  dst.flags[i] = src.flags[o]; dst.mask[i] = src.mask[o]; dst.mat[i] = src.mat[o];
  int id = src.id[o]; dst.id[i] = id; dst.up[i] = src.up[o];
  id2slot[id] = i;
}

// cellStart[k] = number of spheres with key < k : scatter run ends, then inclusive max-scan (done by the host with cub)
__global__ void __launch_bounds__(256) k_cellends(const uint32_t* __restrict__ keys, int n, int* cellStart) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == n - 1 || keys[i] != keys[i + 1]) cellStart[keys[i] + 1] = i + 1;
}

__device__ __forceinline__ bool aabbOverlapSS(double4 a, double4 b, double vd) {
  // Bo1_Sphere_Aabb::goGeneric + AabbCollider::updateAabb (Sphere.cpp:52-66, Collision.cpp:156-159): min=(pos-r)-vd, max=(pos+r)+vd
  return ((a.x - a.w) - vd <= (b.x + b.w) + vd) && ((a.x + a.w) + vd >= (b.x - b.w) - vd) &&
         ((a.y - a.w) - vd <= (b.y + b.w) + vd) && ((a.y + a.w) + vd >= (b.y - b.w) - vd) &&
         ((a.z - a.w) - vd <= (b.z + b.w) + vd) && ((a.z + a.w) + vd >= (b.z - b.w) - vd);
}
__device__ __forceinline__ bool mayCollide(uint32_t mA, uint32_t mB, uint32_t lone) { return (mA & mB) && !(mA & mB & lone); }  // Collision.cpp:21-43

struct OldCsr {  // incidence rows of the previous contact table (for history carry-over)
  const int* rowStart; const uint32_t* rowEntry; const int32_t* rowOther; int valid;
  int nRows;   // spheres the rows were built for; slots beyond (spheres appended since, wdem_add_spheres) have no previous contacts
  // history rows that arrived with spheres migrating in from another subdomain (sorted by key (idA<<32|idB)); see wdem_domain.cuh
  const unsigned long long* immKeys; const int* immVals; int nImm;
  int immAll;   // rows imported by wdem_upload_contacts: consult them for every pair, not only for spheres that just arrived
};

__device__ __forceinline__ void copy_history(const Contacts& N, long long cn, const Contacts& O, long long co, int law, int rolling, int periodic) {
  // the whole 128-byte record (frame, shear force, contact point, uN, global force) + the Cp2 constants
  const double4 r0 = O.rA[co], r1 = O.rB[co], r2 = O.rC[co], r3 = O.rD[co], p = O.ph[co];
  N.rA[cn] = r0; N.rB[cn] = r1; N.rC[cn] = r2; N.rD[cn] = r3; N.ph[cn] = p;
  if (law == 1) N.ph4[cn] = O.ph4[co];
  if (rolling) N.tq[cn] = O.tq[co];
  N.stepCreated[cn] = O.stepCreated[co];
  if (periodic) { N.cdx[cn] = O.cdx[co]; N.cdy[cn] = O.cdy[co]; N.cdz[cn] = O.cdz[co]; }
}

// find the previous contact between the sphere in (old) slot `oslot` and the particle with id `otherId`
__device__ __forceinline__ long long find_old(const OldCsr& oc, int oslot, int otherId) {
  if (!oc.valid || oslot < 0 || oslot >= oc.nRows) return -1;
  int lo = oc.rowStart[oslot], hi = oc.rowStart[oslot + 1] - 1;   // rows are sorted by the partner's id (k_row_sort)
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1; const int v = oc.rowOther[mid];
    if (v == otherId) return (long long)(oc.rowEntry[mid] >> 1);
    if (v < otherId) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}
// same, falling back to the rows that migrated in when one of the two spheres arrived in this full run
__device__ __forceinline__ long long find_old2(const OldCsr& oc, int oslot, int otherId, uint32_t arrivedBits, int idA, int idB) {
  long long co = find_old(oc, oslot, otherId);
  if (co < 0 && oc.nImm > 0 && (oc.immAll || (arrivedBits & F_ARRIVED))) {
    const unsigned long long key = ((unsigned long long)(uint32_t)idA << 32) | (uint32_t)idB;
    int lo = 0, hi = oc.nImm - 1;
    while (lo <= hi) { const int mid = (lo + hi) >> 1; const unsigned long long k = oc.immKeys[mid]; if (k == key) return oc.immVals[mid]; if (k < key) lo = mid + 1; else hi = mid - 1; }
  }
  return co;
}

// Sphere+Sphere candidates (see HGrid for the ownership rule).  FILL=false counts, FILL=true writes pairs (canonical pA = lower id) and looks up the
// previous contact of the pair (history carry-over by k_carry).
// wall candidates found in the same pass as the sphere candidates (aperiodic scenes with at most 32 walls): the counting pass leaves the walls whose boxes overlap
// the sphere's as a bit mask, the filling pass writes those pairs without touching the sphere's position again.  walls == nullptr: separate k_pairs_ws launches.
struct WallPass { const Wall* walls; int nWalls; double halfThick; uint32_t* bits; };
#ifndef WD_CAND_CAP
#define WD_CAND_CAP 16
#endif
#ifndef WD_PAIRS_MINB
#define WD_PAIRS_MINB 10   // 48 registers (round 1 measured 40/48/64 registers within 3 %; the dormant test needs a few more than 40)
#endif
// (Measured and dropped: a single-precision rejection test on a float4 copy of the positions in front of the exact one — half the bytes and a quarter of the fp64
// work per visited sphere, yet 1.08 ms per full run instead of 1.05: the search waits for dependent loads, and the survivors then pay a second one.)
// DORMANT potential contacts.  Between two collider full runs no sphere is farther than verletDist from where it was at the last one (the Verlet test runs in every
// step, before the contacts, and triggers the full run in the same step: Collision.cpp:113-121), so a candidate pair whose centres were farther apart than
// |distFactor|(r1+r2) + 2 verletDist at the full run cannot touch before the next one: Cg2_Sphere_Sphere_L6Geom::go would return false in every step in between
// (Sphere.cpp:101-104) and leave no trace.  In a dense bed that is ~80 % of the potential contacts (the enlarged boxes overlap long before the spheres can touch).
// Such pairs stay in the table (the candidate set is the reference's), flagged CF_DORMANT in a segment of their own behind the active sphere-sphere contacts,
// but the contact kernel does not visit them and they take no place in the incidence rows.  dormMargin2 < 0 switches this off (periodic cells: the shift
// hSize*cellDist moves with the cell; Law2.noBreak: a real contact may live at any distance).
__device__ __forceinline__ bool isDormant(double4 a, double4 b, double adf, double vd2x, double on) {
  if (!(on > 0.)) return false;
  const double rs = adf * (a.w + b.w) + vd2x;                 // |distFactor|(r1+r2) + 2 verletDist
  const d3 dl = xyz(b) - xyz(a);
  return sqn(dl) > (rs * rs) * (1. + 1e-9);
}
template <bool FILL>
__global__ void __launch_bounds__(128, WD_PAIRS_MINB) k_pairs_ss(Spheres S, int n, HGrid g, const int* __restrict__ cellStart, double vd, uint32_t loneMask, int* counts,
                                                 const int* __restrict__ offsets, Contacts N, Contacts O, OldCsr oc, const int* __restrict__ perm,
                                                 int* __restrict__ rowCount, int* __restrict__ srcOld, int* __restrict__ cand, WallPass wp, double adf, double dormOn) {
  // counts / offsets blocks: [0,n) active sphere-sphere pairs by emitting sphere, [n,2n) dormant ones, [2n,3n) wall pairs, [3n,4n) facet pairs
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t fi = S.flags[i];
  int cntLocal = 0, dormLocal = 0;
  long long out = FILL ? offsets[i] : 0, outD = FILL ? offsets[n + i] : 0;
  int idI = 0, oslotI = 0;
  if (FILL) { idI = S.id[i]; oslotI = perm ? perm[i] : i; }
  if (FILL && wp.walls && !(fi & F_GHOST)) {   // wall pairs of this sphere (segment [nSS, nSS+nWS) of the table: offsets of the third counts block)
    uint32_t m = wp.bits[i]; long long c = offsets[2 * n + i];
    while (m) {
      const int w = __ffs(m) - 1; m &= m - 1;
      const int wid = wp.walls[w].id;
      N.pair[c] = make_int2(w, i);
      srcOld[c] = (int)find_old2(oc, oslotI, wid, fi, wid, idI);
      N.flags[c] = (KIND_WS << CF_KIND_SHIFT);
      c++;
    }
  }
  // the counting pass leaves the partners it found (up to WD_CAND_CAP per sphere, in discovery order; bit 31: dormant) in `cand`; the filling pass then only has to
  // look up the previous contact and write the pair, without walking the cells again.  Spheres with more partners than the cache holds repeat the search.
  if (FILL && cand) {
    const int cnt = counts[i] + counts[n + i];
    if (cnt <= WD_CAND_CAP) {
      int ka = 0, kd = 0;
      for (int k = 0; k < cnt; k++) {
        const int je = cand[(long long)k * n + i];
        const int j = je & 0x7fffffff; const bool dorm = je < 0;
        const int idJ = S.id[j];
        const bool iIsA = idI < idJ;
        const long long c = dorm ? outD + (kd++) : out + (ka++);
        N.pair[c] = iIsA ? make_int2(i, j) : make_int2(j, i);
        srcOld[c] = dorm ? -1 : (int)find_old2(oc, oslotI, idJ, oc.nImm > 0 ? (fi | S.flags[j]) : 0u, iIsA ? idI : idJ, iIsA ? idJ : idI);
        N.flags[c] = (KIND_SS << CF_KIND_SHIFT) | (dorm ? (uint32_t)CF_DORMANT : 0u);
      }
      return;
    }
  }
  const double4 pi = S.posr[i];
  const int Li = levelOf(g, pi.w);
  const uint32_t mi = S.mask[i];
  for (int L = Li; L < g.nL; L++) {
    const double H = ((pi.w + vd) + g.reach[L][0]) * (1 + 1e-12) + 1e-300;
    const int nx = g.n[L][0], ny = g.n[L][1], nz = g.n[L][2];
    const int x0 = cellCoord(pi.x - H, g.ox, g.inv[L][0], nx), x1 = cellCoord(pi.x + H, g.ox, g.inv[L][0], nx);
    const int y0 = cellCoord(pi.y - H, g.oy, g.inv[L][1], ny), y1 = cellCoord(pi.y + H, g.oy, g.inv[L][1], ny);
    const int z0 = cellCoord(pi.z - H, g.oz, g.inv[L][2], nz), z1 = cellCoord(pi.z + H, g.oz, g.inv[L][2], nz);
    for (int iz = z0; iz <= z1; iz++)
      for (int iy = y0; iy <= y1; iy++) {
        const int k0 = g.base[L] + (iz * ny + iy) * nx;
        int jb = cellStart[k0 + x0]; const int je = cellStart[k0 + x1 + 1];
        if (L == Li && jb <= i) jb = i + 1;   // same level: the lower slot emits (slots of a level are contiguous and ascending along the row)
        for (int j = jb; j < je; j++) {
          const double4 pj = S.posr[j];
          if (!aabbOverlapSS(pi, pj, vd)) continue;
          if (!mayCollide(mi, S.mask[j], loneMask)) continue;
          if ((fi & F_GHOST) && (S.flags[j] & F_GHOST)) continue;   // ghost-ghost pairs belong to other subdomains
          if ((fi & F_CLUMPED) && S.clumpOfId && (S.flags[j] & F_CLUMPED) && S.clumpOfId[S.id[i]] == S.clumpOfId[S.id[j]]) continue;   // members of one clump (Collision.cpp:34-39)
          const bool dorm = isDormant(pi, pj, adf, 2. * vd, dormOn);
          if (FILL) {
            const int idJ = S.id[j];
            const long long c = dorm ? outD + dormLocal : out + cntLocal;
            const bool iIsA = idI < idJ;
            N.pair[c] = iIsA ? make_int2(i, j) : make_int2(j, i);
            srcOld[c] = dorm ? -1 : (int)find_old2(oc, oslotI, idJ, oc.nImm > 0 ? (fi | S.flags[j]) : 0u, iIsA ? idI : idJ, iIsA ? idJ : idI);
            N.flags[c] = (KIND_SS << CF_KIND_SHIFT) | (dorm ? (uint32_t)CF_DORMANT : 0u);
          } else {
            if (!dorm) atomicAdd(&rowCount[j], 1);
            if (cand && cntLocal + dormLocal < WD_CAND_CAP) cand[(long long)(cntLocal + dormLocal) * n + i] = dorm ? (j | (int)0x80000000) : j;
          }
          if (dorm) dormLocal++; else cntLocal++;
        }
      }
  }
  if (!FILL) {
    int wcnt = 0;
    if (wp.walls) {   // Bo1_Wall_Aabb (Wall.cpp:25-38), the test of k_pairs_ws
      uint32_t m = 0;
      if (!(fi & F_GHOST)) {
        for (int w = 0; w < wp.nWalls; w++) {
          const Wall W = wp.walls[w];
          const double pc = comp(xyz(pi), W.axis);
          double wmin = W.pos[W.axis], wmax = wmin;
          if (wp.halfThick > 0) { wmin -= wp.halfThick; wmax += wp.halfThick; }
          wmin = wmin - vd; wmax = wmax + vd;
          if (!(((pc - pi.w) - vd <= wmax) && ((pc + pi.w) + vd >= wmin))) continue;
          if (!mayCollide(mi, W.mask, loneMask)) continue;
          m |= 1u << w; wcnt++;
        }
      }
      wp.bits[i] = m; counts[2 * n + i] = wcnt;
    }
    counts[i] = cntLocal; counts[n + i] = dormLocal; if (cntLocal + wcnt) atomicAdd(&rowCount[i], cntLocal + wcnt);
  }
}

// Wall+Sphere candidates: Bo1_Wall_Aabb (Wall.cpp:25-38): the wall box is the plane coordinate (± halfThick) ± verletDist on its axis, infinite elsewhere.
template <bool FILL>
__global__ void __launch_bounds__(128) k_pairs_ws(Spheres S, int n, const Wall* __restrict__ walls, int nWalls, double vd, double wallHalfThick, uint32_t loneMask,
                                                 int* counts, const int* __restrict__ offsets, Contacts N, Contacts O, OldCsr oc, const int* __restrict__ perm,
                                                 int* __restrict__ rowCount, int* __restrict__ srcOld) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 pi = S.posr[i];
  const uint32_t mi = S.mask[i];
  const uint32_t fi = S.flags[i];
  if (fi & F_GHOST) { if (!FILL) counts[i] = 0; return; }   // boundary contacts of a ghost are its owner's business
  int cntLocal = 0;
  long long out = FILL ? offsets[i] : 0;
  const int oslotI = FILL ? (perm ? perm[i] : i) : 0;
  for (int w = 0; w < nWalls; w++) {
    const Wall W = walls[w];
    const double pc = comp(xyz(pi), W.axis);
    double wmin = W.pos[W.axis], wmax = wmin;
    if (wallHalfThick > 0) { wmin -= wallHalfThick; wmax += wallHalfThick; }
    wmin = wmin - vd; wmax = wmax + vd;
    if (!(((pc - pi.w) - vd <= wmax) && ((pc + pi.w) + vd >= wmin))) continue;
    if (!mayCollide(mi, W.mask, loneMask)) continue;
    if (FILL) {
      const long long c = out + cntLocal;
      N.pair[c] = make_int2(w, i);
      srcOld[c] = (int)find_old2(oc, oslotI, W.id, fi, W.id, S.id[i]);
      N.flags[c] = (KIND_WS << CF_KIND_SHIFT);
    }
    cntLocal++;
  }
  if (!FILL) { counts[i] = cntLocal; if (cntLocal) atomicAdd(&rowCount[i], cntLocal); }
}

// device-side binning of the facets into the (fixed) facet grid at a full run: count per cell, scan (host: cub), fill through per-cell cursors
__global__ void __launch_bounds__(128) k_fgrid_bin(const Facet* __restrict__ facets, int n, FacetGrid fg, double vd, int* __restrict__ cellCount /*+1 shifted*/, int* __restrict__ cursor,
                                                  int* __restrict__ cellFacet, int fill) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const Facet* F = facets + f;
  const int x0 = cellCoord(F->bmin[0] - vd, fg.ox, fg.inv, fg.nx), x1 = cellCoord(F->bmax[0] + vd, fg.ox, fg.inv, fg.nx);
  const int y0 = cellCoord(F->bmin[1] - vd, fg.oy, fg.inv, fg.ny), y1 = cellCoord(F->bmax[1] + vd, fg.oy, fg.inv, fg.ny);
  const int z0 = cellCoord(F->bmin[2] - vd, fg.oz, fg.inv, fg.nz), z1 = cellCoord(F->bmax[2] + vd, fg.oz, fg.inv, fg.nz);
  for (int iz = z0; iz <= z1; iz++) for (int iy = y0; iy <= y1; iy++) for (int ix = x0; ix <= x1; ix++) {
    const int k = (iz * fg.ny + iy) * fg.nx + ix;
    if (!fill) atomicAdd(&cellCount[k], 1);
    else cellFacet[fg.cellStart[k] + atomicAdd(&cursor[k], 1)] = f;
  }
}

// Facet+Sphere candidates: Bo1_Facet_Aabb (Facet.cpp:275-298): vertex min/max ± halfThick, then ± verletDist.  Facets are static and are
// binned once (host) into a uniform grid; each sphere visits the bins its enlarged AABB touches and emits a facet from the single bin
// that holds the lower corner of the intersection of the two boxes (no duplicates).
template <bool FILL>
__global__ void __launch_bounds__(128) k_pairs_fs(Spheres S, int n, const Facet* __restrict__ facets, FacetGrid fg, double vd, uint32_t loneMask, int* counts,
                                                 const int* __restrict__ offsets, Contacts N, Contacts O, OldCsr oc, const int* __restrict__ perm,
                                                 int* __restrict__ rowCount, int* __restrict__ srcOld) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 pi = S.posr[i];
  const uint32_t mi = S.mask[i];
  const uint32_t fi = S.flags[i];
  if (fi & F_GHOST) { if (!FILL) counts[i] = 0; return; }
  const double smin[3] = {(pi.x - pi.w) - vd, (pi.y - pi.w) - vd, (pi.z - pi.w) - vd}, smax[3] = {(pi.x + pi.w) + vd, (pi.y + pi.w) + vd, (pi.z + pi.w) + vd};
  const int x0 = cellCoord(smin[0], fg.ox, fg.inv, fg.nx), x1 = cellCoord(smax[0], fg.ox, fg.inv, fg.nx);
  const int y0 = cellCoord(smin[1], fg.oy, fg.inv, fg.ny), y1 = cellCoord(smax[1], fg.oy, fg.inv, fg.ny);
  const int z0 = cellCoord(smin[2], fg.oz, fg.inv, fg.nz), z1 = cellCoord(smax[2], fg.oz, fg.inv, fg.nz);
  int cntLocal = 0;
  long long out = FILL ? offsets[i] : 0;
  const int oslotI = FILL ? (perm ? perm[i] : i) : 0;
  for (int iz = z0; iz <= z1; iz++)
    for (int iy = y0; iy <= y1; iy++)
      for (int ix = x0; ix <= x1; ix++) {
        const int k = (iz * fg.ny + iy) * fg.nx + ix;
        for (int e = fg.cellStart[k]; e < fg.cellStart[k + 1]; e++) {
          const int f = fg.cellFacet[e];
          const Facet* F = facets + f;
          const double fbmin[3] = {F->bmin[0] - vd, F->bmin[1] - vd, F->bmin[2] - vd}, fbmax[3] = {F->bmax[0] + vd, F->bmax[1] + vd, F->bmax[2] + vd};
          if (!(smin[0] <= fbmax[0] && smax[0] >= fbmin[0] && smin[1] <= fbmax[1] && smax[1] >= fbmin[1] && smin[2] <= fbmax[2] && smax[2] >= fbmin[2])) continue;
          // unique bin of the pair
          if (cellCoord(fmax(smin[0], fbmin[0]), fg.ox, fg.inv, fg.nx) != ix || cellCoord(fmax(smin[1], fbmin[1]), fg.oy, fg.inv, fg.ny) != iy ||
              cellCoord(fmax(smin[2], fbmin[2]), fg.oz, fg.inv, fg.nz) != iz) continue;
          if (!mayCollide(mi, F->mask, loneMask)) continue;
          if (FILL) {
            const long long c = out + cntLocal;
            N.pair[c] = make_int2(f, i);
            srcOld[c] = (int)find_old2(oc, oslotI, F->id, fi, F->id, S.id[i]);
            N.flags[c] = (KIND_FS << CF_KIND_SHIFT);
          }
          cntLocal++;
        }
      }
  if (!FILL) { counts[i] = cntLocal; if (cntLocal) atomicAdd(&rowCount[i], cntLocal); }
}

// ================================================================================================
// Periodic broad phase (replaces insertionSortPeri/handleBoundInversionPeri/spatialOverlapPeri, InsertionSortCollider.cpp:637-1060):
// AABBs live in the cell's unsheared coordinates (Bo1_Sphere_Aabb::goGeneric, Sphere.cpp:67-81), overlap is per-axis periodic and yields
// the integer cellDist.  Positions are never wrapped (InsertionSortCollider.hpp:28-29); only the grid hash wraps.
// ================================================================================================
struct PeriU { m33 unshear; d3 size; d3 cosv; int hasShear; };   // Cell::unshearPt, getSize, getCos, hasShear (Cell.hpp:82-97)

__device__ __forceinline__ void periAabb(const PeriU& pu, double4 p, double vd, d3& mn, d3& mx) {
  d3 half = mk3(p.w, p.w, p.w);
  if (pu.hasShear) {  // Cell::shearAlignedExtents (Cell.cpp:68-79): sequential += in the order ax=0,1,2
    const d3 perp = half;
    half.y += .5 * perp.y * (1 / pu.cosv.x - 1); half.z += .5 * perp.z * (1 / pu.cosv.x - 1);
    half.z += .5 * perp.z * (1 / pu.cosv.y - 1); half.x += .5 * perp.x * (1 / pu.cosv.y - 1);
    half.x += .5 * perp.x * (1 / pu.cosv.z - 1); half.y += .5 * perp.y * (1 / pu.cosv.z - 1);
  }
  const d3 u = mulMV(pu.unshear, xyz(p));
  mn = u - half; mx = u + half;
  if (vd > 0) { mx = mk3(mx.x + vd, mx.y + vd, mx.z + vd); mn = mk3(mn.x - vd, mn.y - vd, mn.z - vd); }  // AabbCollider::updateAabb (Collision.cpp:156-159)
}
__device__ __forceinline__ double cellWrapRel(double x, double x0, double x1, int& period) {  // InsertionSortCollider.cpp:624-628
  double xNorm = (x - x0) / (x1 - x0);
  double fl = floor(xNorm);
  period = (int)fl;
  return (xNorm - fl) * (x1 - x0);
}
// spatialOverlapPeri_axis, "new algorithm" (InsertionSortCollider.cpp:897-1026). Returns 0/1, or -1 for the ambiguous double contact.
__device__ __forceinline__ int overlapPeriAxis(double min1, double max1, double min2, double max2, double dim, int& period) {
  if (isinf(min1) || isinf(min2)) { period = 0; return 1; }
  int pmn1_2, pmx1_2, pmn2_1, pmx2_1;
  double mn1_2 = cellWrapRel(min1, min2, min2 + dim, pmn1_2); (void)cellWrapRel(max1, min2, min2 + dim, pmx1_2);
  double mn2_1 = cellWrapRel(min2, min1, min1 + dim, pmn2_1); (void)cellWrapRel(max2, min1, min1 + dim, pmx2_1);
  if (pmn1_2 != pmx1_2 && pmn2_1 != pmx2_1) return -1;
  if (pmn2_1 == pmx2_1) { period = -pmn2_1; return mn2_1 < (max1 - min1) ? 1 : 0; }
  period = pmn1_2; return mn1_2 < (max2 - min2) ? 1 : 0;
}
// spatialOverlapPeri (InsertionSortCollider.cpp:1042-1060), box 1 belonging to the particle with the lower id
__device__ __forceinline__ int overlapPeri(d3 mn1, d3 mx1, d3 mn2, d3 mx2, d3 size, int per[3]) {
  int r = overlapPeriAxis(mn1.x, mx1.x, mn2.x, mx2.x, size.x, per[0]); if (r <= 0) return r;
  r = overlapPeriAxis(mn1.y, mx1.y, mn2.y, mx2.y, size.y, per[1]); if (r <= 0) return r;
  return overlapPeriAxis(mn1.z, mx1.z, mn2.z, mx2.z, size.z, per[2]);
}
__device__ __forceinline__ int imod(int a, int n) { int m = a % n; return m < 0 ? m + n : m; }
__device__ __forceinline__ double wrap01(double u, double size) { double xn = u / size; xn -= floor(xn); return xn * size; }

__global__ void __launch_bounds__(256) k_cellkeys_peri(const double4* __restrict__ posr, int n, HGrid g, PeriU pu, uint32_t* keys, int* idx, const uint32_t* __restrict__ deadFlags, uint32_t deadKey) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (deadFlags && (deadFlags[i] & F_DEADSLOT)) { keys[i] = deadKey; idx[i] = i; return; }
  const double4 p = posr[i];
  const int L = levelOf(g, p.w);
  const d3 u = mulMV(pu.unshear, xyz(p));
  int ix = min(g.n[L][0] - 1, max(0, (int)floor(wrap01(u.x, pu.size.x) * g.inv[L][0])));
  int iy = min(g.n[L][1] - 1, max(0, (int)floor(wrap01(u.y, pu.size.y) * g.inv[L][1])));
  int iz = min(g.n[L][2] - 1, max(0, (int)floor(wrap01(u.z, pu.size.z) * g.inv[L][2])));
  keys[i] = (uint32_t)(g.base[L] + (iz * g.n[L][1] + iy) * g.n[L][0] + ix);
  idx[i] = i;
}

// Sphere+Sphere candidates in a periodic cell; same ownership rule as k_pairs_ss, cells wrap (positions do not).
template <bool FILL>
__global__ void __launch_bounds__(128) k_pairs_ss_peri(Spheres S, int n, HGrid g, PeriU pu, const int* __restrict__ cellStart, double vd, uint32_t loneMask, int* counts,
                                                      const int* __restrict__ offsets, Contacts N, OldCsr oc, const int* __restrict__ perm,
                                                      int* __restrict__ rowCount, int* __restrict__ srcOld, Counters* cnt, int* __restrict__ cand) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // partner cache as in k_pairs_ss; an entry holds the partner's slot and the three periods (4 ints)
  if (FILL && cand) {
    const int cntI = counts[i];
    if (cntI <= WD_CAND_CAP) {
      const int idI_ = S.id[i]; const uint32_t fi_ = S.flags[i]; const int oslot_ = perm ? perm[i] : i; const long long out_ = offsets[i];
      for (int k = 0; k < cntI; k++) {
        const int j = cand[(4ll * k + 0) * n + i];
        const int idJ = S.id[j];
        const long long c = out_ + k;
        const bool iIsA = idI_ < idJ;
        N.pair[c] = iIsA ? make_int2(i, j) : make_int2(j, i);
        N.cdx[c] = cand[(4ll * k + 1) * n + i]; N.cdy[c] = cand[(4ll * k + 2) * n + i]; N.cdz[c] = cand[(4ll * k + 3) * n + i];
        srcOld[c] = (int)find_old2(oc, oslot_, idJ, oc.nImm > 0 ? (fi_ | S.flags[j]) : 0u, iIsA ? idI_ : idJ, iIsA ? idJ : idI_);
        N.flags[c] = (KIND_SS << CF_KIND_SHIFT);
      }
      return;
    }
  }
  const double4 pi = S.posr[i];
  const int Li = levelOf(g, pi.w);
  d3 mnI, mxI; periAabb(pu, pi, vd, mnI, mxI);
  const d3 ui = mulMV(pu.unshear, xyz(pi));
  const double cx = wrap01(ui.x, pu.size.x), cy = wrap01(ui.y, pu.size.y), cz = wrap01(ui.z, pu.size.z);
  const uint32_t mi = S.mask[i];
  const uint32_t fi = S.flags[i];
  const int idI = S.id[i];
  const int oslotI = FILL ? (perm ? perm[i] : i) : 0;
  int cntLocal = 0;
  long long out = FILL ? offsets[i] : 0;
  for (int L = Li; L < g.nL; L++) {
    const int nx = g.n[L][0], ny = g.n[L][1], nz = g.n[L][2];
    // |Δcentre| <= half_i + half_j <= half_i + reach of the level; small relative + absolute slack
    const double Hx = (.5 * (mxI.x - mnI.x) + g.reach[L][0]) * (1 + 1e-9) + 1e-9 / g.inv[L][0], Hy = (.5 * (mxI.y - mnI.y) + g.reach[L][1]) * (1 + 1e-9) + 1e-9 / g.inv[L][1],
                 Hz = (.5 * (mxI.z - mnI.z) + g.reach[L][2]) * (1 + 1e-9) + 1e-9 / g.inv[L][2];
    int x0 = (int)floor((cx - Hx) * g.inv[L][0]), x1 = (int)floor((cx + Hx) * g.inv[L][0]); if (x1 - x0 + 1 >= nx) { x0 = 0; x1 = nx - 1; }
    int y0 = (int)floor((cy - Hy) * g.inv[L][1]), y1 = (int)floor((cy + Hy) * g.inv[L][1]); if (y1 - y0 + 1 >= ny) { y0 = 0; y1 = ny - 1; }
    int z0 = (int)floor((cz - Hz) * g.inv[L][2]), z1 = (int)floor((cz + Hz) * g.inv[L][2]); if (z1 - z0 + 1 >= nz) { z0 = 0; z1 = nz - 1; }
    // the x range wraps into at most two runs of consecutive cells
    const int wx0 = imod(x0, nx), wx1 = imod(x1, nx);
    const int nseg = (wx0 <= wx1) ? 1 : 2;
    for (int iz = z0; iz <= z1; iz++)
      for (int iy = y0; iy <= y1; iy++) {
        const int k0 = g.base[L] + (imod(iz, nz) * ny + imod(iy, ny)) * nx;
        for (int seg = 0; seg < nseg; seg++) {
          const int sa = (nseg == 1) ? wx0 : (seg == 0 ? wx0 : 0), sb = (nseg == 1) ? wx1 : (seg == 0 ? nx - 1 : wx1);
          const int jb = cellStart[k0 + sa], je = cellStart[k0 + sb + 1];
          for (int j = jb; j < je; j++) {
            if (L == Li && j <= i) continue;   // same level: the lower slot emits
            const double4 pj = S.posr[j];
            if (!mayCollide(mi, S.mask[j], loneMask)) continue;
            if ((fi & F_GHOST) && (S.flags[j] & F_GHOST)) continue;
            if ((fi & F_CLUMPED) && S.clumpOfId && (S.flags[j] & F_CLUMPED) && S.clumpOfId[S.id[i]] == S.clumpOfId[S.id[j]]) continue;   // members of one clump (Collision.cpp:34-39)
            d3 mnJ, mxJ; periAabb(pu, pj, vd, mnJ, mxJ);
            const int idJ = S.id[j];
            const bool iIsA = idI < idJ;
            int per[3];
            const int ov = iIsA ? overlapPeri(mnI, mxI, mnJ, mxJ, pu.size, per) : overlapPeri(mnJ, mxJ, mnI, mxI, pu.size, per);
            if (ov < 0) { cnt->error = ERR_AMBIGUOUS_PERI; cnt->errA = iIsA ? idI : idJ; cnt->errB = iIsA ? idJ : idI; continue; }
            if (!ov) continue;
            if (FILL) {
              const long long c = out + cntLocal;
              N.pair[c] = iIsA ? make_int2(i, j) : make_int2(j, i);
              N.cdx[c] = per[0]; N.cdy[c] = per[1]; N.cdz[c] = per[2];
              srcOld[c] = (int)find_old2(oc, oslotI, idJ, oc.nImm > 0 ? (fi | S.flags[j]) : 0u, iIsA ? idI : idJ, iIsA ? idJ : idI);
              N.flags[c] = (KIND_SS << CF_KIND_SHIFT);
            } else {
              atomicAdd(&rowCount[j], 1);
              if (cand && cntLocal < WD_CAND_CAP) {
                cand[(4ll * cntLocal + 0) * n + i] = j; cand[(4ll * cntLocal + 1) * n + i] = per[0]; cand[(4ll * cntLocal + 2) * n + i] = per[1]; cand[(4ll * cntLocal + 3) * n + i] = per[2];
              }
            }
            cntLocal++;
          }
        }
      }
  }
  if (!FILL) { counts[i] = cntLocal; if (cntLocal) atomicAdd(&rowCount[i], cntLocal); }
}

// Wall+Sphere candidates in a (non-skewed) periodic cell: the wall box is infinite on two axes (period 0 there) and takes part in the
// periodic overlap test on its own axis.
template <bool FILL>
__global__ void __launch_bounds__(128) k_pairs_ws_peri(Spheres S, int n, const Wall* __restrict__ walls, int nWalls, PeriU pu, double vd, double wallHalfThick, uint32_t loneMask,
                                                      int* counts, const int* __restrict__ offsets, Contacts N, OldCsr oc, const int* __restrict__ perm,
                                                      int* __restrict__ rowCount, int* __restrict__ srcOld, Counters* cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 pi = S.posr[i];
  d3 mnI, mxI; periAabb(pu, pi, vd, mnI, mxI);
  const uint32_t mi = S.mask[i];
  const uint32_t fi = S.flags[i];
  if (fi & F_GHOST) { if (!FILL) counts[i] = 0; return; }
  const int idI = S.id[i];
  int cntLocal = 0;
  long long out = FILL ? offsets[i] : 0;
  const int oslotI = FILL ? (perm ? perm[i] : i) : 0;
  for (int w = 0; w < nWalls; w++) {
    const Wall W = walls[w];
    if (!mayCollide(mi, W.mask, loneMask)) continue;
    double wmin = W.pos[W.axis], wmax = wmin;
    if (wallHalfThick > 0) { wmin -= wallHalfThick; wmax += wallHalfThick; }
    if (vd > 0) { wmax = wmax + vd; wmin = wmin - vd; }
    const double smin = comp(mnI, W.axis), smax = comp(mxI, W.axis), dim = comp(pu.size, W.axis);
    int period;
    const bool wallLo = W.id < idI;
    const int ov = wallLo ? overlapPeriAxis(wmin, wmax, smin, smax, dim, period) : overlapPeriAxis(smin, smax, wmin, wmax, dim, period);
    if (ov < 0) { cnt->error = ERR_AMBIGUOUS_PERI; cnt->errA = W.id; cnt->errB = idI; continue; }
    if (!ov) continue;
    if (FILL) {
      const long long c = out + cntLocal;
      N.pair[c] = make_int2(w, i);
      const int cd = wallLo ? period : -period;   // canonical pA = wall (Contact::swapOrder negates cellDist, Contact.cpp:12-16)
      N.cdx[c] = W.axis == 0 ? cd : 0; N.cdy[c] = W.axis == 1 ? cd : 0; N.cdz[c] = W.axis == 2 ? cd : 0;
      srcOld[c] = (int)find_old2(oc, oslotI, W.id, fi, W.id, idI);
      N.flags[c] = (KIND_WS << CF_KIND_SHIFT);
    }
    cntLocal++;
  }
  if (!FILL) { counts[i] = cntLocal; if (cntLocal) atomicAdd(&rowCount[i], cntLocal); }
}

// carry L6Geom/FrictPhys history of contacts that were real into the new table: one thread per NEW contact.  T holds (pair, kind, cellDist) in
// emission order; sphere-sphere contacts are placed real-first (stable): dst = realPos[c] for carried-real ones, nReal + (c - realPos[c]) otherwise.
// Facets in a periodic cell.  Bo1_Facet_Aabb in a periodic cell (Facet.cpp:286-297): the box of the UNSHEARED node positions +- halfThick, enlarged by verletDist
// (AabbCollider::updateAabb); nodes are never wrapped.  The boxes depend on the cell's shear, so they are refreshed at every full run.
__global__ void __launch_bounds__(128) k_facet_boxes_peri(const Facet* __restrict__ facets, int nF, PeriU pu, double vd, double* __restrict__ box /*6 per facet: min xyz, max xyz*/) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nF) return;
  const Facet* F = facets + f; const double ht = F->halfThick;
  d3 mn, mx;
  for (int i = 0; i < 3; i++) {
    const d3 p = mulMV(pu.unshear, mk3(F->v[3 * i], F->v[3 * i + 1], F->v[3 * i + 2]));
    const d3 lo = mk3(p.x - ht, p.y - ht, p.z - ht), hi = mk3(p.x + ht, p.y + ht, p.z + ht);
    if (i == 0) { mn = lo; mx = hi; }
    else { mn = mk3(fmin(mn.x, lo.x), fmin(mn.y, lo.y), fmin(mn.z, lo.z)); mx = mk3(fmax(mx.x, hi.x), fmax(mx.y, hi.y), fmax(mx.z, hi.z)); }
  }
  if (vd > 0) { mx = mk3(mx.x + vd, mx.y + vd, mx.z + vd); mn = mk3(mn.x - vd, mn.y - vd, mn.z - vd); }
  box[6 * f] = mn.x; box[6 * f + 1] = mn.y; box[6 * f + 2] = mn.z; box[6 * f + 3] = mx.x; box[6 * f + 4] = mx.y; box[6 * f + 5] = mx.z;
}
// Facet+Sphere candidates in a periodic cell: spatialOverlapPeri of the two boxes (the one of the lower id first) gives Contact.cellDist.  Every sphere visits every
// facet: meshes in periodic cells are small (a plate, an obstacle); the engine refuses more than WD_MAX_PERI_FACETS of them.
#define WD_MAX_PERI_FACETS 4096
template <bool FILL>
__global__ void __launch_bounds__(128) k_pairs_fs_peri(Spheres S, int n, const Facet* __restrict__ facets, int nF, const double* __restrict__ box, PeriU pu, double vd, uint32_t loneMask,
                                                      int* counts, const int* __restrict__ offsets, Contacts N, OldCsr oc, const int* __restrict__ perm,
                                                      int* __restrict__ rowCount, int* __restrict__ srcOld, Counters* cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t fi = S.flags[i];
  if (fi & F_GHOST) { if (!FILL) counts[i] = 0; return; }
  const double4 pi = S.posr[i];
  d3 mnI, mxI; periAabb(pu, pi, vd, mnI, mxI);
  const uint32_t mi = S.mask[i];
  const int idI = S.id[i];
  int cntLocal = 0;
  long long out = FILL ? offsets[i] : 0;
  const int oslotI = FILL ? (perm ? perm[i] : i) : 0;
  for (int f = 0; f < nF; f++) {
    const Facet* F = facets + f;
    if (!mayCollide(mi, F->mask, loneMask)) continue;
    const d3 mnF = mk3(box[6 * f], box[6 * f + 1], box[6 * f + 2]), mxF = mk3(box[6 * f + 3], box[6 * f + 4], box[6 * f + 5]);
    int per[3];
    const bool facetLo = F->id < idI;
    const int ov = facetLo ? overlapPeri(mnF, mxF, mnI, mxI, pu.size, per) : overlapPeri(mnI, mxI, mnF, mxF, pu.size, per);
    if (ov < 0) { cnt->error = ERR_AMBIGUOUS_PERI; cnt->errA = F->id; cnt->errB = idI; continue; }
    if (!ov) continue;
    if (FILL) {
      const long long c = out + cntLocal;
      N.pair[c] = make_int2(f, i);
      const int sg = facetLo ? 1 : -1;      // canonical pA = facet (Contact::swapOrder negates cellDist, Contact.cpp:12-16)
      N.cdx[c] = sg * per[0]; N.cdy[c] = sg * per[1]; N.cdz[c] = sg * per[2];
      srcOld[c] = (int)find_old2(oc, oslotI, F->id, fi, F->id, idI);
      N.flags[c] = (KIND_FS << CF_KIND_SHIFT);
    }
    cntLocal++;
  }
  if (!FILL) { counts[i] = cntLocal; if (cntLocal) atomicAdd(&rowCount[i], cntLocal); }
}

__global__ void __launch_bounds__(256) k_was_real(const int* __restrict__ srcOld, const uint32_t* __restrict__ oflags, long long nSS, int* flag) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nSS) return;
  int f = 0;
  if (c < nSS) { const int co = srcOld[c]; f = (co >= 0 && (oflags[co] & CF_REAL)) ? 1 : 0; }
  flag[c] = f;
}
__global__ void __launch_bounds__(256) k_carry(Contacts N, long long nNew, Contacts O, const int* __restrict__ srcOld, Contacts T, const int* __restrict__ realPos, long long nSS,
                                              int law, int rolling, int periodic, long long stepNow) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nNew) return;
  const int co = srcOld[c];
  const uint32_t ofl = co >= 0 ? O.flags[co] : 0u;
  const bool carried = co >= 0 && (ofl & CF_REAL);
  long long d = c;
  if (c < nSS) { const long long rp = realPos[c]; d = carried ? rp : (long long)realPos[nSS] + (c - rp); }
  N.pair[d] = T.pair[c];
  uint32_t fl = T.flags[c];
  if (periodic) { N.cdx[d] = T.cdx[c]; N.cdy[d] = T.cdy[c]; N.cdz[d] = T.cdz[c]; }
  if (carried) {
    copy_history(N, d, O, co, law, rolling, periodic);      // four 32-byte sectors + the Cp2 constants (round 1: 14 scalar planes, one sector each)
    fl |= CF_REAL;
    O.flags[co] = ofl | CF_CLAIMED;
  } else {
    // a potential contact that survives a full run keeps the cellDist it was created with (handleBoundInversionPeri only creates
    // when !hasInter, InsertionSortCollider.cpp:844-893); one whose particle was removed meanwhile (CF_DEAD, k_kill_rows) is a new contact
    if (periodic && co >= 0 && !(ofl & CF_DEAD)) { N.cdx[d] = O.cdx[co]; N.cdy[d] = O.cdy[co]; N.cdz[d] = O.cdz[co]; }
    N.stepCreated[d] = (int)stepNow;
    N.rD[d] = make_double4(0, 0, 0, 0);
  }
  N.flags[d] = fl;
}

// old real contacts nobody claimed = orphans: count (and add their incidences to the row counts), then append behind the grid candidates
__global__ void __launch_bounds__(256) k_orphans(Contacts O, long long nOld, Contacts N, long long nNewBase, const int32_t* __restrict__ oldSid, const int* __restrict__ id2slot,
                                                const uint32_t* __restrict__ sflags, int law, int rolling, int periodic, int* counter, int* __restrict__ rowCount, int doAppend) {
  long long co = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (co >= nOld) return;
  const uint32_t fl = O.flags[co];
  if (!(fl & CF_REAL) || (fl & (CF_CLAIMED | CF_DEAD))) return;
  const int kind = cf_kind(fl);
  int2 pr = O.pair[co];
  pr.y = id2slot[oldSid[pr.y]];
  if (kind == KIND_SS) pr.x = id2slot[oldSid[pr.x]];
  // decomposed runs: a partner left this subdomain, or the pair is now ghost-ghost / a ghost's boundary contact -> not ours any more
  if (pr.y < 0 || (kind == KIND_SS && pr.x < 0)) return;
  if (sflags) {
    const bool gy = sflags[pr.y] & F_GHOST;
    if (kind == KIND_SS ? (gy && (sflags[pr.x] & F_GHOST)) : gy) return;
  }
  const int k = atomicAdd(counter, 1);
  if (!doAppend) { atomicAdd(&rowCount[pr.y], 1); if (kind == KIND_SS) atomicAdd(&rowCount[pr.x], 1); return; }
  const long long c = nNewBase + k;
  N.pair[c] = pr;
  copy_history(N, c, O, co, law, rolling, periodic);
  N.flags[c] = CF_REAL | CF_ORPHAN | (fl & CF_KIND_MASK);
}

// per-particle incidence rows: rowStart = exclusive scan of rowCount (host: cub), then fill through per-row cursors and sort each row by the
// other particle's id (the canonical summation order).  Replaces a global 64-bit radix sort.
__global__ void __launch_bounds__(256) k_csr_fill(Contacts C, long long first, long long nTot, const int32_t* __restrict__ sid, const Wall* __restrict__ walls, const Facet* __restrict__ facets,
                                                 const int* __restrict__ rowStart, int* __restrict__ cursor, uint32_t* __restrict__ rowEntry, int32_t* __restrict__ rowOther) {
  long long c = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nTot) return;
  const int2 pr = C.pair[c];
  const uint32_t cfl = C.flags[c];
  if (cfl & CF_DORMANT) return;     // dormant potential contacts take no place in the incidence rows (see isDormant)
  const int kind = cf_kind(cfl);
  int idA;
  if (kind == KIND_SS) {
    idA = sid[pr.x];
    const int p = rowStart[pr.x] + atomicAdd(&cursor[pr.x], 1);
    rowEntry[p] = (uint32_t)(c << 1); rowOther[p] = sid[pr.y];
  } else idA = (kind == KIND_WS) ? walls[pr.x].id : facets[pr.x].id;
  const int p = rowStart[pr.y] + atomicAdd(&cursor[pr.y], 1);
  rowEntry[p] = (uint32_t)(c << 1) | 1u; rowOther[p] = idA;
}
// Short rows (most of them) are insertion-sorted by one thread; long rows (large spheres of a wide size distribution) are queued and
// sorted by one warp each (bitonic network in shared memory), so that a few long rows do not serialise whole warps.
#define ROW_SHORT 16
#define ROW_WARP_MAX 512
__device__ __forceinline__ void row_insertion_sort(int b, int e, uint32_t* __restrict__ rowEntry, int32_t* __restrict__ rowOther) {
  for (int a = b + 1; a < e; a++) {
    const int32_t ko = rowOther[a]; const uint32_t ke = rowEntry[a];
    int q = a - 1;
    while (q >= b && rowOther[q] > ko) { rowOther[q + 1] = rowOther[q]; rowEntry[q + 1] = rowEntry[q]; q--; }
    rowOther[q + 1] = ko; rowEntry[q + 1] = ke;
  }
}
// short rows: all entries are loaded up front (independent loads), sorted in registers by an odd-even transposition network, written back
template <int M>
__device__ __forceinline__ void row_sort_regs(int b, int len, uint32_t* __restrict__ rowEntry, int32_t* __restrict__ rowOther) {
  int32_t ko[M]; uint32_t ke[M];
#pragma unroll
  for (int k = 0; k < M; k++) { ko[k] = k < len ? rowOther[b + k] : 0x7fffffff; ke[k] = k < len ? rowEntry[b + k] : 0u; }
#pragma unroll
  for (int r = 0; r < M; r++) {
#pragma unroll
    for (int k = (r & 1); k + 1 < M; k += 2) {
      const bool sw = ko[k] > ko[k + 1];
      const int32_t a = sw ? ko[k + 1] : ko[k], c = sw ? ko[k] : ko[k + 1]; ko[k] = a; ko[k + 1] = c;
      const uint32_t x = sw ? ke[k + 1] : ke[k], y = sw ? ke[k] : ke[k + 1]; ke[k] = x; ke[k + 1] = y;
    }
  }
#pragma unroll
  for (int k = 0; k < M; k++) if (k < len) { rowOther[b + k] = ko[k]; rowEntry[b + k] = ke[k]; }
}
__global__ void __launch_bounds__(128) k_row_sort(int n, const int* __restrict__ rowStart, uint32_t* __restrict__ rowEntry, int32_t* __restrict__ rowOther, int* __restrict__ longRows,
                                                 int* __restrict__ nLong) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = rowStart[i], len = rowStart[i + 1] - b;
  if (len > ROW_SHORT) { longRows[atomicAdd(nLong, 1)] = i; return; }
  if (len <= 1) return;
  if (len <= 8) row_sort_regs<8>(b, len, rowEntry, rowOther);
  else row_sort_regs<ROW_SHORT>(b, len, rowEntry, rowOther);
}
__global__ void __launch_bounds__(128) k_row_sort_long(const int* __restrict__ longRows, const int* __restrict__ nLong, const int* __restrict__ rowStart, uint32_t* __restrict__ rowEntry,
                                                      int32_t* __restrict__ rowOther) {
  __shared__ int32_t sK[4][ROW_WARP_MAX]; __shared__ uint32_t sV[4][ROW_WARP_MAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nl = *nLong;
  for (int q = blockIdx.x * 4 + w; q < nl; q += gridDim.x * 4) {
    const int i = longRows[q];
    const int b = rowStart[i], len = rowStart[i + 1] - b;
    if (len > ROW_WARP_MAX) { if (lane == 0) row_insertion_sort(b, b + len, rowEntry, rowOther); __syncwarp(); continue; }
    int P = 32; while (P < len) P <<= 1;
    for (int k = lane; k < P; k += 32) { sK[w][k] = k < len ? rowOther[b + k] : 0x7fffffff; sV[w][k] = k < len ? rowEntry[b + k] : 0u; }
    __syncwarp();
    for (int k = 2; k <= P; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < P; t += 32) {
          const int x = t ^ j;
          if (x > t) {
            const bool up = (t & k) == 0;
            const int32_t a = sK[w][t], c = sK[w][x];
            if ((a > c) == up) { sK[w][t] = c; sK[w][x] = a; const uint32_t v = sV[w][t]; sV[w][t] = sV[w][x]; sV[w][x] = v; }
          }
        }
        __syncwarp();
      }
    for (int k = lane; k < len; k += 32) { rowOther[b + k] = sK[w][k]; rowEntry[b + k] = sV[w][k]; }
    __syncwarp();
  }
}

// scatter per-slot data into upload order (downloads)
__global__ void __launch_bounds__(256) k_gather_state(Spheres S, int n, double* pos, double* ori, double* vel, double* angVel, double* force, double* torque) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long u = S.up[i];
  if (pos) { double4 p = S.posr[i]; pos[3 * u] = p.x; pos[3 * u + 1] = p.y; pos[3 * u + 2] = p.z; }
  if (ori) { double4 q = S.ori[i]; ori[4 * u] = q.x; ori[4 * u + 1] = q.y; ori[4 * u + 2] = q.z; ori[4 * u + 3] = q.w; }
  if (vel) { double4 v = S.velm[i]; vel[3 * u] = v.x; vel[3 * u + 1] = v.y; vel[3 * u + 2] = v.z; }
  if (angVel) { double4 w = S.angi[i]; angVel[3 * u] = w.x; angVel[3 * u + 1] = w.y; angVel[3 * u + 2] = w.z; }
  if (force) { force[3 * u] = S.fx[i]; force[3 * u + 1] = S.fy[i]; force[3 * u + 2] = S.fz[i]; }
  if (torque) { torque[3 * u] = S.tx[i]; torque[3 * u + 1] = S.ty[i]; torque[3 * u + 2] = S.tz[i]; }
}
__global__ void __launch_bounds__(256) k_scatter_state(Spheres S, int n, const int* __restrict__ up2slot, long long first, long long cnt, const double* pos, const double* vel,
                                                      const double* angVel) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const int i = up2slot[first + t];
  if (pos) { double4 p = S.posr[i]; p.x = pos[3 * t]; p.y = pos[3 * t + 1]; p.z = pos[3 * t + 2]; S.posr[i] = p; }
  if (vel) { double4 v = S.velm[i]; v.x = vel[3 * t]; v.y = vel[3 * t + 1]; v.z = vel[3 * t + 2]; S.velm[i] = v; }
  if (angVel) { double4 w = S.angi[i]; w.x = angVel[3 * t]; w.y = angVel[3 * t + 1]; w.z = angVel[3 * t + 2]; S.angi[i] = w; }
}
// ParticleContainer::remove (ParticleContainer.cpp:121-157) on the device: the kept spheres move to the front of the other generation in slot order; their upload index is
// renumbered (newUpOf: old upload index -> new one, -1 = removed), origSlot remembers the slot they had when the incidence rows were built (id2slot still holds
// it), and the removed ids vanish from id2slot so that their contacts are dropped at the next full run (k_orphans).
__global__ void __launch_bounds__(256) k_keep_flag(const int32_t* __restrict__ up, int n, const int* __restrict__ newUpOf, int* flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = newUpOf[up[i]] >= 0 ? 1 : 0;
}
// the contacts of a removed particle are deleted with it (ParticleContainer::remove -> ContactContainer::remove, ParticleContainer.cpp:121-157): every entry of
// its incidence row dies here, so that neither the history carry-over nor the orphan pass of the next full run can hand its frame / shear force to a NEW particle that
// reuses the id (ParticleContainer::findFreeId hands freed ids out again).  rows are addressed by the slot the sphere had when they were built (id2slot).
__global__ void __launch_bounds__(256) k_kill_rows(const int32_t* __restrict__ sid, int n, const int* __restrict__ keepFlag, const int* __restrict__ id2slot, const int* __restrict__ rowStart,
                                                  const uint32_t* __restrict__ rowEntry, int nRows, uint32_t* cflags, uint32_t* realBits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || keepFlag[i]) return;
  const int rs = id2slot[sid[i]];
  if (rs < 0 || rs >= nRows) return;
  for (int e = rowStart[rs]; e < rowStart[rs + 1]; e++) {
    const uint32_t c = rowEntry[e] >> 1;
    cflags[c] = (cflags[c] & ~(uint32_t)CF_REAL) | CF_DEAD | CF_CLAIMED;
    atomicAnd(&realBits[c >> 5], ~(1u << (c & 31)));
  }
}
__global__ void __launch_bounds__(256) k_compact_spheres(Spheres dst, Spheres src, int n, const int* __restrict__ flag, const int* __restrict__ off, const int* __restrict__ newUpOf,
                                                        int* __restrict__ id2slot, int* __restrict__ origSlot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = src.id[i];
  if (!flag[i]) { id2slot[id] = -1; return; }
  const int k = off[i];
  dst.posr[k] = src.posr[i]; dst.velm[k] = src.velm[i]; dst.angi[k] = src.angi[i]; dst.ori[k] = src.ori[i];
  dst.fx[k] = src.fx[i]; dst.fy[k] = src.fy[i]; dst.fz[k] = src.fz[i]; dst.tx[k] = src.tx[i]; dst.ty[k] = src.ty[i]; dst.tz[k] = src.tz[i];
  dst.lx[k] = src.lx[i]; dst.ly[k] = src.ly[i]; dst.lz[k] = src.lz[i];
  dst.flags[k] = src.flags[i]; dst.mask[k] = src.mask[i]; dst.mat[k] = src.mat[i]; dst.id[k] = id; dst.up[k] = newUpOf[src.up[i]];
  origSlot[k] = id2slot[id];
}
__global__ void __launch_bounds__(256) k_up2slot(const int32_t* __restrict__ up, int n, int* up2slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) up2slot[up[i]] = i;
}
// realBits of a freshly built table: one thread per contact, one word per warp (tables start warp-aligned)
__global__ void __launch_bounds__(256) k_realbits(const uint32_t* __restrict__ flags, long long first /*multiple of 32*/, long long n, uint32_t* __restrict__ bits) {
  const long long c = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t f = 0;
  if (c < n) f = flags[c];
  const unsigned m = __ballot_sync(0xffffffffu, (f & CF_REAL) && !(f & CF_DEAD));
  if ((threadIdx.x & 31) == 0 && c < n + 31) bits[c >> 5] = m;
}
__global__ void __launch_bounds__(256) k_count_flags(const uint32_t* __restrict__ flags, long long n, unsigned long long* out /*[0]=real,[1]=alive*/) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned r = 0, a = 0;
  if (c < n) { uint32_t f = flags[c]; r = (f & CF_REAL) ? 1 : 0; a = (f & CF_DEAD) ? 0 : 1; }
  r = __reduce_add_sync(0xffffffffu, r); a = __reduce_add_sync(0xffffffffu, a);
  if ((threadIdx.x & 31) == 0) { if (r) atomicAdd(&out[0], (unsigned long long)r); if (a) atomicAdd(&out[1], (unsigned long long)a); }
}

}  // namespace wd
