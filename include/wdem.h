/* wdem.h — C-ABI of the B200 DEM hot-path engine (libwdem.so).
 *
 * Drop-in boundary for woodem/woo's per-timestep DEM loop
 *     [Leapfrog, InsertionSortCollider, ContactLoop(Cg2_*, Cp2_*, Law2_*), DynDt]
 * (reference: py/utils.py:122-129 defaultEngines; Scene::doOneStep src/core/Scene.cpp:366-432).
 * One handle owns all device state of one DemField on one GPU.  Plain pointers and sizes only; every call
 * returns 0 on success and a non-zero code otherwise, with the message available from wdem_last_error().
 * No C++ exception crosses this boundary (the C++ engine GpuDemLoop converts non-zero to std::runtime_error,
 * which is what the reference's engines throw out of run(): src/core/Scene.cpp:110-117).
 *
 * All host arrays are row-major, float64 unless stated; vectors are xyz triples, quaternions are (x,y,z,w)
 * (Eigen coeffs() order), matrices are row-major 3x3.  Spheres are addressed by *upload index* on download.
 *
 * There is no CPU fallback: every entry point fails with WDEM_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef WDEM_H
#define WDEM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wdem_engine* wdem_handle;

enum { WDEM_OK = 0, WDEM_ERR_INVALID = 1, WDEM_ERR_CUDA = 2, WDEM_ERR_UNSUPPORTED = 3, WDEM_ERR_RUNTIME = 4 };
enum { WDEM_MAT_FRICT = 0, WDEM_MAT_HERTZ = 1 };          /* FrictMat (FrictMat.hpp:20-31), HertzMat (Hertz.hpp:9-16) */
enum { WDEM_LAW_IDEALELPL = 0, WDEM_LAW_HERTZ_DMT = 1 };  /* Law2_L6Geom_FrictPhys_IdealElPl / Law2_L6Geom_HertzPhys_DMT */
enum { WDEM_HOMO_NONE = 0, WDEM_HOMO_POS = 1, WDEM_HOMO_VEL = 2, WDEM_HOMO_VEL_2ND = 3, WDEM_HOMO_GRADV2 = 4 }; /* Cell.hpp */

/* Engine attributes; field names are the reference's attribute names.  Replaces the attribute
 * reads each reference engine does in its run():
 *   Leapfrog.{damping,reset}                     src/dem/Leapfrog.hpp:50-53
 *   InsertionSortCollider.verletDist             src/dem/Collision.hpp:94;  Bo1_Wall_Aabb.halfThick Wall.hpp
 *   Cg2_*.{noRatch,iniLensTouch,trsfRenorm}      src/dem/L6Geom.hpp:35-45   (approxMask must be 0)
 *   ContactLoop.applyForces                      src/dem/ContactLoop.hpp:115-127
 *   Law2_L6Geom_FrictPhys_IdealElPl.*            src/dem/IdealElPl.hpp:17-24
 *   Cp2_HertzMat_HertzPhys.{poisson,en}          src/dem/Hertz.hpp:58-59
 *   Law2_L6Geom_HertzPhys_DMT.noAttraction       src/dem/Hertz.hpp:96
 *   DynDt.{stepPeriod,maxRelInc}                 src/dem/DynDt.hpp:21-23, src/core/Engine.hpp:149
 *   Scene.{dt,dtSafety}, DemField.{gravity,loneMask,distFactor}   Scene.hpp:191-205, Particle.hpp:359-361 */
typedef struct wdem_config {
  double dt, dtSafety;
  double gravity[3];
  double distFactor;
  double damping;
  double verletDist;
  double wallHalfThick;
  double relRollStiff, relTwistStiff, rollTanPhi;
  double poisson, en;
  double maxRelInc;
  uint32_t loneMask;
  int32_t reset;
  int32_t noRatch, iniLensTouch, trsfRenorm;
  int32_t applyForces;
  int32_t law;
  int32_t noSlip, noBreak, noFrict;
  int32_t noAttraction;
  int32_t dynDtPeriod;
  int32_t periodic;
  int32_t device;      /* CUDA device ordinal */
  int32_t evalStress;  /* ContactLoop.evalStress (ContactLoop.hpp:117): with Scene.trackEnergy in a periodic scene the stress is evaluated every step and the work of the
                        * velocity gradient is booked as the "gradV" energy (ContactLoop.cpp:243-250); the stress itself is available on demand either way (wdem_get_stress) */
  int32_t reserved[6];
} wdem_config;

/* Counters of the last wdem_step call (cumulative ones say so). */
typedef struct wdem_stats {
  int64_t n_spheres, n_contacts, n_real;      /* after the last step */
  int64_t n_rebuilds;                         /* collider full runs inside this call */
  int64_t n_full_runs_total;                  /* InsertionSortCollider.nFullRuns (cumulative) */
  int64_t n_launches;                         /* kernels launched inside this call (ours + CUB at rebuilds) */
  int64_t step;                               /* Scene.step after the call */
  double dt, time;                            /* Scene.dt / Scene.time after the call */
  double device_ms;                           /* CUDA-event time of the whole call on the engine's stream */
  double ms_integrate, ms_contact, ms_reduce, ms_rebuild, ms_dyndt;  /* per-phase event times (if profiling enabled) */
  int64_t sum_contacts, sum_real, sum_potential_evaluated;           /* summed over the steps of this call (roofline accounting) */
} wdem_stats;

/* One Impose object (src/dem/Particle.hpp:130-160, src/dem/Impose.hpp, Impose.cpp): the hooks Leapfrog calls on a node that carries DemData.impose — velocity
 * after the accelerations were applied and again at the end of the node's step (Leapfrog.cpp:203-205,248), force at the reset (:244), readForce before the reset (:238).
 * kind and the meaning of v[16] / flags:
 *  1 HarmonicOscillation          v = freq, amp, dir(3, normalized), t0; flags&1 perpFree
 *  2 AlignedHarmonicOscillations  v = freqs(3), amps(3); NaN: axis untouched
 *  3 ConstantForce                v = F(3)
 *  4 RadialForce                  v = nodeA.pos(3), nodeB.pos(3), F
 *  5 Local6Dofs                   v = values(6), whats(6; 0 nothing, 1 velocity, 2 force); ori
 *  6 VelocityAndReadForce         v = dir(3, normalized), vel, dist (its value at upload); flags&1 latBlock, &2 invF (energyName is not supported)
 *  7 ReadForce                    flags&1: node given, v = node.pos(3), ori = node.ori
 *  8 VariableVelocity3d           v = t0; ori; times[nTimes], vals = vels (3 per time), vals2 = angVels (3 per time) or NULL; flags&1 diff, &2 wrap (hooks are not supported)
 *  9 VariableAlignedRotation      v = axis; times/vals = timeAngVel columns (1 value per time); flags&1 wrap
 * 10 CircularOrbit                v = node.pos(3), omega, -, angle (its value at upload); ori = node.ori; flags&1 rotate      (Impose.cpp:22-37)
 * 11 StableCircularOrbit          v = node.pos(3), omega, radius, angle; ori = node.ori; flags&1 rotate   (Impose.cpp:38-48; py/triangulated.py:59-64: a cylinder that really rotates)
 *    The two orbit kinds act on spheres and on facet nodes (wdem_upload_facet_imposes).  They take atan2 / sin / cos of the node's position on the device, so nodes that
 *    carry them follow the reference to rounding (1e-12 relative), not bit for bit like everything else.
 * 12 InterpolatedMotion           v = t0; times[nTimes], vals = poss (3 per time), vals2 = oris (4 per time: x,y,z,w)     (Impose.cpp:120-133)
 *    on the master nodes of kinematic facet groups only (a mesh that follows a recorded trajectory); evaluated by the host with libm: bit-identical to the oracle. */
typedef struct wdem_impose {
  int32_t kind, flags;
  double v[16];
  double ori[4];                 /* x,y,z,w */
  int32_t nTimes; int32_t pad;
  const double* times; const double* vals; const double* vals2;
} wdem_impose;

int wdem_default_config(wdem_config* cfg);
int wdem_create(const wdem_config* cfg, wdem_handle* out);
int wdem_destroy(wdem_handle h);
const char* wdem_last_error(wdem_handle h);  /* h may be NULL: message of the last failed wdem_create */
int wdem_set_config(wdem_handle h, const wdem_config* cfg);  /* re-read attributes (as the reference does each run()) */
int wdem_set_profiling(wdem_handle h, int per_phase_events);

/* Materials: replaces the shared_ptr<Material> each Particle carries (Particle.hpp:91-97). */
int wdem_upload_materials(wdem_handle h, int n, const int32_t* kind, const double* density, const double* young,
                          const double* ktDivKn, const double* tanPhi, const double* surfEnergy, const double* alpha);
/* Spheres: Node{pos,ori} + DemData{vel,angVel,mass,inertia,flags} + Sphere::radius + Particle{id,mask,material}
 * (core/Field.hpp:102-103, dem/Particle.hpp:281-292,91-97, dem/Sphere.hpp:25).  ori/vel/angVel/flags/mask may be NULL. */
int wdem_upload_spheres(wdem_handle h, int64_t n, const int32_t* ids, const double* pos, const double* ori, const double* vel,
                        const double* angVel, const double* mass, const double* inertia, const double* radius,
                        const uint32_t* flags, const uint32_t* mask, const int32_t* matId);
/* DemData.force/torque accumulated by the previous step's ContactLoop (consumed by the next Leapfrog, Leapfrog.cpp:144-160); upload order,
 * either may be NULL (= zero).  Only right after wdem_upload_spheres. */
/* ParticleContainer::add in mid-run (ParticleContainer.cpp:23-73; what Inlets do, Inlet.cpp): append spheres to the device state without re-uploading
 * the scene.  Same arrays as wdem_upload_spheres; ids are required and must be unused.  Existing spheres, their forces and the whole contact history stay;
 * the new spheres start with zero force and the collider re-initialises at the next step (contacts dirty).  Upload order continues:
 * the new spheres follow the old ones in wdem_download_state / wdem_set_state.
 * Decomposed runs: call on EVERY rank with the SAME arguments (collective in spirit, no communication inside).  Each rank keeps the spheres whose centre lies in
 * its slab (periodic cells: every world-th sphere of the list) and the full run that follows sends every sphere home and exports its ghosts.  The id check then
 * covers what the subdomain holds (owned spheres, ghost copies, boundaries); global uniqueness is the caller's business. */
int wdem_add_spheres(wdem_handle h, int64_t n, const int32_t* ids, const double* pos, const double* ori, const double* vel, const double* angVel,
                     const double* mass, const double* inertia, const double* radius, const uint32_t* flags, const uint32_t* mask,
                     const int32_t* matId);
/* ParticleContainer::remove in mid-run (ParticleContainer.cpp:121-157; what Outlets do): the spheres with these ids leave the device state, the remaining ones keep
 * their state, their forces and their contact history; contacts of the removed spheres are dropped and the collider re-initialises at the next step.  The upload
 * order closes up (remaining spheres keep their relative order).
 * Decomposed runs: call on EVERY rank with the SAME list; each rank lets go of what it holds of it (owned spheres and ghost copies) and ignores the ids it does not hold. */
int wdem_remove_spheres(wdem_handle h, int64_t n, const int32_t* ids);
int wdem_upload_forces(wdem_handle h, const double* force, const double* torque);
/* Walls (dem/Wall.hpp): static; vel may be NULL. */
int wdem_upload_walls(wdem_handle h, int n, const int32_t* ids, const double* pos, const int32_t* axis, const int32_t* sense,
                      const double* vel, const uint32_t* mask, const int32_t* matId);
/* Facets (dem/Facet.hpp): verts 9 per facet, vvel 9 per facet (node velocities; NULL = 0), fakeVel 3, n21lim 3 (NaN = none). */
int wdem_upload_facets(wdem_handle h, int n, const int32_t* ids, const double* verts, const double* vvel, const double* fakeVel,
                       const double* halfThick, const double* n21lim, const uint32_t* mask, const int32_t* matId);
/* Periodic cell (core/Cell.hpp): Cell.setHSize + gradV/nextGradV/homoDeform. */
/* Kinematic groups of facets: a clump of boundary particles whose massless, fully blocked master node carries a prescribed velocity and angular velocity
 * (ParticleContainer::addClumped + ClumpData, Clump.cpp:187-261; examples/mill.py:38-42: a rotating drum).  Every step the master node is moved by Leapfrog's
 * fixed-particle branch (Leapfrog.cpp:195-199,275-297) and its members follow rigidly (ClumpData::applyToMembers, Clump.cpp:293-307): node position
 * pos + ori*relPos, node velocity vel + angVel x (p - pos); the facets' Aabbs take part in the Verlet test and the collider run like any other particle.
 * groupOfFacet: per uploaded facet the group index or -1 (static); pos 3, ori 4 (x,y,z,w), vel 3, angVel 3 per group; at most 8 groups, aperiodic scenes.
 * Call after wdem_upload_facets, before the first step.  Forces on the members are not fed back (all DOFs of the master are blocked). */
int wdem_upload_facet_groups(wdem_handle h, int nGroups, const int32_t* groupOfFacet, const double* pos, const double* ori, const double* vel,
                             const double* angVel);
/* Clumps of spheres with mass (ClumpData, src/dem/Clump.hpp; ParticleContainer::addClumped, S.dem.par.addClumped): rigid aggregates whose master node carries
 * the mass and the principal inertia.  Replaces, for sphere members, what Leapfrog::run does with a clump node (Leapfrog.cpp:148-255): every step
 * ClumpData::forceTorqueFromMembers (Clump.cpp:272-280), the spherical or — when the principal inertias differ — the aspherical integrator
 * (leapfrogAsphericalRotate, Leapfrog.cpp:299-321), the reset (gravity acts on the master node only) and ClumpData::applyToMembers (Clump.cpp:293-307).
 * Members of one clump never collide (Collider::mayCollide, Collision.cpp:34-39); DynDt weighs the members' contact stiffnesses against the clump's mass
 * and inertia (DynDt.cpp:48-52).
 * memberStart: nClumps+1 offsets into memberIds (particle ids of spheres uploaded with the CLUMPED flag, DemData.flags bit 64, in ClumpData.nodes order);
 * pos 3, ori 4 (x,y,z,w), vel 3, angVel 3, angMom 3 (NULL or NaN: initialised by the first step, Leapfrog.cpp:303-305), mass 1, inertia 3 (principal), flags
 * (DemData.flags of the master node: blocked DOFs, *_SKIP bits) per clump; relPos 3 / relOri 4 per member (ClumpData.relPos/relOri; NULL: derived from the
 * uploaded member state as ClumpData::makeClump does, Clump.cpp:252-253).  Call right after wdem_upload_spheres; every sphere with the CLUMPED flag must be a
 * member of exactly one clump.  Single-device scenes (aperiodic or periodic) without Scene.trackEnergy; wdem_add_spheres / wdem_remove_spheres are refused while clumps exist. */
int wdem_upload_clumps(wdem_handle h, int32_t nClumps, const int32_t* memberStart, const int32_t* memberIds, const double* pos, const double* ori,
                       const double* vel, const double* angVel, const double* angMom, const double* mass, const double* inertia, const uint32_t* flags,
                       const double* relPos, const double* relOri);
/* State of the clump master nodes in upload order: 3 doubles per clump (ori: 4, x,y,z,w); any pointer may be NULL.  force/torque are what the next step's
 * forceTorqueFromMembers adds to (mass*gravity after Leapfrog's reset). */
int wdem_download_clumps(wdem_handle h, double* pos, double* ori, double* vel, double* angVel, double* angMom, double* force, double* torque);
int64_t wdem_num_clumps(wdem_handle h);
/* DemData.force / torque of the master nodes at upload time (3 per clump; either may be NULL = zero): what a scene that has already run carries there
 * (mass*gravity after Leapfrog's reset).  Only right after wdem_upload_clumps. */
int wdem_upload_clump_forces(wdem_handle h, const double* force, const double* torque);
/* DemData.impose of the scene's nodes.  imp[nImp]: the Impose objects; a node refers to a SET of them (CombinedImpose, Impose.cpp:283-300: applied in order):
 * set s = setOps[setStart[s] .. setStart[s+1]).  Spheres: sphereSet[k] is the set of the sphere with particle id sphereIds[k]; wallSet / groupSet: per uploaded wall /
 * per facet group, -1 none, NULL = none at all.  A wall that carries an imposition is integrated like one added with a velocity (DemData::guessMoving, Particle.cpp:638-640).
 * Walls and facet-group master nodes (all DOFs blocked) take velocity impositions only; the read-force kinds work on spheres, single-device runs.
 * Time-dependent factors are evaluated by the host with libm once per step (they depend on Scene.time and Scene.dt only).  Call after the particles are uploaded, before
 * the first step; at most 16 Impose objects.  Not with wdem_step_e2e. */
int wdem_upload_imposes(wdem_handle h, int32_t nImp, const wdem_impose* imp, int32_t nSets, const int32_t* setStart, const int32_t* setOps, int64_t nSphereIds,
                        const int32_t* sphereIds, const int32_t* sphereSet, const int32_t* wallSet, const int32_t* groupSet);
/* DemData.impose of facet NODES: facetSet[f] = the set of impositions all three nodes of uploaded facet f carry, -1 none.  Such nodes are in DemField.nodes
 * (DemData::guessMoving, Particle.cpp:638-640): Leapfrog's fixed-node branch translates them with their velocity and Impose::velocity at the end of the node's step sets
 * the velocity of the next one (Leapfrog.cpp:195-199,248).  Velocity impositions only; not for facets of a kinematic group.  Call after wdem_upload_imposes. */
int wdem_upload_facet_imposes(wdem_handle h, const int32_t* facetSet);
/* DemData.impose of the master nodes of clumps of spheres: clumpSet[c] = the set of impositions of uploaded clump c, -1 none.  Leapfrog calls the hooks on the master
 * node like on any other node (velocity after the accelerations and at the end of the node's step, before ClumpData::applyToMembers; force at the reset;
 * Leapfrog.cpp:203-205,244,248,252).  All kinds but the read-force ones and InterpolatedMotion.  Call after wdem_upload_clumps and wdem_upload_imposes. */
int wdem_upload_clump_imposes(wdem_handle h, const int32_t* clumpSet);
/* Velocities of the boundary nodes an imposition may have changed: 3 per uploaded wall, 3 + 3 per facet group; any pointer may be NULL. */
int wdem_download_boundary_velocities(wdem_handle h, double* wallVel, double* groupVel, double* groupAngVel);
/* out[8]: VelocityAndReadForce.sumF, .dist (CircularOrbit / StableCircularOrbit: .angle in the place of .dist); ReadForce.F(3), .T(3) of Impose object ix after the last step (sums of what the last step's Leapfrog read). */
int wdem_get_impose_state(wdem_handle h, int32_t ix, double* out);
/* Facet node positions (9 per facet, upload order) and the groups' master nodes after the last step; any pointer may be NULL. */
int wdem_download_facets(wdem_handle h, double* verts, double* groupPos, double* groupOri);
int wdem_set_cell(wdem_handle h, const double* hSize, const double* gradV, const double* nextGradV, int homoDeform);
int wdem_set_next_gradv(wdem_handle h, const double* nextGradV);
int wdem_get_cell(wdem_handle h, double* hSize);

/* nsteps times Scene::doOneStep restricted to this engine group (Scene.cpp:366-432): cell.integrateAndUpdate,
 * Leapfrog, collider, ContactLoop, DynDt, epilogue (step++, time+=dt, dt=nextDt). stats may be NULL. */
int wdem_step(wdem_handle h, int64_t nsteps, wdem_stats* stats);
/* End-to-end variant (host buffers in and out; what a host program that owns the particle state calls every step, cf. Scene::doOneStep with the state in
 * DemField.nodes): uploads pos/vel/angVel (host, upload order), steps, downloads pos/vel (+ kinetic energy).  The copies overlap with the integrator chunk by chunk.
 * With stats == NULL (and nsteps == 1) the call returns as soon as the step's positions and velocities are on the host: the contact pass that prepares the forces
 * of the NEXT step is still running and overlaps the next call's upload.  Every other entry point waits for it as usual; an error raised by that contact pass
 * (NaN force, ...) is reported by the next call.  With stats != NULL the call waits for the whole step (the counters are part of the statistics). */
int wdem_step_e2e(wdem_handle h, int64_t nsteps, const double* pos_in, const double* vel_in, const double* angVel_in,
                  double* pos_out, double* vel_out, double* ek_out, wdem_stats* stats);

/* The same for one subdomain of a decomposed run (also accepted by a single-domain engine): the host sends the state of the m spheres this handle OWNS, addressed by
 * Particle.id in any order, and receives their new pos / vel in the same order.  m must equal wdem_num_owned.  *owned_changed != 0 (may be NULL): spheres migrated in
 * this step's collider run; the outputs are still those of the spheres sent, the next call must send the new owned set (wdem_download_owned gives it). */
int wdem_step_e2e_owned(wdem_handle h, int64_t m, const int32_t* ids, const double* pos_in, const double* vel_in, const double* angVel_in,
                        double* pos_out, double* vel_out, double* ek_out, int32_t* owned_changed, wdem_stats* stats);

/* Downloads, in upload order; any pointer may be NULL. */
int wdem_download_state(wdem_handle h, double* pos, double* ori, double* vel, double* angVel);
int wdem_download_forces(wdem_handle h, double* force, double* torque);
int wdem_set_state(wdem_handle h, int64_t first, int64_t n, const double* pos, const double* vel, const double* angVel);
int64_t wdem_num_contacts(wdem_handle h, int realOnly);
/* Contacts sorted by canonical key (id1,id2): boundary first, sphere-sphere lower id first. Returns the count (<0: error).
 * ids 2/contact, cellDist 3, real 1, uN 1, forceLocal 3, torqueLocal 3, trsf 9 (column-major), contPt 3, phys 3 (kn,kt,tanPhi),
 * stepCreated 1, physRaw 5 (the stored constants, see wdem_upload_contacts). Any pointer may be NULL. */
int64_t wdem_download_contacts(wdem_handle h, int realOnly, int64_t cap, int32_t* ids, int32_t* cellDist, int32_t* real, double* uN,
                               double* forceLocal, double* torqueLocal, double* trsf, double* contPt, double* phys,
                               int32_t* stepCreated, double* physRaw);
/* Import the persistent state of existing REAL contacts (resume from a saved Scene, or keep the contact history when the particle set is
 * re-uploaded): Contact{cellDist,stepCreated} (Contact.hpp:95-106), L6Geom{trsf,uN}+CGeom.node.pos (L6Geom.hpp:16-21, Contact.hpp:21),
 * FrictPhys{force,torque,kn,kt,tanPhi} (Contact.hpp:32-33, FrictMat.hpp:33-35) / HertzPhys (Hertz.hpp:20-26).  Same layout and canonical order
 * as wdem_download_contacts (boundary particle first; sphere-sphere: id1<id2); physRaw is 5 doubles per contact exactly as downloaded
 * (linear law: kn,kt,tanPhi,0,0; Hertz: K*sqrt(R), kt0, tanPhi, alpha_sqrtMK, 4*pi*R*gamma).  Call after the particles are uploaded and before
 * the first wdem_step; the rows are matched to the candidate pairs of the first collider run by (id1,id2) — a real contact whose boxes no
 * longer overlap is kept, as the reference's collider never deletes a real contact (InsertionSortCollider.cpp:78-113).  cellDist, stepCreated and
 * torqueLocal may be NULL. */
int wdem_upload_contacts(wdem_handle h, int64_t n, const int32_t* ids, const int32_t* cellDist, const int32_t* stepCreated, const double* uN,
                         const double* forceLocal, const double* torqueLocal, const double* trsf, const double* contPt, const double* physRaw);
/* Scene.step / Scene.time (Scene.hpp:191-205) of a resumed scene: trsfRenorm and DynDt's period count scene steps. */
int wdem_set_scene(wdem_handle h, int64_t step, double time);
/* PeriodicEngine.stepLast / .stepPrev of DynDt (Engine.hpp:156-157, Engine.cpp:170-195): the step of DynDt's last activation and of the one before, -1 = never ran.
 * wdem_set_scene resets both to -1 (a fresh engine); set them after it to resume a run exactly where a saved scene left it — otherwise DynDt's period restarts at the
 * resumed step.  Either output pointer may be NULL. */
int wdem_set_periodic_state(wdem_handle h, int64_t dynDtStepLast, int64_t dynDtStepPrev);
int wdem_get_periodic_state(wdem_handle h, int64_t* dynDtStepLast, int64_t* dynDtStepPrev);
/* Walls after the last step, in upload order (any pointer may be NULL): position (a wall uploaded with a nonzero velocity is in DemField.nodes and is
 * translated by Leapfrog, pos += vel*dt — Leapfrog.cpp:195-199,275-278, Particle.cpp:638-640) and DemData.force / torque as the reference's scripts see them
 * (ContactLoop::applyForceUninodal, ContactLoop.cpp:270-276): for a MOVING wall the last step's contacts (Leapfrog resets its force every step; summed in ascending id
 * of the sphere, bit-identical to the oracle); for a wall AT REST the value accumulated over all steps since wdem_upload_walls — the reference never resets it
 * (the node is not in DemField.nodes); accumulated on the device per step, equal to the reference to rounding (its own order of additions is that of
 * its contact loop).  Decomposed runs: the share of this subdomain's spheres. */
int wdem_download_walls(wdem_handle h, double* pos, double* force, double* torque);
/* Scene.trackEnergy (Scene.hpp:191-205, core/EnergyTracker.hpp): on = the step kernels also accumulate the EnergyTracker entries of this path, 8 doubles in the order
 * kinetic, grav, nonviscDamp (Leapfrog.cpp:53-98), elast, plast, broken (IdealElPl.cpp:28-40,64-75,85-92), viscN, viscT (Hertz.cpp:181,193,199,203-207; Hertz's
 * sliding and elastic terms go to plast / elast).  "kinetic" and "elast" are resettable: they hold the last step's value (Scene.cpp:416), the others accumulate from
 * the moment tracking was switched on.  Sums are formed with atomics: equal to the reference's to rounding, not bitwise.  The gradV work of periodic scenes
 * (ContactLoop.cpp:243-250) is a separate entry: wdem_get_gradv_work. */
int wdem_track_energy(wdem_handle h, int on);
int wdem_get_energy(wdem_handle h, double* out8);
/* EnergyTracker entry "gradV" (ContactLoop.cpp:243-250): -tr(gradV . midStress) dt midVol accumulated over the steps taken with trackEnergy and evalStress in a
 * periodic scene (midStress / midVol: means of this and the previous step's stress / cell volume).  Decomposed runs: this subdomain's share. */
int wdem_get_gradv_work(wdem_handle h, double* work);
/* ContactLoop.stress with evalStress=True (ContactLoop.hpp:117, ContactLoop.cpp:116,123,218-231,243-244): sum over real contacts of F (x) dPos / cell volume,
 * row-major 3x3, periodic scenes only; evaluated on demand from the state the last step left (decomposed runs: this subdomain's share — add the
 * results of all subdomains).  Contacts with facets fail like the reference (multi-nodal particles). */
int wdem_get_stress(wdem_handle h, double* stress9);
/* DemField::critDt -> DemFuncs::pWaveDt (Funcs.cpp:151-175). */
int wdem_crit_dt(wdem_handle h, double* dt);
/* Σ ½ m v² + ½ ω·I·ω  (DemData::getEk_any, Particle.cpp:247-265), device reduction. */
int wdem_kinetic_energy(wdem_handle h, double* ek);
int wdem_get_scalars(wdem_handle h, double* dt, double* time, int64_t* step, double* verletDist, int64_t* nFullRuns);
/* ---- Slab decomposition over several GPUs.  The reference has no distributed DEM; its nearest interface is the shared-memory
 * loop over all particles/contacts (`#pragma omp parallel for`, Leapfrog.cpp:144-147, ContactLoop.cpp:134-137), which this replaces for scenes that do not
 * fit / should not run on one GPU.  One handle = one subdomain (1-D slab along `axis`, periodic scenes: along the cell axis, ring-closed).
 * Every subdomain uploads the walls/facets/materials of the whole scene and ANY subset of the spheres with scene-wide unique ids (the first collider
 * run migrates them to their owners); then all subdomains call wdem_step with the same arguments.  Per step the owners refresh pos/vel/angVel of
 * their ghosts (72 B each); at collider full runs spheres migrate together with the history of their contacts and ghosts are re-selected.
 * Transport: NCCL (one process per GPU; the 128-byte id from rank 0 must be distributed by the caller, e.g. torch.distributed) or an in-process
 * group (one host thread per handle; any mix of devices, also several subdomains on one GPU). */
int wdem_comm_unique_id(char* out128);
int wdem_comm_init_nccl(wdem_handle h, int rank, int world, const char* id128);
void* wdem_local_group_create(int world);
int wdem_local_group_destroy(void* group);
int wdem_comm_init_local(wdem_handle h, void* group, int rank);
/* cuts: world-1 ascending slab boundaries along axis (absolute coordinates; periodic: fractions of the cell size). globalMaxId: largest Particle.id of the scene. */
int wdem_set_domain(wdem_handle h, int axis, int ncuts, const double* cuts, int64_t globalMaxId);
/* Re-balancing at collider full runs (quantile cuts at rebuild time).  When, at a full run, the largest subdomain owns more than (1+tolerance) times the
 * mean number of spheres — and at least minFullRuns full runs have passed since the last re-balancing — the cuts move to the k/world quantiles of the slab coordinate
 * of all spheres (a 2048-bin histogram over the global extent, summed over the ranks) and the spheres migrate accordingly in the same full run.  Results do not depend on
 * the cuts (N subdomains are bit-identical to one).  tolerance 0 switches it off (the default).  Collective: call with the same values on every rank.
 * wdem_get_domain: the current cuts (world-1 values) and how many times they were moved; either pointer may be NULL. */
int wdem_set_rebalance(wdem_handle h, double tolerance, int minFullRuns);
int wdem_get_domain(wdem_handle h, double* cuts, int64_t* nRebalances);
int64_t wdem_num_owned(wdem_handle h);
/* Owned spheres in subdomain order with their ids (any pointer but ids may be NULL); arrays sized by wdem_num_owned. */
int wdem_download_owned(wdem_handle h, int32_t* ids, double* pos, double* ori, double* vel, double* angVel, double* force, double* torque);
/* Overwrite pos/vel/angVel (any may be NULL) of owned spheres addressed by Particle.id; fails if an id is not owned here. */
int wdem_set_owned(wdem_handle h, int64_t n, const int32_t* ids, const double* pos, const double* vel, const double* angVel);
int wdem_domain_stats(wdem_handle h, int64_t* nOwned, int64_t* nGhosts, int64_t* migratedOut, int64_t* migratedIn, int64_t* rowsIn);
/* 1 if the per-step halo / dirty-flag exchange of this subdomain runs over NVLink peer memory (CUDA-IPC mailboxes), 0 if it uses the NCCL calls
   (single-process transport, WDEM_NO_PEER, or a peer that could not be mapped); -1 on a bad handle. */
int wdem_peer_active(wdem_handle h);

/* Device-level info for bench.py: bytes allocated, stream pointer (cudaStream_t) */
int wdem_device_info(wdem_handle h, int64_t* bytes_allocated, void** stream);

#ifdef __cplusplus
}
#endif
#endif
