// wdem_engine.cu — host orchestration + C-ABI (include/wdem.h) of the B200 DEM hot-path engine.
// One step = Scene::doOneStep restricted to [Leapfrog, InsertionSortCollider, ContactLoop, DynDt]
// (reference: src/core/Scene.cpp:366-432, py/utils.py:122-129).  All compute is in wdem_kernels.cuh; CUB is used only
// inside collider full runs (radix sort / scans).  No CPU fallback: without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>
#include "../../include/wdem.h"
#include <chrono>
#include <unordered_map>
#include <unordered_set>
#include "wdem_kernels.cuh"
#include "wdem_domain.cuh"
#include "wdem_comm.h"

using namespace wd;

namespace {
thread_local std::string g_createError;

struct CudaErr { std::string msg; int code; };
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess) throw CudaErr{std::string(#call) + ": " + cudaGetErrorString(e__), WDEM_ERR_CUDA}; \
  } while (0)
struct EngErr { std::string msg; int code; };

// Growth of the device buffers is stream-ordered (cudaMallocAsync / cudaFreeAsync on the engine's stream, from the device's default pool, whose release threshold
// wdem_create raises so that freed blocks stay in the pool): while a bed forms, the contact tables of a small scene outgrow their capacity dozens of times, and
// cudaMalloc + cudaFree (a device-wide synchronisation each) made such a collider run cost 1 ms instead of 0.26 ms (10k spheres).  Everything the engine does with a
// buffer is ordered on that one stream (side streams are joined before a full run; the end-to-end staging buffers grow only between calls), so freeing in stream order
// is safe.  Buffers handed to NCCL stay with cudaMalloc (`plain`).
thread_local cudaStream_t g_allocStream = nullptr;      // the stream of the handle whose entry point is running (guarded())
static const bool g_noPool = getenv("WDEM_NO_POOL") != nullptr;
template <class T>
struct DBuf {  // growable device buffer
  T* p = nullptr; size_t cap = 0; bool plain = false, pooled = false;
  void ensure(size_t n, int64_t& total, bool keep = false, cudaStream_t st = 0) {
    if (n <= cap) return;
    size_t ncap = std::max(n, cap + cap / 2);
    T* q = nullptr;
    cudaStream_t s = st ? st : g_allocStream;
    const bool pool = !plain && !g_noPool && s != nullptr;
    if (pool) CK(cudaMallocAsync(&q, ncap * sizeof(T), s)); else CK(cudaMalloc(&q, ncap * sizeof(T)));
    if (keep && p && cap) CK(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s ? s : st));
    if (p) {
      if (pooled && s) CK(cudaFreeAsync(p, s));
      else { CK(cudaStreamSynchronize(s ? s : st)); CK(cudaFree(p)); }
      total -= (int64_t)(cap * sizeof(T));
    }
    p = q; cap = ncap; pooled = pool; total += (int64_t)(ncap * sizeof(T));
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; pooled = false; }
};

struct SphereStore {
  DBuf<double4> posr, velm, angi, ori; DBuf<double> f[6], l[3]; DBuf<uint32_t> flags, mask; DBuf<int32_t> mat, id, up;
  void ensure(size_t n, int64_t& tot, bool keep = false, cudaStream_t st = 0) {
    posr.ensure(n, tot, keep, st); velm.ensure(n, tot, keep, st); angi.ensure(n, tot, keep, st); ori.ensure(n, tot, keep, st);
    for (auto& b : f) b.ensure(n, tot, keep, st); for (auto& b : l) b.ensure(n, tot, keep, st);
    flags.ensure(n, tot, keep, st); mask.ensure(n, tot, keep, st); mat.ensure(n, tot, keep, st); id.ensure(n, tot, keep, st); up.ensure(n, tot, keep, st);
  }
  Spheres view() { Spheres S; S.posr = posr.p; S.velm = velm.p; S.angi = angi.p; S.ori = ori.p; S.fx = f[0].p; S.fy = f[1].p; S.fz = f[2].p; S.tx = f[3].p; S.ty = f[4].p; S.tz = f[5].p;
    S.lx = l[0].p; S.ly = l[1].p; S.lz = l[2].p; S.flags = flags.p; S.mask = mask.p; S.mat = mat.p; S.id = id.p; S.up = up.p; S.clumpOfId = nullptr; S.impOfId = nullptr; return S; }
  void release() { posr.release(); velm.release(); angi.release(); ori.release(); for (auto& b : f) b.release(); for (auto& b : l) b.release(); flags.release(); mask.release(); mat.release(); id.release(); up.release(); }
};

struct ContactStore {
  DBuf<int2> pair; DBuf<uint32_t> flags, realBits; DBuf<int32_t> stepCreated, cd[3]; DBuf<double4> r[4], ph, tq, Tg; DBuf<double> ph4;
  size_t cap = 0;
  void ensure(size_t n, int64_t& tot, bool hertz, bool rolling, bool periodic, bool keep = false, cudaStream_t st = 0) {
    if (n <= cap) return;
    n = std::max(n, cap + cap / 2);
    pair.ensure(n, tot, keep, st); flags.ensure(n, tot, keep, st); stepCreated.ensure(n, tot, keep, st);
    if (periodic) for (auto& b : cd) b.ensure(n, tot, keep, st);
    for (auto& b : r) b.ensure(n, tot, keep, st); ph.ensure(n, tot, keep, st);
    if (hertz) ph4.ensure(n, tot, keep, st);
    if (rolling) { tq.ensure(n, tot, keep, st); Tg.ensure(n, tot, keep, st); }
    realBits.ensure(n / 32 + 2, tot, keep, st);
    cap = n;
  }
  Contacts view() { Contacts C; C.pair = pair.p; C.flags = flags.p; C.stepCreated = stepCreated.p; C.cdx = cd[0].p; C.cdy = cd[1].p; C.cdz = cd[2].p;
    C.rA = r[0].p; C.rB = r[1].p; C.rC = r[2].p; C.rD = r[3].p; C.ph = ph.p; C.ph4 = ph4.p; C.tq = tq.p; C.Tg = Tg.p; C.realBits = realBits.p; return C; }
  void release() { pair.release(); flags.release(); stepCreated.release(); for (auto& b : cd) b.release(); for (auto& b : r) b.release(); ph.release(); ph4.release(); tq.release(); Tg.release(); realBits.release(); cap = 0; }
};

// clumps of spheres with mass (wdem_upload_clumps): master nodes integrated by k_clumps_step
struct ClumpStore {
  int n = 0, nMem = 0;
  DBuf<double4> pos, ori, vel, angVel, angMom, force, torque, inertia, relPos, relOri; DBuf<uint32_t> flags; DBuf<int> memStart, impSet; bool anyImp = false; DBuf<int32_t> memId, clumpOfId;
  Clumps view() { Clumps K; K.n = n; K.pos = pos.p; K.ori = ori.p; K.vel = vel.p; K.angVel = angVel.p; K.angMom = angMom.p; K.force = force.p; K.torque = torque.p; K.inertia = inertia.p; K.flags = flags.p;
    K.memStart = memStart.p; K.memId = memId.p; K.relPos = relPos.p; K.relOri = relOri.p; K.impSet = anyImp ? impSet.p : nullptr; return K; }
  void release() { for (auto* b : {&pos, &ori, &vel, &angVel, &angMom, &force, &torque, &inertia, &relPos, &relOri}) b->release(); flags.release(); memStart.release(); impSet.release(); anyImp = false; memId.release(); clumpOfId.release(); n = nMem = 0; }
};

struct CsrStore { DBuf<int> rowStart; DBuf<uint32_t> rowEntry; DBuf<int32_t> rowOther; bool valid = false; int nRows = 0; void release() { rowStart.release(); rowEntry.release(); rowOther.release(); } };

// ---- host-side 3x3 helpers for the periodic cell (src/core/Cell.cpp:81-120), row-major m33
m33 mIdent() { m33 r; for (int i = 0; i < 9; i++) r.m[i] = (i % 4 == 0) ? 1. : 0.; return r; }
m33 mZero() { m33 r; for (int i = 0; i < 9; i++) r.m[i] = 0.; return r; }
m33 mAdd(const m33& a, const m33& b) { m33 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i]; return r; }
m33 mSub(const m33& a, const m33& b) { m33 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] - b.m[i]; return r; }
m33 mScale(const m33& a, double s) { m33 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] * s; return r; }
m33 mT(const m33& a) { m33 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * j + i]; return r; }
m33 mMul(const m33& a, const m33& b) {   // Matrix3*Matrix3: column j of the result = a * (column j of b), row order per WD_MATPROD_ORDER
  m33 r;
  for (int j = 0; j < 3; j++) { const d3 c = mulMV(a, mk3(b.m[j], b.m[3 + j], b.m[6 + j])); r.m[j] = c.x; r.m[3 + j] = c.y; r.m[6 + j] = c.z; }
  return r;
}
double mDet(const m33& a) { const double* m = a.m; return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]); }
m33 mInv(const m33& a) {
  const double* m = a.m; double id = 1. / mDet(a); m33 r;
  r.m[0] = (m[4] * m[8] - m[5] * m[7]) * id; r.m[1] = (m[2] * m[7] - m[1] * m[8]) * id; r.m[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  r.m[3] = (m[5] * m[6] - m[3] * m[8]) * id; r.m[4] = (m[0] * m[8] - m[2] * m[6]) * id; r.m[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  r.m[6] = (m[3] * m[7] - m[4] * m[6]) * id; r.m[7] = (m[1] * m[6] - m[0] * m[7]) * id; r.m[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return r;
}
d3 hLevi(const m33& a) { return mk3(a.m[5] - a.m[7], a.m[6] - a.m[2], a.m[1] - a.m[3]); }  // Math.hpp:223-228

struct HostCell {
  m33 hSize = mIdent(), trsf = mIdent(), gradV = mZero(), nextGradV = mZero(), pprevHsize = mIdent();
  d3 spinVec = mk3(0, 0, 0), size = mk3(1, 1, 1), cosv = mk3(1, 1, 1); m33 unshearTrsf = mIdent(); bool hasShear = false; int homoDeform = WDEM_HOMO_GRADV2;
  void integrateAndUpdate(double dt) {  // Cell.cpp:81-120
    m33 inc = mScale(gradV, dt);
    trsf = mAdd(trsf, mMul(inc, trsf));
    m33 prev = hSize;
    if (homoDeform == WDEM_HOMO_GRADV2) {
      m33 A = mInv(mSub(mIdent(), mScale(mScale(gradV, dt), .5)));
      m33 B = mAdd(mIdent(), mScale(mScale(gradV, dt), .5));
      hSize = mMul(mMul(A, B), hSize);
    } else hSize = mAdd(hSize, mMul(inc, hSize));
    pprevHsize = mScale(mAdd(prev, hSize), .5);
    if (mDet(hSize) == 0) throw EngErr{"Cell is degenerate (zero volume).", WDEM_ERR_RUNTIME};
    m33 Hn; d3 hcol[3];
    for (int i = 0; i < 3; i++) {
      d3 b = mk3(hSize.m[i], hSize.m[3 + i], hSize.m[6 + i]); double l = nrm(b); setc(size, i, l);
      b = b / l; hcol[i] = b; Hn.m[i] = b.x; Hn.m[3 + i] = b.y; Hn.m[6 + i] = b.z;
    }
    for (int i = 0; i < 3; i++) setc(cosv, i, sqn(cross(hcol[(i + 1) % 3], hcol[(i + 2) % 3])));
    unshearTrsf = mInv(Hn);   // _shearTrsf=Hnorm; _unshearTrsf=_shearTrsf.inverse() (Cell.cpp:104-106)
    hasShear = (hSize.m[1] != 0 || hSize.m[2] != 0 || hSize.m[3] != 0 || hSize.m[5] != 0 || hSize.m[6] != 0 || hSize.m[7] != 0);
  }
  void setNextGradV() { gradV = nextGradV; m33 W = mScale(mSub(gradV, mT(gradV)), .5); spinVec = .5 * hLevi(W); }  // Cell.cpp:60-66
};
}  // namespace

// Slab decomposition state of one subdomain; inactive (comm == nullptr) in single-domain runs.
struct DomainState {
  int axis = 0; std::vector<double> cuts;   // world-1 interior boundaries: absolute coordinates, or fractions of the cell size if periodic
  int nOwn = 0; int globalMaxId = -1; bool globalsValid = false; double rmaxGlobal = 0;
  std::vector<size_t> sendCnt, recvCnt; size_t nSendTot = 0, nRecvTot = 0;   // ghosts exported to / imported from every rank
  DBuf<int> sendSlot, recvSlot; DBuf<int32_t> exportId, importId, oldIds; DBuf<double> haloSend, haloRecv;
  DBuf<uint32_t> dkeys, dkeysAlt, emask; DBuf<int> dslots, dslotsAlt, dstart, origSlot, perm2, rowCnt, rowOff, perRank, flag, off, syncFlag;
  DBuf<char> sendBuf, recvBuf; DBuf<unsigned long long> immKeys, immKeysAlt; DBuf<int> immVals, immValsAlt; int nImm = 0;
  long long migratedOut = 0, migratedIn = 0, rowsIn = 0;
  // per-step exchange over peer memory (k_halo_push / k_peer_sync): my mailbox (words + 2 halo buffers), the peers' mailboxes mapped with CUDA IPC
  // re-balancing (wdem_set_rebalance): when the largest subdomain owns more than (1+rebalTol) x the mean at a full run, the cuts move to the quantiles of the slab coordinate
  double rebalTol = 0.; int rebalMinRuns = 1; long long lastRebalRun = -1, nRebalances = 0; DBuf<int> hist;
  int deadSlots = 0;      // slots of the current generation flagged F_DEADSLOT (domainExchange's in-place path), cut off by the collider run that follows
  struct Peer { bool active = false, disabled = false; unsigned long long syncSeq = 0; char* box = nullptr; long long haloCap = 0; void* base[WDEM_MAX_RANKS] = {}; std::vector<long long> dstOff; long long maxRecvTot = 0; } peer;
  DomainState() { sendBuf.plain = true; recvBuf.plain = true; perRank.plain = true; haloSend.plain = true; haloRecv.plain = true; }      // handed to NCCL / read by peers: plain cudaMalloc
  void release() {
    sendSlot.release(); recvSlot.release(); exportId.release(); importId.release(); oldIds.release(); haloSend.release(); haloRecv.release(); dkeys.release(); dkeysAlt.release();
    emask.release(); dslots.release(); dslotsAlt.release(); dstart.release(); origSlot.release(); perm2.release(); rowCnt.release(); rowOff.release(); perRank.release();
    flag.release(); off.release(); syncFlag.release(); sendBuf.release(); recvBuf.release(); immKeys.release(); immKeysAlt.release(); immVals.release(); immValsAlt.release();
  }
};

struct wdem_engine {
  wdem_config cfg;
  Transport* comm = nullptr; DomainState dom;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;       // boundary (wall/facet/orphan) contact kernels run beside the sphere-sphere kernel
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evFlag = nullptr, evFork = nullptr, evJoin = nullptr, evOrph = nullptr;
  // steps are enqueued in batches of up to batchK without a host wait (Counters::gflag); batchK adapts to the number of steps between collider full runs
  int batchMax = 16, batchK = 1; long long stepsSinceFullRun = 0;
  static const int NEVSET = 4;          // per-phase events are harvested a few steps later, so profiling does not stall the pipeline
  cudaEvent_t evr[NEVSET][6] = {}; bool evPending[NEVSET] = {false, false, false, false}; double phaseAcc[5] = {0, 0, 0, 0, 0};
  std::string err;
  int64_t bytes = 0;
  bool profiling = false;
  // particles
  int n = 0; SphereStore sph[2]; int cur = 0; DBuf<int> id2slot, up2slot; int maxId = -1; bool up2slotValid = false;
  std::vector<Material> hMats; DBuf<Material> dMats;
  std::vector<Wall> hWalls; DBuf<Wall> dWalls; bool anyMovingWall = false, anyStaticWall = false;
  DBuf<double> wallAcc, wallPartial;     // accumulated force/torque on walls at rest (never reset in the reference: ContactLoop.cpp:270-276), per-step block partials
  std::vector<Facet> hFacets; DBuf<Facet> dFacets; DBuf<int> fgCellStart, fgCellFacet, fgCursor; FacetGrid fgrid; bool fgridValid = false;
  ClumpStore clumps;                     // clumps of spheres with mass (wdem_upload_clumps)
  // Impose (wdem_upload_imposes): descriptors and their host state, sets, node attachments, and the per-step ring of resolved impositions
  struct ImposeHost { int kind = 0, flags = 0; double v[16] = {0}; double ori[4] = {0, 0, 0, 1}; std::vector<double> times, vals, vals2; size_t ipos = 0, ipos2 = 0; double dist = 0; int nNodes = 0; };
  std::vector<ImposeHost> imposes; std::vector<int> impSetStart, impSetOps, wallImpSet, groupImpSet; std::vector<double> impWallVel;
  DBuf<int> dImpSetStart, dImpSetOps; DBuf<int32_t> dImpOfId; bool impSpheres = false, impRead = false, impFacetNodes = false; ImpSlot* hImpRing = nullptr; DBuf<ImpSlot> dImpRing; int lastImpSlot = 0;
  ImposeDev impDev(int slot) { ImposeDev I; I.slot = dImpRing.p ? dImpRing.p + slot : nullptr; I.setStart = dImpSetStart.p; I.setOps = dImpSetOps.p; return I; }
  GroupsU groups{};                      // kinematic facet groups (wdem_upload_facet_groups): master nodes integrated on the host, members on the device
  std::vector<double> groupRadius;       // largest distance of a member node from its master node (bounding sphere of the sweep of a rotating group)
  int nClumpedUploaded = 0;             // spheres uploaded with the CLUMPED flag that no clump has claimed yet
  double minRadius = INFINITY, pwaveDt = INFINITY;
  std::vector<double> hRadius; std::vector<int32_t> hMatOfSphere;
  std::vector<int32_t> hIds;                // Particle.id of the spheres in upload order
  std::vector<CRec> pendingRows;            // contact history imported by wdem_upload_contacts, consumed by the next collider full run
  // contacts
  ContactStore con[2]; int curC = 0; long long nSS = 0, nSSa = 0, nWS = 0, nFS = 0, nOrph = 0, nC = 0;   // table: [SS active (nSSa) | SS dormant | WS | FS | orphans]
  CsrStore csr[2]; int curR = 0;
  // scratch
  DBuf<uint32_t> keys, keysAlt; DBuf<int> idx, idxAlt, cellStart, counts, offsets; DBuf<unsigned char> cubTemp; DBuf<double> blk;
  DBuf<unsigned long long> ckeys, ckeysAlt, cntScratch; DBuf<int> orphCnt, rowCount, srcOld, candBuf; DBuf<uint32_t> wallBits; DBuf<double> ioBuf; DBuf<uint32_t> cvals, cvalsAlt; DBuf<int> realFlag, realPos; DBuf<int32_t> tmpCd[3];
  DBuf<double> facetBoxes;      // periodic cells: unsheared boxes of the facets at the last full run
  double* hBounds = nullptr; bool boundsValid = false; DBuf<double> boundsBuf;   // pinned, 2 x 8: bounding box of the sphere centres at the last two full runs
  Counters* dCnt = nullptr; Counters* hCnt = nullptr; int* hScratch = nullptr;   // pinned: per-step counters, small read-backs of the full run
  // scene scalars
  double dt = NAN, time = 0, nextDt = NAN, verletDist = NAN; long long step = 0, nFullRuns = 0; bool conDirty = true;
  long long dynStepLast = -1, dynStepPrev = -1;
  HostCell cell;
  long long launches = 0; int numSMs = 148;
  bool compactPending = false;   // spheres were removed since the last full run: the current generation is not the one the incidence rows were built for
  double gradVWork = 0, prevVol = NAN, prevStress[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // EnergyTracker "gradV" (ContactLoop.cpp:243-254)
  bool trackEnergy = false; DBuf<double> dEnergy;   // Scene.trackEnergy: EN_* accumulators on the device
  // end-to-end pipeline (wdem_step_e2e): host<->device copies of the state overlap with the integrator chunk by chunk
  static const int NCHUNK = 16; int nChunk = 8;   // NCHUNK = capacity, nChunk = pieces in use (WDEM_E2E_CHUNKS)
  struct E2E { bool active = false; bool defer = false; bool ioBusy = false; long long count = 0; const int* slotMap = nullptr; const double *pos = nullptr, *vel = nullptr, *angVel = nullptr; double *posOut = nullptr, *velOut = nullptr; } e2e;
  cudaStream_t sH2D = nullptr, sD2H = nullptr; cudaEvent_t evUp[NCHUNK] = {}, evLf[NCHUNK] = {}, evEk = nullptr;
  int lastLfChunk = 0; DBuf<int> ownedSlot; DBuf<int32_t> ownedIds;      // wdem_step_e2e_owned: slots of the ids the host sent, in its order
  DBuf<double> ioIn, ioOut; double* hEk = nullptr; size_t hEkCap = 0; int ekBlocks = 0;
  Spheres S() { Spheres v = sph[cur].view(); v.clumpOfId = clumps.n ? clumps.clumpOfId.p : nullptr; v.impOfId = impSpheres ? dImpOfId.p : nullptr; return v; }
};

namespace {

inline int nblk(long long n, int b) { return (int)((n + b - 1) / b); }

void validateConfig(const wdem_config& c) {
  if (c.law != WDEM_LAW_IDEALELPL && c.law != WDEM_LAW_HERTZ_DMT) throw EngErr{"unknown law", WDEM_ERR_INVALID};
  if (!c.iniLensTouch) throw EngErr{"Cg2_*.iniLensTouch=False is not supported by the GPU engine", WDEM_ERR_UNSUPPORTED};
  if (c.distFactor > 1.) throw EngErr{"DemField.distFactor>1 is not supported by the GPU engine", WDEM_ERR_UNSUPPORTED};
  if (c.dynDtPeriod > 0 && 1. + c.maxRelInc == 1.) throw EngErr{"DynDt: maxRelInc too small (1.0+maxRelInc==1.0)", WDEM_ERR_INVALID};  // DynDt.cpp:12-14
}

StepU makeUniforms(wdem_engine* E) {
  const wdem_config& c = E->cfg;
  StepU u; memset(&u, 0, sizeof(u));
  u.dt = E->dt; u.damping = c.damping; u.gravity = mk3(c.gravity[0], c.gravity[1], c.gravity[2]);
  u.hasGravity = (c.gravity[0] != 0 || c.gravity[1] != 0 || c.gravity[2] != 0); u.reset = c.reset;
  u.maxD2 = E->verletDist * E->verletDist;
  u.periodic = c.periodic; u.homoDeform = c.periodic ? E->cell.homoDeform : -1;
  u.hSize = E->cell.hSize;
  u.shiftVelM = mZero();
  if (c.periodic) {
    const HostCell& ce = E->cell;
    if (ce.homoDeform == WDEM_HOMO_VEL || ce.homoDeform == WDEM_HOMO_VEL_2ND) u.shiftVelM = mMul(ce.gradV, ce.hSize);
    else if (ce.homoDeform == WDEM_HOMO_GRADV2) u.shiftVelM = mMul(ce.gradV, ce.pprevHsize);
    u.gradV = ce.gradV; u.nextGradV = ce.nextGradV; u.dGradV = mSub(ce.nextGradV, ce.gradV); u.midGradV = mScale(mAdd(ce.gradV, ce.nextGradV), .5);
    u.spinVec = ce.spinVec;
    if (ce.homoDeform == WDEM_HOMO_GRADV2) {  // Leapfrog.cpp:114-127
      m33 s4 = mScale(mScale(mAdd(ce.nextGradV, ce.gradV), E->dt), .25);
      u.ImLL4hInv = mInv(mSub(mIdent(), s4)); u.IpLL4h = mAdd(mIdent(), s4); u.LmL = mSub(ce.nextGradV, ce.gradV);
      d3 a = hLevi(mScale(mSub(ce.gradV, mT(ce.gradV)), .5)), b = hLevi(mScale(mSub(ce.nextGradV, mT(ce.nextGradV)), .5));
      u.deltaSpinVec = (-.5) * a + .5 * b;
    }
  }
  u.gIn = E->dCnt->gflag; u.gOut = E->dCnt->gflag + 1;
  u.step = E->step; u.renormNow = (c.trsfRenorm > 0 && (E->step % c.trsfRenorm) == 0); u.noRatch = c.noRatch;
  u.absDistFactor = std::abs(c.distFactor);
  u.law = c.law; u.noSlip = c.noSlip; u.noBreak = c.noBreak; u.noFrict = c.noFrict; u.noAttraction = c.noAttraction;
  u.rolling = (c.relRollStiff > 0. && c.rollTanPhi > 0.); u.relRollStiff = c.relRollStiff; u.relTwistStiff = c.relTwistStiff; u.rollTanPhi = c.rollTanPhi;
  u.poisson = c.poisson; u.en = c.en;
  u.energy = E->trackEnergy ? E->dEnergy.p : nullptr;
  {   // Leapfrog.cpp:130-132
    bool ident = true; for (int k = 0; k < 9; k++) if (E->cell.gradV.m[k] != ((k % 4 == 0) ? 1. : 0.)) ident = false;
    u.trackLeapfrog = (E->trackEnergy && (!c.periodic || E->step > 0 || ident)) ? 1 : 0;
  }
  u.viscAlpha = (c.en > 0 && c.en < 1.) ? (-std::sqrt(5.) * std::log(c.en) / (std::sqrt(std::log(c.en) * std::log(c.en) + M_PI * M_PI))) : 0.;
  return u;
}

void checkDeviceError(wdem_engine* E) {
  const Counters& h = *E->hCnt;
  if (h.error == ERR_NONE) return;
  std::string pr = "##" + std::to_string(h.errA) + "+" + std::to_string(h.errB);
  if (E->comm) pr += " [subdomain " + std::to_string(E->comm->rank) + "; the ids are those of the failing subdomain]";
  switch (h.error) {
    case ERR_NAN_FORCE: throw EngErr{"NaN force in contact " + pr + " (IdealElPl.cpp:79-83)", WDEM_ERR_RUNTIME};
    case ERR_NAN_POS: throw EngErr{"NaN position of particle #" + std::to_string(h.errA), WDEM_ERR_RUNTIME};
    case ERR_CLUMP_INERTIA: throw EngErr{"Leapfrog::run: inertia==(0,0,0), but the node wants to use aspherical integrator (clump " + std::to_string(h.errA) + "; Leapfrog.cpp:225)", WDEM_ERR_RUNTIME};
    case ERR_NO_PHYS: throw EngErr{"ContactLoop: " + pr + ": no Contact.phys created (Cp2_HertzMat_HertzPhys needs HertzMat on both particles; ContactLoop.cpp:181)", WDEM_ERR_RUNTIME};
    case ERR_SCHWARZ: throw EngErr{"Law2_L6Geom_HertzPhys_DMT: Schwarz model (alpha!=0 and gamma!=0) is out of scope " + pr, WDEM_ERR_UNSUPPORTED};
    case ERR_PEER_TIMEOUT: throw EngErr{"slab decomposition: a neighbouring subdomain did not publish its step flags in time (it failed or stalled)", WDEM_ERR_RUNTIME};
    case ERR_AMBIGUOUS_PERI: throw EngErr{"InsertionSortCollider: ambiguous periodic contact " + pr + " (particle too large for the cell; InsertionSortCollider.cpp:979-981)", WDEM_ERR_RUNTIME};
    case ERR_FACET_IMPOSSIBLE: throw EngErr{"Cg2_Facet_Sphere_L6Geom: impossible sphere-facet intersection " + pr + " (Facet.cpp:358)", WDEM_ERR_RUNTIME};
    default: throw EngErr{"device error " + std::to_string(h.error) + " " + pr, WDEM_ERR_RUNTIME};
  }
}

void buildFacetGrid(wdem_engine* E);
void rebinFacets(wdem_engine* E);

// -------------------------------------------------------------------------------------------------
// Slab decomposition: host orchestration of migration, ghost export and the per-step ghost refresh.
// -------------------------------------------------------------------------------------------------
template <class T> T* as(DBuf<char>& b, size_t byteOff = 0) { return reinterpret_cast<T*>(b.p + byteOff); }

PeriU makePeriU(wdem_engine* E) {
  PeriU pu; memset(&pu, 0, sizeof(pu));
  if (E->cfg.periodic) { pu.unshear = E->cell.unshearTrsf; pu.size = E->cell.size; pu.cosv = E->cell.cosv; pu.hasShear = E->cell.hasShear ? 1 : 0; }
  return pu;
}

DomainU makeDomainU(wdem_engine* E) {
  DomainU D; memset(&D, 0, sizeof(D));
  const int W = E->comm->world;
  D.rank = E->comm->rank; D.world = W; D.axis = E->dom.axis; D.periodic = E->cfg.periodic; D.vd = E->verletDist; D.pu = makePeriU(E);
  D.dim = E->cfg.periodic ? comp(E->cell.size, D.axis) : 0.;
  D.bounds[0] = E->cfg.periodic ? 0. : -INFINITY; D.bounds[W] = E->cfg.periodic ? D.dim : INFINITY;
  for (int r = 1; r < W; r++) D.bounds[r] = E->cfg.periodic ? E->dom.cuts[r - 1] * D.dim : E->dom.cuts[r - 1];
  // largest half-extent of an enlarged box along the slab axis
  double h = E->dom.rmaxGlobal;
  if (E->cfg.periodic && E->cell.hasShear) {  // Cell::shearAlignedExtents (Cell.cpp:68-79) of the largest sphere
    const d3 c = E->cell.cosv; double hx = h, hy = h, hz = h;
    hy += .5 * h * (1 / c.x - 1); hz += .5 * h * (1 / c.x - 1); hz += .5 * h * (1 / c.y - 1); hx += .5 * h * (1 / c.y - 1); hx += .5 * h * (1 / c.z - 1); hy += .5 * h * (1 / c.z - 1);
    h = D.axis == 0 ? hx : (D.axis == 1 ? hy : hz);
  }
  D.reach = (h + E->verletDist) * (1 + 1e-9);
  return D;
}

// values every subdomain must agree on: smallest radius (verletDist), p-wave dt, largest radius (ghost reach)
void syncGlobals(wdem_engine* E) {
  if (!E->comm || E->dom.globalsValid) return;
  const int W = E->comm->world;
  double mine[3] = {E->minRadius, E->pwaveDt, E->dom.rmaxGlobal};
  int64_t m64[3]; memcpy(m64, mine, sizeof(mine));
  std::vector<int64_t> all((size_t)3 * W);
  E->comm->allgather64(m64, 3, all.data(), E->stream);
  for (int r = 0; r < W; r++) { double v[3]; memcpy(v, &all[(size_t)3 * r], sizeof(v)); E->minRadius = std::min(E->minRadius, v[0]); E->pwaveDt = std::min(E->pwaveDt, v[1]); E->dom.rmaxGlobal = std::max(E->dom.rmaxGlobal, v[2]); }
  E->dom.globalsValid = true;
}

void sortPairs32(wdem_engine* E, uint32_t* k, uint32_t* kAlt, int* v, int* vAlt, int n, int bits) {
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, k, kAlt, v, vAlt, n, 0, bits, E->stream);
  E->cubTemp.ensure(tb, E->bytes);
  cub::DeviceRadixSort::SortPairs(E->cubTemp.p, tb, k, kAlt, v, vAlt, n, 0, bits, E->stream);
}
void exclusiveSum(wdem_engine* E, const int* in, int* out, int n) {
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, E->stream);
  E->cubTemp.ensure(tb, E->bytes);
  cub::DeviceScan::ExclusiveSum(E->cubTemp.p, tb, in, out, n, E->stream);
}

// Runs at the start of every collider full run of a decomposed scene, on all subdomains in the same step:
//   1. owned spheres whose slab coordinate left the slab migrate (full state + the history rows of their real contacts);
//   2. ghosts of the previous generation are dropped; every owned sphere whose enlarged AABB can reach a box owned by rank r is exported to r.
// Leaves the new generation (owned first, then ghosts; unsorted) in the current sphere store.
static const size_t PEER_WORDS_BYTES = 2 * WDEM_MAX_RANKS * sizeof(unsigned long long);
PeerU makePeerU(wdem_engine* E);
// New cuts = the k/W quantiles of the slab coordinate of all owned spheres ("quantile cuts at rebuild time"), from a histogram that every rank
// builds of its own spheres over the global extent; all ranks add up the same table in the same order, so they arrive at the same cuts.  Collective.
void rebalanceCuts(wdem_engine* E) {
  Transport* T = E->comm; DomainState& M = E->dom; const int W = T->world; cudaStream_t st = E->stream;
  const DomainU D = makeDomainU(E);
  double lo = 0., hi = D.dim;
  if (!E->cfg.periodic) {      // global extent along the axis: the boxes measured at the last full run (a superset is fine; what lies outside falls into the end bins)
    const double* hb = E->hBounds + 8 * (int)((E->nFullRuns + 1) & 1);
    double mine[2] = {hb[M.axis], hb[3 + M.axis]}; int64_t mi[2], all[2 * WDEM_MAX_RANKS];
    if (!E->boundsValid) { mine[0] = INFINITY; mine[1] = -INFINITY; }      // no box measured here yet (e.g. nothing owned so far): an empty interval
    memcpy(mi, mine, sizeof(mine));
    T->allgather64(mi, 2, all, st);
    lo = INFINITY; hi = -INFINITY;
    for (int r = 0; r < W; r++) { double v[2]; memcpy(v, all + 2 * r, sizeof(v)); if (v[0] <= v[1]) { lo = std::min(lo, v[0]); hi = std::max(hi, v[1]); } }
    if (!(lo < hi) || !std::isfinite(lo) || !std::isfinite(hi)) return;
  }
  const double inv = WD_REBAL_BINS / (hi - lo);
  M.hist.ensure(WD_REBAL_BINS, E->bytes);
  CK(cudaMemsetAsync(M.hist.p, 0, WD_REBAL_BINS * sizeof(int), st));
  if (E->n > 0) { k_dom_coord_hist<<<nblk(E->n, 256), 256, 0, st>>>(E->S(), E->n, D, lo, inv, M.hist.p); E->launches++; }
  std::vector<int> h(WD_REBAL_BINS);
  CK(cudaMemcpyAsync(h.data(), M.hist.p, WD_REBAL_BINS * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::vector<int64_t> mine(WD_REBAL_BINS), all((size_t)WD_REBAL_BINS * W);
  for (int k = 0; k < WD_REBAL_BINS; k++) mine[k] = h[k];
  T->allgather64(mine.data(), WD_REBAL_BINS, all.data(), st);
  std::vector<long long> tot(WD_REBAL_BINS, 0); long long N = 0;
  for (int k = 0; k < WD_REBAL_BINS; k++) { for (int r = 0; r < W; r++) tot[k] += all[(size_t)r * WD_REBAL_BINS + k]; N += tot[k]; }
  if (N <= 0) return;
  std::vector<double> cuts(W - 1);
  long long acc = 0; int k = 0;
  for (int c = 1; c < W; c++) {
    const double target = (double)N * c / W;
    while (k < WD_REBAL_BINS - 1 && (double)(acc + tot[k]) < target) { acc += tot[k]; k++; }
    const double frac = tot[k] > 0 ? std::min(1., std::max(0., (target - (double)acc) / (double)tot[k])) : 0.;
    double x = lo + ((double)k + frac) / inv;
    if (c > 1) x = std::max(x, E->cfg.periodic ? cuts[c - 2] * D.dim : cuts[c - 2]);
    cuts[c - 1] = E->cfg.periodic ? x / D.dim : x;
  }
  M.cuts = cuts; M.nRebalances++; M.lastRebalRun = E->nFullRuns;
}
void domainExchange(wdem_engine* E) {
  static const bool tr = getenv("WDEM_REBUILD_TRACE") != nullptr;
  static double tacc[12] = {0}; static int tcnt = 0; std::chrono::steady_clock::time_point tp[12];
  auto TM = [&](int k) { if (tr) { cudaStreamSynchronize(E->stream); tp[k] = std::chrono::steady_clock::now(); } };
  Transport* T = E->comm; DomainState& M = E->dom;
  const int W = T->world, R = T->rank; cudaStream_t st = E->stream; int64_t& tot = E->bytes;
  const wdem_config& c = E->cfg;
  const bool hertz = c.law == WDEM_LAW_HERTZ_DMT, rolling = (c.relRollStiff > 0. && c.rollTanPhi > 0.);
  const int n = E->n;
  DomainU D = makeDomainU(E);
  SphereStore& Ast = E->sph[E->cur]; SphereStore& Bst = E->sph[1 - E->cur];
  Spheres A = Ast.view();
  ContactStore& Ost = E->con[E->curC]; CsrStore& Ocsr = E->csr[E->curR];
  OldCsr oc; memset(&oc, 0, sizeof(oc)); oc.valid = (Ocsr.valid && E->nC > 0) ? 1 : 0; oc.nRows = Ocsr.nRows; oc.rowStart = Ocsr.rowStart.p; oc.rowEntry = Ocsr.rowEntry.p; oc.rowOther = Ocsr.rowOther.p;
  TM(0);
  // ---- 0. nothing migrates anywhere (the usual case in a bed at rest): the owned spheres stay where they are, the ghosts of the previous generation are flagged
  //         dead in place and the new ones appended behind them — no regrouping sort, no copy of the sphere arrays, one host wait
  static const bool noFast = getenv("WDEM_NO_FAST_EXCHANGE") != nullptr;
  M.deadSlots = 0;
  if (!noFast) {
    M.emask.ensure((size_t)std::max(n, 1), tot); M.perRank.ensure((size_t)2 * WDEM_MAX_RANKS * (W + 1), tot);
    int* dMineF = M.perRank.p; int* dAllF = M.perRank.p + 2 * WDEM_MAX_RANKS;      // [emigrants + (no owned sphere), ghosts per destination...] : mine, then everybody's
    CK(cudaMemsetAsync(dMineF, 0, 2 * WDEM_MAX_RANKS * sizeof(int), st));
    if (n > 0) { k_dom_fast_scan<<<nblk(n, 256), 256, 0, st>>>(A, n, D, M.emask.p, dMineF); E->launches++; }
    if (!(M.nOwn > 0 && n == M.nOwn + (int)M.nRecvTot)) { k_fill_int<<<1, 1, 0, st>>>(dMineF + 1 + W, 1, 1); E->launches++; }      // my veto (e.g. nothing owned here)
    const bool rebalDue = M.rebalTol > 0. && (E->nFullRuns - M.lastRebalRun) >= M.rebalMinRuns;      // the same on every rank (no per-rank state in this condition: what follows is collective)
    if (rebalDue) { k_fill_int<<<1, 1, 0, st>>>(dMineF + 2 + W, 1, M.nOwn); E->launches++; }      // how many spheres I own: everybody sees the imbalance in the same table
    T->allgatherDev(dMineF, dAllF, (size_t)2 * WDEM_MAX_RANKS * sizeof(int), st);
    std::vector<int> fa((size_t)W * 2 * WDEM_MAX_RANKS, 0);
    CK(cudaMemcpyAsync(fa.data(), dAllF, fa.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // every rank reaches the same verdict: nobody emigrates anywhere, and every rank's arrays are [owned + ghosts of the last full run] with at least one owned sphere
    long long vetoAll = 0;
    for (int r = 0; r < W; r++) vetoAll += (long long)fa[(size_t)r * 2 * WDEM_MAX_RANKS] + fa[(size_t)r * 2 * WDEM_MAX_RANKS + 1 + W];
    if (rebalDue) {      // the same table on every rank => the same decision
      long long totOwn = 0, maxOwn = 0;
      for (int r = 0; r < W; r++) { const long long o = fa[(size_t)r * 2 * WDEM_MAX_RANKS + 2 + W]; totOwn += o; maxOwn = std::max(maxOwn, o); }
      if (totOwn > 0 && (double)maxOwn * W > (1. + M.rebalTol) * (double)totOwn) { rebalanceCuts(E); D = makeDomainU(E); vetoAll = 1; }
    }
    if (vetoAll == 0) {
      TM(1);
      const int nGhostOld = (int)M.nRecvTot;
      M.sendCnt.assign(W, 0); M.recvCnt.assign(W, 0);
      size_t accS = 0, accG = 0; GhostSeg seg; memset(&seg, 0, sizeof(seg));
      auto cnt = [&](int src, int dst) { return (size_t)fa[(size_t)src * 2 * WDEM_MAX_RANKS + 1 + dst]; };
      for (int r = 0; r < W; r++) { M.sendCnt[r] = cnt(R, r); M.recvCnt[r] = cnt(r, R); seg.base[r] = (int)accS; accS += M.sendCnt[r]; accG += M.recvCnt[r]; }
      std::vector<size_t> sOffG(W, 0), sBytG(W, 0), rOffG(W, 0), rBytG(W, 0);
      { size_t a = 0, b = 0; for (int r = 0; r < W; r++) { sOffG[r] = a * sizeof(GRec); sBytG[r] = M.sendCnt[r] * sizeof(GRec); a += M.sendCnt[r]; rOffG[r] = b * sizeof(GRec); rBytG[r] = M.recvCnt[r] * sizeof(GRec); b += M.recvCnt[r]; } }
      M.nSendTot = accS; M.nRecvTot = accG;
      M.peer.dstOff.assign(W, 0); M.peer.maxRecvTot = 0;
      for (int r = 0; r < W; r++) {
        long long off = 0, totR = 0;
        for (int q = 0; q < W; q++) { if (q < R) off += (long long)cnt(q, r); totR += (long long)cnt(q, r); }
        M.peer.dstOff[r] = off; M.peer.maxRecvTot = std::max(M.peer.maxRecvTot, totR);
      }
      M.sendBuf.ensure(accS * sizeof(GRec) + 64, tot); M.recvBuf.ensure(accG * sizeof(GRec) + 64, tot);
      M.exportId.ensure(std::max<size_t>(accS, 1), tot); M.importId.ensure(std::max<size_t>(accG, 1), tot);
      DomainState::Peer& P = M.peer;
      const bool viaPeer = P.active && 9 * P.maxRecvTot <= P.haloCap;
      const size_t stageOff = PEER_WORDS_BYTES + 2 * (size_t)P.haloCap * sizeof(double);
      GhostDst gd; memset(&gd, 0, sizeof(gd));
      for (int r = 0; r < W; r++) gd.out[r] = viaPeer ? reinterpret_cast<GRec*>(static_cast<char*>(P.base[r]) + stageOff) + P.dstOff[r] : as<GRec>(M.sendBuf) + seg.base[r];
      int* dCur = M.perRank.p;      // per-destination cursors
      CK(cudaMemsetAsync(dCur, 0, WDEM_MAX_RANKS * sizeof(int), st));
      if (accS > 0) { k_dom_pack_ghosts_all<<<nblk(n, 256), 256, 0, st>>>(A, n, M.emask.p, W, seg, dCur, gd, M.exportId.p); E->launches++; }
      const GRec* ghostsIn = as<GRec>(M.recvBuf);
      if (viaPeer) {
        M.syncFlag.ensure(4, tot); CK(cudaMemsetAsync(M.syncFlag.p, 0, 4 * sizeof(int), st));
        const PeerU u = makePeerU(E); const unsigned long long seq = ++P.syncSeq;
        k_peer_sync<<<1, WDEM_MAX_RANKS, 0, st>>>(u, E->dCnt, M.syncFlag.p, seq, (int)(seq & 1), 20000000000ll); E->launches++;
        ghostsIn = reinterpret_cast<const GRec*>(P.box + stageOff);
      } else T->alltoallv(M.sendBuf.p, sOffG.data(), sBytG.data(), M.recvBuf.p, rOffG.data(), rBytG.data(), st);
      const int nNew = n + (int)accG;
      Ast.ensure((size_t)std::max(nNew + 1024, 1), tot, /*keep*/ true, st); A = Ast.view();
      if (n > 0) { k_dom_mark_dead<<<nblk(n, 256), 256, 0, st>>>(A.flags, n); E->launches++; }
      if (accG > 0) { k_dom_unpack_ghosts<<<nblk(accG, 256), 256, 0, st>>>(A, n, (int)accG, ghostsIn, M.importId.p); E->launches++; }
      M.oldIds.ensure((size_t)std::max(n, 1), tot); M.origSlot.ensure((size_t)std::max(nNew, 1), tot);
      if (n > 0) CK(cudaMemcpyAsync(M.oldIds.p, A.id, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
      if (nNew > 0) { k_dom_orig_slot<<<nblk(nNew, 256), 256, 0, st>>>(A.id, nNew, E->id2slot.p, M.origSlot.p); E->launches++; }
      if (n > 0) { k_dom_clear_id2slot<<<nblk(n, 256), 256, 0, st>>>(A.id, n, E->id2slot.p); E->launches++; }
      E->n = nNew; M.deadSlots = nGhostOld; M.nImm = 0; E->up2slotValid = false;
      if (tr) { TM(2); tacc[8] += std::chrono::duration<double, std::milli>(tp[1] - tp[0]).count(); tacc[9] += std::chrono::duration<double, std::milli>(tp[2] - tp[1]).count();
        static int fcnt = 0; if (++fcnt % 25 == 0) { fprintf(stderr, "domainExchange trace (rank %d, last 25 of %d in-place calls): scan+counts %.3f ms, ghosts+bookkeeping %.3f ms\n", R, fcnt, tacc[8] / 25, tacc[9] / 25); tacc[8] = tacc[9] = 0; } }
      return;
    }
  }
  // ---- 1. classify and group by destination; 2. what every rank sends to every other: counted on the device, exchanged from device memory (one host wait)
  std::vector<int> start(W + 3, 0);
  const size_t nn = std::max(n, 1);
  M.dkeys.ensure(nn, tot); M.dkeysAlt.ensure(nn, tot); M.dslots.ensure(nn, tot); M.dslotsAlt.ensure(nn, tot); M.dstart.ensure(W + 3, tot);
  M.rowCnt.ensure(nn + 1, tot); M.rowOff.ensure(nn + 1, tot); M.perRank.ensure((size_t)2 * WDEM_MAX_RANKS * (W + 1), tot);
  int* dMine = M.perRank.p; int* dAll = M.perRank.p + 2 * WDEM_MAX_RANKS;      // (spheres, rows) per destination: mine, then everybody's
  std::vector<int> all((size_t)2 * W * W, 0);
  if (n > 0) {
    k_dom_classify<<<nblk(n, 256), 256, 0, st>>>(A, n, D, M.dkeys.p, M.dslots.p);
    int bits = 1; while ((1 << bits) < W + 2) bits++;
    sortPairs32(E, M.dkeys.p, M.dkeysAlt.p, M.dslots.p, M.dslotsAlt.p, n, bits);
    k_fill_int<<<1, 64, 0, st>>>(M.dstart.p, W + 3, n);
    k_dom_starts<<<nblk(n, 256), 256, 0, st>>>(M.dkeysAlt.p, n, M.dstart.p);
    k_dom_count_rows_all<<<nblk(n + 1, 128), 128, 0, st>>>(M.dkeysAlt.p, M.dslotsAlt.p, n, W, oc, Ost.flags.p, M.rowCnt.p);
    exclusiveSum(E, M.rowCnt.p, M.rowOff.p, n + 1);
    k_dom_send_counts<<<1, 32, 0, st>>>(M.dstart.p, n, W, M.rowOff.p, dMine);
    E->launches += 11;
  } else { CK(cudaMemsetAsync(dMine, 0, 2 * W * sizeof(int), st)); CK(cudaMemsetAsync(M.dstart.p, 0, (W + 3) * sizeof(int), st)); }
  T->allgatherDev(dMine, dAll, (size_t)2 * W * sizeof(int), st);
  int* hTab = E->hScratch + 16;      // pinned: [W+3 starts | W*2W table]  (WDEM_MAX_RANKS = 16: 19 + 512 ints < 256?  no -> table goes through a vector below)
  CK(cudaMemcpyAsync(hTab, M.dstart.p, (W + 3) * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(all.data(), dAll, all.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int k = 0; k < W + 3; k++) start[k] = hTab[k];
  if (n == 0) for (int k = 0; k < W + 3; k++) start[k] = 0;
  const int nKeep = start[1] - start[0], nE = start[W + 1] - start[1];
  TM(1);
  std::vector<size_t> sOffP(W, 0), sBytP(W, 0), rOffP(W, 0), rBytP(W, 0), sOffR(W, 0), sBytR(W, 0), rOffR(W, 0), rBytR(W, 0);
  size_t nImmP = 0, nImmR = 0, accP = 0, accR = 0, accSR = 0;
  for (int r = 0; r < W; r++) {
    const size_t mp = (size_t)all[(size_t)R * 2 * W + 2 * r], mr = (size_t)all[(size_t)R * 2 * W + 2 * r + 1];
    sOffP[r] = (size_t)(start[1 + r] - start[1]) * sizeof(PRec); sBytP[r] = mp * sizeof(PRec);
    sOffR[r] = accSR * sizeof(CRec); sBytR[r] = mr * sizeof(CRec); accSR += mr;
    const size_t np = (size_t)all[(size_t)r * 2 * W + 2 * R], nr = (size_t)all[(size_t)r * 2 * W + 2 * R + 1];
    rOffP[r] = accP * sizeof(PRec); rBytP[r] = np * sizeof(PRec); accP += np;
    rOffR[r] = accR * sizeof(CRec); rBytR[r] = nr * sizeof(CRec); accR += nr;
  }
  nImmP = accP; nImmR = accR;
  const size_t sendRowsBase = (size_t)nE * sizeof(PRec), recvRowsBase = nImmP * sizeof(PRec);
  M.sendBuf.ensure(sendRowsBase + accSR * sizeof(CRec) + 64, tot); M.recvBuf.ensure(recvRowsBase + nImmR * sizeof(CRec) + 64, tot);
  if (nE > 0) {
    k_dom_pack_particles<<<nblk(nE, 256), 256, 0, st>>>(A, M.dslotsAlt.p, start[1], nE, as<PRec>(M.sendBuf));
    if (accSR > 0)      // rowOff is an exclusive scan over ALL spheres in destination order; the kept ones (key 0, in front) have no rows, so the emigrants' offsets start at 0
      k_dom_pack_rows<<<nblk(nE, 128), 128, 0, st>>>(M.dslotsAlt.p, start[1], nE, oc, Ost.view(), A.id, E->dWalls.p, E->dFacets.p, M.rowOff.p + start[1], c.law, rolling, c.periodic, as<CRec>(M.sendBuf, sendRowsBase));
    E->launches += 2;
  }
  T->alltoallv(M.sendBuf.p, sOffP.data(), sBytP.data(), M.recvBuf.p, rOffP.data(), rBytP.data(), st);
  T->alltoallv(M.sendBuf.p + sendRowsBase, sOffR.data(), sBytR.data(), M.recvBuf.p + recvRowsBase, rOffR.data(), rBytR.data(), st);
  TM(2);
  // ---- 3. new owned set = kept + immigrants
  const int nOwn = nKeep + (int)nImmP;
  Bst.ensure((size_t)std::max(nOwn + (int)M.nRecvTot + 1024, 1), tot);
  Spheres B = Bst.view();
  if (nKeep > 0) { k_dom_gather_kept<<<nblk(nKeep, 256), 256, 0, st>>>(B, A, M.dslotsAlt.p + start[0], nKeep); E->launches++; }
  if (nImmP > 0) { k_dom_unpack_particles<<<nblk(nImmP, 256), 256, 0, st>>>(B, nKeep, (int)nImmP, as<PRec>(M.recvBuf)); E->launches++; }
  // ---- 4. immigrant history rows: appended behind the previous table, keyed by (idA,idB)
  M.nImm = (int)nImmR;
  if (nImmR > 0) {
    Ost.ensure((size_t)(E->nC + nImmR), tot, hertz, rolling, c.periodic, /*keep*/ true, st);
    M.immKeys.ensure(nImmR, tot); M.immKeysAlt.ensure(nImmR, tot); M.immVals.ensure(nImmR, tot); M.immValsAlt.ensure(nImmR, tot);
    k_dom_unpack_rows<<<nblk(nImmR, 256), 256, 0, st>>>(Ost.view(), E->nC, (int)nImmR, as<CRec>(M.recvBuf, recvRowsBase), c.law, rolling, c.periodic, M.immKeys.p, M.immVals.p, 0);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, M.immKeys.p, M.immKeysAlt.p, M.immVals.p, M.immValsAlt.p, (int)nImmR, 0, 64, st);
    E->cubTemp.ensure(tb, tot);
    cub::DeviceRadixSort::SortPairs(E->cubTemp.p, tb, M.immKeys.p, M.immKeysAlt.p, M.immVals.p, M.immValsAlt.p, (int)nImmR, 0, 64, st);
    E->launches += 4;
  }
  TM(3);
  // ---- 5. ghost export: per-destination counts on the device, exchanged from device memory (the second and last host wait), one packing pass for all destinations
  M.emask.ensure((size_t)std::max(nOwn, 1), tot);
  int* dPer = M.perRank.p; int* dPerAll = M.perRank.p + WDEM_MAX_RANKS;
  CK(cudaMemsetAsync(dPer, 0, WDEM_MAX_RANKS * sizeof(int), st));
  if (nOwn > 0) { k_dom_export_mask<<<nblk(nOwn, 256), 256, 0, st>>>(B, nOwn, D, M.emask.p, dPer); E->launches++; }
  T->allgatherDev(dPer, dPerAll, WDEM_MAX_RANKS * sizeof(int), st);
  std::vector<int> ga((size_t)W * WDEM_MAX_RANKS, 0);
  CK(cudaMemcpyAsync(ga.data(), dPerAll, ga.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  TM(4);
  M.sendCnt.assign(W, 0); M.recvCnt.assign(W, 0);
  std::vector<size_t> sOffG(W, 0), sBytG(W, 0), rOffG(W, 0), rBytG(W, 0);
  size_t accS = 0, accG = 0;
  GhostSeg seg; memset(&seg, 0, sizeof(seg));
  for (int r = 0; r < W; r++) {
    M.sendCnt[r] = (size_t)ga[(size_t)R * WDEM_MAX_RANKS + r]; M.recvCnt[r] = (size_t)ga[(size_t)r * WDEM_MAX_RANKS + R];
    seg.base[r] = (int)accS;
    sOffG[r] = accS * sizeof(GRec); sBytG[r] = M.sendCnt[r] * sizeof(GRec); accS += M.sendCnt[r];
    rOffG[r] = accG * sizeof(GRec); rBytG[r] = M.recvCnt[r] * sizeof(GRec); accG += M.recvCnt[r];
  }
  M.nSendTot = accS; M.nRecvTot = accG;
  M.peer.dstOff.assign(W, 0); M.peer.maxRecvTot = 0;
  for (int r = 0; r < W; r++) {   // where my segment starts in rank r's receive layout (segments ordered by source rank), and the largest halo of any rank
    long long off = 0, totR = 0;
    for (int q = 0; q < W; q++) { if (q < R) off += ga[(size_t)q * WDEM_MAX_RANKS + r]; totR += ga[(size_t)q * WDEM_MAX_RANKS + r]; }
    M.peer.dstOff[r] = off; M.peer.maxRecvTot = std::max(M.peer.maxRecvTot, totR);
  }
  M.sendBuf.ensure(accS * sizeof(GRec) + 64, tot); M.recvBuf.ensure(accG * sizeof(GRec) + 64, tot);
  M.exportId.ensure(std::max<size_t>(accS, 1), tot); M.importId.ensure(std::max<size_t>(accG, 1), tot);
  if (accS > 0) CK(cudaMemsetAsync(dPer, 0, WDEM_MAX_RANKS * sizeof(int), st));      // reused as the per-destination cursors
  // ghost records travel like the per-step halo when the mailboxes are up and large enough (the same decision on every rank: maxRecvTot is a global maximum; it is
  // also peerSetup's criterion, so the mailboxes used here are not re-created by this full run): stored straight into the receiver's staging area over NVLink, then one
  // release/acquire word per peer.  Otherwise (first full run, NCCL-only transports) ncclSend/ncclRecv — measured 0.4-2.5 ms per exchange for < 1 MB on the B200 boxes.
  DomainState::Peer& P = M.peer;
  const bool viaPeer = P.active && 9 * P.maxRecvTot <= P.haloCap;
  const size_t stageOff = PEER_WORDS_BYTES + 2 * (size_t)P.haloCap * sizeof(double);
  GhostDst gd; memset(&gd, 0, sizeof(gd));
  for (int r = 0; r < W; r++) gd.out[r] = viaPeer ? reinterpret_cast<GRec*>(static_cast<char*>(P.base[r]) + stageOff) + P.dstOff[r] : as<GRec>(M.sendBuf) + seg.base[r];
  if (accS > 0) { k_dom_pack_ghosts_all<<<nblk(nOwn, 256), 256, 0, st>>>(B, nOwn, M.emask.p, W, seg, dPer, gd, M.exportId.p); E->launches++; }
  TM(5);
  const GRec* ghostsIn = as<GRec>(M.recvBuf);
  if (viaPeer) {
    M.syncFlag.ensure(4, tot); CK(cudaMemsetAsync(M.syncFlag.p, 0, 4 * sizeof(int), st));
    const PeerU u = makePeerU(E); const unsigned long long seq = ++P.syncSeq;
    k_peer_sync<<<1, WDEM_MAX_RANKS, 0, st>>>(u, E->dCnt, M.syncFlag.p, seq, (int)(seq & 1), 20000000000ll); E->launches++;
    ghostsIn = reinterpret_cast<const GRec*>(P.box + stageOff);
  } else T->alltoallv(M.sendBuf.p, sOffG.data(), sBytG.data(), M.recvBuf.p, rOffG.data(), rBytG.data(), st);
  TM(6);
  const int nNew = nOwn + (int)accG;
  Bst.ensure((size_t)std::max(nNew, 1), tot, /*keep*/ true, st); B = Bst.view();
  if (accG > 0) { k_dom_unpack_ghosts<<<nblk(accG, 256), 256, 0, st>>>(B, nOwn, (int)accG, ghostsIn, M.importId.p); E->launches++; }
  TM(7);
  // ---- 6. bookkeeping for the history lookup of the full run that follows
  M.oldIds.ensure(nn, tot); M.origSlot.ensure((size_t)std::max(nNew, 1), tot);
  if (n > 0) CK(cudaMemcpyAsync(M.oldIds.p, A.id, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  if (nNew > 0) { k_dom_orig_slot<<<nblk(nNew, 256), 256, 0, st>>>(B.id, nNew, E->id2slot.p, M.origSlot.p); E->launches++; }
  if (n > 0) { k_dom_clear_id2slot<<<nblk(n, 256), 256, 0, st>>>(A.id, n, E->id2slot.p); E->launches++; }
  E->n = nNew; E->cur = 1 - E->cur; M.nOwn = nOwn; E->up2slotValid = false;
  M.migratedOut += nE; M.migratedIn += (long long)nImmP; M.rowsIn += (long long)nImmR;
  if (tr) {
    TM(8); for (int k = 0; k < 8; k++) tacc[k] += std::chrono::duration<double, std::milli>(tp[k + 1] - tp[k]).count();
    if (++tcnt % 50 == 0) { fprintf(stderr, "domainExchange trace (rank %d, mean of %d, ms): classify+counts %.3f | pack+a2a(migrants) %.3f | kept+unpack %.3f | mask+counts %.3f | ghost pack %.3f | ghost a2a %.3f | ghost unpack %.3f | bookkeeping %.3f | nE %d nImm %zu nGhostOut %zu\n",
                                    R, tcnt, tacc[0] / tcnt, tacc[1] / tcnt, tacc[2] / tcnt, tacc[3] / tcnt, tacc[4] / tcnt, tacc[5] / tcnt, tacc[6] / tcnt, tacc[7] / tcnt, nE, nImmP, accS); }
  }
}

void peerTeardown(wdem_engine* E, bool collective) {
  DomainState::Peer& P = E->dom.peer; Transport* T = E->comm;
  if (!P.box) return;
  cudaStreamSynchronize(E->stream);
  for (int r = 0; r < WDEM_MAX_RANKS; r++) if (P.base[r] && T && r != T->rank) { cudaIpcCloseMemHandle(P.base[r]); P.base[r] = nullptr; }
  if (collective && T) { int64_t z = 0; std::vector<int64_t> all(T->world); T->allgather64(&z, 1, all.data(), E->stream); }   // nobody maps my box any more
  cudaFree(P.box); P.box = nullptr; P.active = false;
}
// (re)create the mailboxes; collective (called by all ranks from the same collider run).  Any failure disables the peer path on ALL ranks (NCCL stays).
void peerSetup(wdem_engine* E) {
  DomainState::Peer& P = E->dom.peer; Transport* T = E->comm; const int W = T->world, R = T->rank; cudaStream_t st = E->stream;
  static const bool off = getenv("WDEM_NO_PEER") != nullptr;
  if (off || P.disabled || !T->multiProcess() || W < 2) return;
  const long long need = 9 * P.maxRecvTot;
  if (P.active && need <= P.haloCap) return;      // same decision on every rank: maxRecvTot is a global maximum
  peerTeardown(E, true);
  P.haloCap = std::max<long long>(2 * need, 1 << 17);
  const size_t bytes = PEER_WORDS_BYTES + 2 * (size_t)P.haloCap * sizeof(double) + (size_t)(P.haloCap / 9 + 1) * sizeof(GRec);      // words | halo x2 | ghost staging area (full runs)
  int64_t ok = 1; cudaIpcMemHandle_t hd; memset(&hd, 0, sizeof(hd));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  if (cudaMalloc(&P.box, bytes) != cudaSuccess) { P.box = nullptr; ok = 0; cudaGetLastError(); }
  if (ok) { CK(cudaMemsetAsync(P.box, 0, bytes, st)); CK(cudaStreamSynchronize(st)); if (cudaIpcGetMemHandle(&hd, P.box) != cudaSuccess) { ok = 0; cudaGetLastError(); } }
  std::vector<int64_t> mine(9, 0), all((size_t)9 * W, 0);
  memcpy(mine.data(), &hd, 64); mine[8] = ok;
  T->allgather64(mine.data(), 9, all.data(), st);
  for (int r = 0; r < W; r++) if (!all[(size_t)9 * r + 8]) ok = 0;
  if (ok) {
    for (int r = 0; r < W && ok; r++) {
      if (r == R) { P.base[r] = P.box; continue; }
      cudaIpcMemHandle_t hr; memcpy(&hr, &all[(size_t)9 * r], 64);
      if (cudaIpcOpenMemHandle(&P.base[r], hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { P.base[r] = nullptr; ok = 0; cudaGetLastError(); }
    }
  }
  int64_t ok2 = ok; std::vector<int64_t> all2(W, 0);
  T->allgather64(&ok2, 1, all2.data(), st);
  for (int r = 0; r < W; r++) if (!all2[r]) ok = 0;
  if (!ok) {   // some rank could not allocate / export / map: everybody falls back to NCCL for good
    for (int r = 0; r < W; r++) if (P.base[r] && r != R) { cudaIpcCloseMemHandle(P.base[r]); } 
    for (auto& b : P.base) b = nullptr;
    std::vector<int64_t> all3(W, 0); int64_t z = 0; T->allgather64(&z, 1, all3.data(), st);
    if (P.box) { cudaFree(P.box); P.box = nullptr; }
    P.disabled = true; P.active = false; return;
  }
  P.active = true;
}
PeerU makePeerU(wdem_engine* E) {
  DomainState& M = E->dom; Transport* T = E->comm; PeerU u; memset(&u, 0, sizeof(u));
  u.rank = T->rank; u.world = T->world; u.haloCap = M.peer.haloCap;
  int acc = 0;
  for (int r = 0; r < T->world; r++) {
    u.words[r] = reinterpret_cast<unsigned long long*>(M.peer.base[r]);
    u.halo[r] = reinterpret_cast<double*>(static_cast<char*>(M.peer.base[r]) + PEER_WORDS_BYTES);
    u.sendBase[r] = acc; acc += (int)M.sendCnt[r]; u.dstOff[r] = M.peer.dstOff[r];
  }
  for (int r = T->world; r <= WDEM_MAX_RANKS; r++) u.sendBase[r] = acc;
  return u;
}
// per step, after Leapfrog: push my exported spheres into the neighbours' halo buffers, publish (stamp|error|dirty), wait for everybody's word
// The exchanges are numbered (syncSeq), not the steps: steps of a batch that turned out void are enqueued again after the full run, and a stamp must never repeat.
// Every rank enqueues the same sequence of exchanges (same batch sizes, same global flags), so the numbers agree.
void peerExchange(wdem_engine* E, int* flag) {
  DomainState& M = E->dom; cudaStream_t st = E->stream;
  const PeerU u = makePeerU(E); const unsigned long long seq = ++M.peer.syncSeq; const int parity = (int)(seq & 1);
  if (M.nSendTot && !E->conDirty) { k_halo_push<<<nblk(M.nSendTot, 256), 256, 0, st>>>(E->S(), M.sendSlot.p, (int)M.nSendTot, u, parity); E->launches++; }
  k_peer_sync<<<1, WDEM_MAX_RANKS, 0, st>>>(u, E->dCnt, flag, seq, parity, 20000000000ll); E->launches++;
}
void peerUnpack(wdem_engine* E) {
  DomainState& M = E->dom; const int parity = (int)(M.peer.syncSeq & 1);
  const double* buf = reinterpret_cast<const double*>(M.peer.box + PEER_WORDS_BYTES) + (long long)parity * M.peer.haloCap;
  if (M.nRecvTot) { k_halo_unpack<<<nblk(M.nRecvTot, 256), 256, 0, E->stream>>>(E->S(), M.recvSlot.p, (int)M.nRecvTot, buf); E->launches++; }
}

// after the full run: exported / imported ids -> slots of the sorted generation
void domainFinishRebuild(wdem_engine* E) {
  DomainState& M = E->dom; cudaStream_t st = E->stream;
  M.sendSlot.ensure(std::max<size_t>(M.nSendTot, 1), E->bytes); M.recvSlot.ensure(std::max<size_t>(M.nRecvTot, 1), E->bytes);
  if (M.nSendTot) { k_dom_ids_to_slots<<<nblk(M.nSendTot, 256), 256, 0, st>>>(M.exportId.p, (int)M.nSendTot, E->id2slot.p, M.sendSlot.p); E->launches++; }
  if (M.nRecvTot) { k_dom_ids_to_slots<<<nblk(M.nRecvTot, 256), 256, 0, st>>>(M.importId.p, (int)M.nRecvTot, E->id2slot.p, M.recvSlot.p); E->launches++; }
  M.haloSend.ensure(std::max<size_t>(M.nSendTot * 9, 1), E->bytes); M.haloRecv.ensure(std::max<size_t>(M.nRecvTot * 9, 1), E->bytes);
  peerSetup(E);
}

// per step (no full run): owners send pos, vel, angVel of the spheres they exported; 72 B per ghost
void haloExchange(wdem_engine* E) {
  DomainState& M = E->dom; Transport* T = E->comm; cudaStream_t st = E->stream; const int W = T->world;
  if (M.nSendTot) { k_halo_pack<<<nblk(M.nSendTot, 256), 256, 0, st>>>(E->S(), M.sendSlot.p, (int)M.nSendTot, M.haloSend.p); E->launches++; }
  std::vector<size_t> sOff(W, 0), sByt(W, 0), rOff(W, 0), rByt(W, 0);
  size_t a = 0, b = 0;
  for (int r = 0; r < W; r++) { sOff[r] = a * 72; sByt[r] = M.sendCnt[r] * 72; a += M.sendCnt[r]; rOff[r] = b * 72; rByt[r] = M.recvCnt[r] * 72; b += M.recvCnt[r]; }
  T->alltoallv((const char*)M.haloSend.p, sOff.data(), sByt.data(), (char*)M.haloRecv.p, rOff.data(), rByt.data(), st);
  if (M.nRecvTot) { k_halo_unpack<<<nblk(M.nRecvTot, 256), 256, 0, st>>>(E->S(), M.recvSlot.p, (int)M.nRecvTot, M.haloRecv.p); E->launches++; }
}

// -------------------------------------------------------------------------------------------------
// Collider full run (replaces InsertionSortCollider::run's fullRun branch, InsertionSortCollider.cpp:407-613)
// -------------------------------------------------------------------------------------------------
void rebuild(wdem_engine* E) {
  const wdem_config& c = E->cfg;
  if (!E->comm) E->dom.deadSlots = 0;
  cudaStream_t st = E->stream;
  static const bool traceReb = getenv("WDEM_REBUILD_TRACE") != nullptr;
  auto now = [&] { if (traceReb) cudaStreamSynchronize(E->stream); return std::chrono::steady_clock::now(); };
  auto t0 = now();
  if (E->comm) domainExchange(E);   // migration + ghosts; leaves the new (unsorted) generation in the current store
  auto t1 = now();
  int n = E->n;
  const bool hertz = c.law == WDEM_LAW_HERTZ_DMT, rolling = (c.relRollStiff > 0. && c.rollTanPhi > 0.);
  const double vd = E->verletDist;
  int64_t& tot = E->bytes;
  long long nSS = 0, nSSa = 0, nWS = 0, nFS = 0, nOrphOut = 0;   // nSSa: active sphere-sphere contacts (the first nSSa of the nSS); the rest is dormant
  if (!E->fgridValid) buildFacetGrid(E);
  if ((E->groups.n > 0 || E->impFacetNodes) && !E->hFacets.empty()) rebinFacets(E);
  const int nFacets = (int)E->hFacets.size();
  ContactStore& Nst = E->con[1 - E->curC];
  ContactStore& Ost = E->con[E->curC];
  CsrStore& Ocsr = E->csr[E->curR];
  CsrStore& Ncsr = E->csr[1 - E->curR];
  OldCsr oc; memset(&oc, 0, sizeof(oc)); oc.valid = Ocsr.valid ? 1 : 0; oc.nRows = Ocsr.nRows; oc.rowStart = Ocsr.rowStart.p; oc.rowEntry = Ocsr.rowEntry.p; oc.rowOther = Ocsr.rowOther.p;
  if (E->comm && E->dom.nImm > 0) { oc.immKeys = E->dom.immKeysAlt.p; oc.immVals = E->dom.immValsAlt.p; oc.nImm = E->dom.nImm; }
  bool imported = false;
  if (!E->pendingRows.empty()) {
    // wdem_upload_contacts: append the imported history rows behind the previous table and key them by (idA,idB), like rows that migrate in
    DomainState& M = E->dom; const size_t nR = E->pendingRows.size();
    Ost.ensure((size_t)(E->nC + nR), tot, hertz, rolling, c.periodic, /*keep*/ true, st);
    M.recvBuf.ensure(nR * sizeof(CRec) + 64, tot);
    CK(cudaMemcpyAsync(M.recvBuf.p, E->pendingRows.data(), nR * sizeof(CRec), cudaMemcpyHostToDevice, st));
    M.immKeys.ensure(nR, tot); M.immKeysAlt.ensure(nR, tot); M.immVals.ensure(nR, tot); M.immValsAlt.ensure(nR, tot);
    k_dom_unpack_rows<<<nblk(nR, 256), 256, 0, st>>>(Ost.view(), E->nC, (int)nR, as<CRec>(M.recvBuf), c.law, rolling, c.periodic, M.immKeys.p, M.immVals.p, 1);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, M.immKeys.p, M.immKeysAlt.p, M.immVals.p, M.immValsAlt.p, (int)nR, 0, 64, st);
    E->cubTemp.ensure(tb, tot);
    cub::DeviceRadixSort::SortPairs(E->cubTemp.p, tb, M.immKeys.p, M.immKeysAlt.p, M.immVals.p, M.immValsAlt.p, (int)nR, 0, 64, st);
    CK(cudaStreamSynchronize(st));   // pendingRows is pageable host memory
    oc.immKeys = M.immKeysAlt.p; oc.immVals = M.immValsAlt.p; oc.nImm = (int)nR; oc.immAll = 1;
    E->nC += (long long)nR; E->pendingRows.clear(); E->pendingRows.shrink_to_fit(); imported = true; E->launches += 4;
  }
  const int* perm = nullptr;
  HGrid g; memset(&g, 0, sizeof(g));
  PeriU pu; memset(&pu, 0, sizeof(pu));
  if (c.periodic) {
    if (nFacets > WD_MAX_PERI_FACETS) throw EngErr{"more than " + std::to_string(WD_MAX_PERI_FACETS) + " facets in a periodic cell are not supported by the GPU engine (every sphere visits every facet there)", WDEM_ERR_UNSUPPORTED};
    if (!E->hWalls.empty() && E->cell.hasShear) throw EngErr{"Walls not supported in skewed periodic boundary conditions.", WDEM_ERR_RUNTIME};  // Wall.cpp:30
    pu.unshear = E->cell.unshearTrsf; pu.size = E->cell.size; pu.cosv = E->cell.cosv; pu.hasShear = E->cell.hasShear ? 1 : 0;
  }
  if (n > 0) {
    int nb = nblk(n, 256);
    long long ncells = 0;
    const int nDead = E->dom.deadSlots; E->dom.deadSlots = 0;      // slots flagged F_DEADSLOT: they get the key behind the last cell and are cut off after the sort
    E->keys.ensure(n, tot); E->keysAlt.ensure(n, tot); E->idx.ensure(n, tot); E->idxAlt.ensure(n, tot);
    if (!c.periodic) {
      // 1. bounds of sphere centres
      // The grid only has to COVER the spheres for speed, not for correctness (cell coordinates are clamped on both sides of every lookup), so the box measured at
      // the previous full run (read back without waiting, pinned memory) lays out this run's grid; only the first run after an upload waits for its own.
      E->blk.ensure((size_t)nb * 7, tot);
      if (!E->boundsBuf.p) { E->boundsBuf.ensure(16, tot); CK(cudaMemsetAsync(E->boundsBuf.p, 0, 16 * sizeof(double), st)); }      // 7 results + the ticket word (the kernel leaves it at 0)
      double* dB = E->boundsBuf.p; unsigned* tick = reinterpret_cast<unsigned*>(dB + 8);
      k_bounds<<<nb, 256, 0, st>>>(E->S().posr, n, E->blk.p, dB, tick); E->launches++;
      const int slot = (int)(E->nFullRuns & 1);
      CK(cudaMemcpyAsync(E->hBounds + 8 * slot, dB, 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
      const double* hb = E->hBounds + 8 * (1 - slot);
      if (!E->boundsValid) { CK(cudaStreamSynchronize(st)); hb = E->hBounds + 8 * slot; }
      E->boundsValid = true;
      double lo[3], hi[3];
      for (int k = 0; k < 3; k++) { lo[k] = hb[k] - 2 * vd; hi[k] = hb[3 + k] + 2 * vd; }
      for (int k = 0; k < 3; k++) if (!std::isfinite(lo[k]) || !std::isfinite(hi[k])) throw EngErr{"non-finite particle position at collider run", WDEM_ERR_RUNTIME};
      // 2. size-class grids (HGrid): level L = radii [rmin 2^L, rmin 2^(L+1)), cell edge = largest enlarged box of the level, at most ~8 cells per sphere
      const double rmin = E->minRadius, rmax = std::max(E->dom.rmaxGlobal, rmin);
      int nL = 1; while (nL < WD_MAX_LEVELS && rmin * (double)(1 << nL) <= rmax) nL++;
      g.nL = nL; g.periodic = 0; g.rmin = rmin; g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
      double ext[3]; for (int k = 0; k < 3; k++) ext[k] = hi[k] - lo[k];
      for (int L = 0; L < nL; L++) {
        const double rL = (L == nL - 1) ? rmax : std::min(rmax, rmin * (double)(2 << L));
        const double reach = rL + vd;
        double cell = 2 * reach * (1 + 1e-9) + 1e-300;
        const double vol = std::max(ext[0], cell) * std::max(ext[1], cell) * std::max(ext[2], cell);
        cell = std::max(cell, std::cbrt(vol / (8.0 * n)));
        g.base[L] = (int)ncells;
        for (int k = 0; k < 3; k++) { g.inv[L][k] = 1. / cell; g.n[L][k] = std::max(1, (int)std::floor(ext[k] / cell) + 1); g.reach[L][k] = reach; }
        ncells += (long long)g.n[L][0] * g.n[L][1] * g.n[L][2];
        if (ncells > (1ll << 31) - 2) throw EngErr{"collider grid too large", WDEM_ERR_RUNTIME};
      }
      g.base[nL] = (int)ncells;
      // 3. cell keys
      k_cellkeys<<<nb, 256, 0, st>>>(E->S().posr, n, g, E->keys.p, E->idx.p, nDead ? E->S().flags : nullptr, (uint32_t)ncells); E->launches++;
    } else {
      // periodic: per level nc[k] cells along each (unsheared) cell axis; the hash wraps, positions do not
      const d3 sz = E->cell.size;
      const double rmin = E->minRadius, rmax = std::max(E->dom.rmaxGlobal, rmin);
      int nL = 1; while (nL < WD_MAX_LEVELS && rmin * (double)(1 << nL) <= rmax) nL++;
      g.nL = nL; g.periodic = 1; g.rmin = rmin;
      double fac[3] = {1, 1, 1};   // Cell::shearAlignedExtents of a unit extent (Cell.cpp:68-79)
      if (E->cell.hasShear) {
        const d3 cv = E->cell.cosv; double hx = 1, hy = 1, hz = 1;
        hy += .5 * (1 / cv.x - 1); hz += .5 * (1 / cv.x - 1); hz += .5 * (1 / cv.y - 1); hx += .5 * (1 / cv.y - 1); hx += .5 * (1 / cv.z - 1); hy += .5 * (1 / cv.z - 1);
        fac[0] = hx; fac[1] = hy; fac[2] = hz;
      }
      const double szk[3] = {sz.x, sz.y, sz.z};
      const double capCell = std::cbrt((sz.x * sz.y * sz.z) / (8.0 * n));
      for (int L = 0; L < nL; L++) {
        const double rL = (L == nL - 1) ? rmax : std::min(rmax, rmin * (double)(2 << L));
        g.base[L] = (int)ncells;
        for (int k = 0; k < 3; k++) {
          const double reach = (fac[k] * rL) * (1 + 1e-9) + vd;
          const double cell = std::max(2 * reach * (1 + 1e-9), capCell);
          g.reach[L][k] = reach; g.n[L][k] = (int)std::min(1024., std::max(1., std::floor(szk[k] / cell))); g.inv[L][k] = g.n[L][k] / szk[k];
        }
        ncells += (long long)g.n[L][0] * g.n[L][1] * g.n[L][2];
      }
      g.base[nL] = (int)ncells;
      k_cellkeys_peri<<<nb, 256, 0, st>>>(E->S().posr, n, g, pu, E->keys.p, E->idx.p, nDead ? E->S().flags : nullptr, (uint32_t)ncells); E->launches++;
    }
    // radix sort by cell, reorder all sphere arrays into cell order
    int keyBits = 1; while ((1ll << keyBits) < ncells + (nDead ? 1 : 0)) keyBits++;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, E->keys.p, E->keysAlt.p, E->idx.p, E->idxAlt.p, n, 0, keyBits, st);
    E->cubTemp.ensure(tb, tot);
    cub::DeviceRadixSort::SortPairs(E->cubTemp.p, tb, E->keys.p, E->keysAlt.p, E->idx.p, E->idxAlt.p, n, 0, keyBits, st); E->launches += 4;
    if (nDead) { n -= nDead; E->n = n; nb = nblk(n, 256); }      // the dead slots sorted to the end: gone
    SphereStore& dst = E->sph[1 - E->cur];
    dst.ensure(n, tot);
    k_reorder<<<nb, 256, 0, st>>>(dst.view(), E->S(), E->idxAlt.p, n, E->id2slot.p); E->launches++;
    E->cur = 1 - E->cur; E->up2slotValid = false;
    perm = E->idxAlt.p;
    if (E->comm || E->compactPending) {   // slots of the previous table's generation: through the pre-sort generation built by domainExchange / wdem_remove_spheres
      E->dom.perm2.ensure(n, tot);
      k_dom_compose<<<nb, 256, 0, st>>>(E->dom.origSlot.p, E->idxAlt.p, n, E->dom.perm2.p); E->launches++;
      perm = E->dom.perm2.p;
    }
    // 4. cellStart[k] = #spheres with key<k
    E->cellStart.ensure((size_t)ncells + 1, tot);
    CK(cudaMemsetAsync(E->cellStart.p, 0, ((size_t)ncells + 1) * sizeof(int), st));
    k_cellends<<<nb, 256, 0, st>>>(E->keysAlt.p, n, E->cellStart.p); E->launches++;
    tb = 0;
    cub::DeviceScan::InclusiveScan(nullptr, tb, E->cellStart.p, E->cellStart.p, cub::Max(), (int)(ncells + 1), st);
    E->cubTemp.ensure(tb, tot);
    cub::DeviceScan::InclusiveScan(E->cubTemp.p, tb, E->cellStart.p, E->cellStart.p, cub::Max(), (int)(ncells + 1), st); E->launches += 2;
    // 5. count candidates: [active SS by sphere | dormant SS by sphere | WS by sphere | FS by sphere]
    const int nWalls = (int)E->hWalls.size();
    E->counts.ensure((size_t)4 * n + 1, tot); E->offsets.ensure((size_t)4 * n + 1, tot);
    static const bool noDormant = getenv("WDEM_NO_DORMANT") != nullptr;
    const double adf = std::abs(c.distFactor), dormOn = (!noDormant && !c.periodic && !c.noBreak && vd > 0.) ? 1. : -1.;
    Contacts dummy; memset(&dummy, 0, sizeof(dummy));
    const int nb128 = nblk(n, 128);
    E->rowCount.ensure((size_t)n + 1, tot);
    CK(cudaMemsetAsync(E->rowCount.p, 0, ((size_t)n + 1) * sizeof(int), st));
    static const bool noCand = getenv("WDEM_NO_CAND_CACHE") != nullptr;
    int* cand = nullptr;
    if (!noCand) { E->candBuf.ensure((size_t)WD_CAND_CAP * n * (c.periodic ? 4 : 1), tot); cand = E->candBuf.p; }
    static const bool noWallMerge = getenv("WDEM_NO_WALL_MERGE") != nullptr;
    WallPass wp; memset(&wp, 0, sizeof(wp));
    const bool mergeWalls = !c.periodic && nWalls > 0 && nWalls <= 32 && !noWallMerge;
    if (mergeWalls) { E->wallBits.ensure((size_t)n, tot); wp.walls = E->dWalls.p; wp.nWalls = nWalls; wp.halfThick = c.wallHalfThick; wp.bits = E->wallBits.p; }
    if (!c.periodic) k_pairs_ss<false><<<nb128, 128, 0, st>>>(E->S(), n, g, E->cellStart.p, vd, c.loneMask, E->counts.p, nullptr, dummy, dummy, oc, perm, E->rowCount.p, nullptr, cand, wp, adf, dormOn);
    else {
      k_pairs_ss_peri<false><<<nb128, 128, 0, st>>>(E->S(), n, g, pu, E->cellStart.p, vd, c.loneMask, E->counts.p, nullptr, dummy, oc, perm, E->rowCount.p, nullptr, E->dCnt, cand);
      CK(cudaMemsetAsync(E->counts.p + n, 0, (size_t)n * sizeof(int), st));    // no dormant contacts in a periodic cell
    }
    E->launches++;
    if (nWalls && !mergeWalls) {
      if (!c.periodic) k_pairs_ws<false><<<nb128, 128, 0, st>>>(E->S(), n, E->dWalls.p, nWalls, vd, c.wallHalfThick, c.loneMask, E->counts.p + 2 * n, nullptr, dummy, dummy, oc, perm, E->rowCount.p, nullptr);
      else k_pairs_ws_peri<false><<<nb128, 128, 0, st>>>(E->S(), n, E->dWalls.p, nWalls, pu, vd, c.wallHalfThick, c.loneMask, E->counts.p + 2 * n, nullptr, dummy, oc, perm, E->rowCount.p, nullptr, E->dCnt);
      E->launches++;
    }
    else if (!nWalls) CK(cudaMemsetAsync(E->counts.p + 2 * n, 0, (size_t)n * sizeof(int), st));
    if (nFacets && c.periodic) {      // boxes of the unsheared node positions (they follow the cell's shear), then every sphere against every facet
      E->facetBoxes.ensure((size_t)6 * nFacets, tot);
      k_facet_boxes_peri<<<nblk(nFacets, 128), 128, 0, st>>>(E->dFacets.p, nFacets, pu, vd, E->facetBoxes.p);
      k_pairs_fs_peri<false><<<nb128, 128, 0, st>>>(E->S(), n, E->dFacets.p, nFacets, E->facetBoxes.p, pu, vd, c.loneMask, E->counts.p + 3 * n, nullptr, dummy, oc, perm, E->rowCount.p, nullptr, E->dCnt);
      E->launches += 2;
    }
    else if (nFacets) { k_pairs_fs<false><<<nb128, 128, 0, st>>>(E->S(), n, E->dFacets.p, E->fgrid, vd, c.loneMask, E->counts.p + 3 * n, nullptr, dummy, dummy, oc, perm, E->rowCount.p, nullptr); E->launches++; }
    else CK(cudaMemsetAsync(E->counts.p + 3 * n, 0, (size_t)n * sizeof(int), st));
    CK(cudaMemsetAsync(E->counts.p + 4 * n, 0, sizeof(int), st));
    tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, E->counts.p, E->offsets.p, 4 * n + 1, st);
    E->cubTemp.ensure(tb, tot);
    cub::DeviceScan::ExclusiveSum(E->cubTemp.p, tb, E->counts.p, E->offsets.p, 4 * n + 1, st); E->launches += 2;
    int* hOff = E->hScratch;   // pinned
    for (int k = 0; k < 4; k++) CK(cudaMemcpyAsync(&hOff[k], E->offsets.p + (size_t)(k + 1) * n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    nSSa = hOff[0]; nSS = hOff[1]; nWS = hOff[2] - hOff[1]; nFS = hOff[3] - hOff[2];
    if (hOff[3] < 0 || (long long)hOff[3] >= (1ll << 30)) throw EngErr{"too many potential contacts for one device", WDEM_ERR_RUNTIME};
    // 6. fill pairs + carry history
    const long long nC0 = nSS + nWS + nFS;
    Nst.ensure((size_t)std::max<long long>(nC0, 1), tot, hertz, rolling, c.periodic);
    Contacts N = Nst.view(), O = Ost.view();
    E->srcOld.ensure((size_t)std::max<long long>(nC0, 1), tot);
    // the fill kernels write (pair, kind, cellDist) in emission order into scratch planes; k_carry moves them to their final place
    const size_t nT = (size_t)std::max<long long>(nC0, 1);
    E->ckeys.ensure(nT, tot); E->cvals.ensure(nT, tot); E->realFlag.ensure(nT + 1, tot); E->realPos.ensure(nT + 1, tot);
    Contacts T = N; T.pair = reinterpret_cast<int2*>(E->ckeys.p); T.flags = E->cvals.p;
    if (c.periodic) { for (auto& b : E->tmpCd) b.ensure(nT, tot); T.cdx = E->tmpCd[0].p; T.cdy = E->tmpCd[1].p; T.cdz = E->tmpCd[2].p; }
    if (!c.periodic) k_pairs_ss<true><<<nb128, 128, 0, st>>>(E->S(), n, g, E->cellStart.p, vd, c.loneMask, E->counts.p, E->offsets.p, T, O, oc, perm, nullptr, E->srcOld.p, cand, wp, adf, dormOn);
    else k_pairs_ss_peri<true><<<nb128, 128, 0, st>>>(E->S(), n, g, pu, E->cellStart.p, vd, c.loneMask, E->counts.p, E->offsets.p, T, oc, perm, nullptr, E->srcOld.p, E->dCnt, cand);
    E->launches++;
    if (nWalls && !mergeWalls) {
      if (!c.periodic) k_pairs_ws<true><<<nb128, 128, 0, st>>>(E->S(), n, E->dWalls.p, nWalls, vd, c.wallHalfThick, c.loneMask, nullptr, E->offsets.p + 2 * n, T, O, oc, perm, nullptr, E->srcOld.p);
      else k_pairs_ws_peri<true><<<nb128, 128, 0, st>>>(E->S(), n, E->dWalls.p, nWalls, pu, vd, c.wallHalfThick, c.loneMask, nullptr, E->offsets.p + 2 * n, T, oc, perm, nullptr, E->srcOld.p, E->dCnt);
      E->launches++;
    }
    if (nFacets && c.periodic) { k_pairs_fs_peri<true><<<nb128, 128, 0, st>>>(E->S(), n, E->dFacets.p, nFacets, E->facetBoxes.p, pu, vd, c.loneMask, nullptr, E->offsets.p + 3 * n, T, oc, perm, nullptr, E->srcOld.p, E->dCnt); E->launches++; }
    else if (nFacets) { k_pairs_fs<true><<<nb128, 128, 0, st>>>(E->S(), n, E->dFacets.p, E->fgrid, vd, c.loneMask, nullptr, E->offsets.p + 3 * n, T, O, oc, perm, nullptr, E->srcOld.p); E->launches++; }
    if (nC0 > 0) {
      // real-first order inside the sphere-sphere segment (the reference reorders its contacts for the same reason, ContactLoop.cpp:75-91):
      // warps of k_contact_ss then see either all-real or all-potential contacts instead of paying the long path for half-empty warps
      k_was_real<<<nblk(nSSa + 1, 256), 256, 0, st>>>(E->srcOld.p, O.flags, nSSa, E->realFlag.p);
      exclusiveSum(E, E->realFlag.p, E->realPos.p, (int)nSSa + 1);
      k_carry<<<nblk(nC0, 256), 256, 0, st>>>(N, nC0, O, E->srcOld.p, T, E->realPos.p, nSSa, c.law, rolling, c.periodic, E->step); E->launches += 4;
    }
    // 6b. orphans: real contacts of the previous table that no candidate claimed (AABBs separated while real).  The counting pass adds their incidences to the row
    // counts; how many there are travels to the host while the GPU goes on with the rows of the grid candidates, so the wait costs no idle time.
    long long nOrph = 0;
    const bool orphPass = (oc.valid || imported) && E->nC > 0;
    const int32_t* oldSid = (E->comm || E->compactPending) ? E->dom.oldIds.p : E->sph[1 - E->cur].id.p;
    const uint32_t* sfl = E->comm ? E->S().flags : nullptr;
    int* hOrph = E->hScratch + 8; *hOrph = 0;
    if (orphPass) {
      E->orphCnt.ensure(1, tot);
      CK(cudaMemsetAsync(E->orphCnt.p, 0, sizeof(int), st));
      k_orphans<<<nblk(E->nC, 256), 256, 0, st>>>(O, E->nC, N, nC0, oldSid, E->id2slot.p, sfl, c.law, rolling, c.periodic, E->orphCnt.p, E->rowCount.p, 0); E->launches++;
      CK(cudaMemcpyAsync(hOrph, E->orphCnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(E->evOrph, st));
    }
    if (nC0 > 0) { k_realbits<<<nblk(nC0, 256), 256, 0, st>>>(N.flags, 0, nC0, N.realBits); E->launches++; }
    // 7. per-particle incidence rows sorted by the other particle's id: scan of row counts, cursor fill, per-row sort
    const long long slack = 4096;
    Ncsr.rowStart.ensure((size_t)n + 2, tot); Ncsr.rowEntry.ensure((size_t)(2 * (nC0 + slack)), tot); Ncsr.rowOther.ensure((size_t)(2 * (nC0 + slack)), tot);
    tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, E->rowCount.p, Ncsr.rowStart.p, n + 1, st);
    E->cubTemp.ensure(tb, tot);
    cub::DeviceScan::ExclusiveSum(E->cubTemp.p, tb, E->rowCount.p, Ncsr.rowStart.p, n + 1, st); E->launches += 2;
    CK(cudaMemsetAsync(E->rowCount.p, 0, ((size_t)n + 1) * sizeof(int), st));   // reused as the per-row cursor
    if (nC0 > 0) { k_csr_fill<<<nblk(nC0, 256), 256, 0, st>>>(N, 0, nC0, E->S().id, E->dWalls.p, E->dFacets.p, Ncsr.rowStart.p, E->rowCount.p, Ncsr.rowEntry.p, Ncsr.rowOther.p); E->launches++; }
    if (orphPass) {
      CK(cudaEventSynchronize(E->evOrph));
      if (*hOrph > 0) {
        nOrph = *hOrph;
        Nst.ensure((size_t)(nC0 + nOrph), tot, hertz, rolling, c.periodic, /*keep*/ true, st);
        N = Nst.view();
        Ncsr.rowEntry.ensure((size_t)(2 * (nC0 + nOrph)), tot, /*keep*/ true, st); Ncsr.rowOther.ensure((size_t)(2 * (nC0 + nOrph)), tot, /*keep*/ true, st);
        CK(cudaMemsetAsync(E->orphCnt.p, 0, sizeof(int), st));
        k_orphans<<<nblk(E->nC, 256), 256, 0, st>>>(O, E->nC, N, nC0, oldSid, E->id2slot.p, sfl, c.law, rolling, c.periodic, E->orphCnt.p, E->rowCount.p, 1);
        const long long w0 = nC0 & ~31ll;
        k_realbits<<<nblk(nC0 + nOrph - w0, 256), 256, 0, st>>>(N.flags, w0, nC0 + nOrph, N.realBits);
        k_csr_fill<<<nblk(nOrph, 256), 256, 0, st>>>(N, nC0, nC0 + nOrph, E->S().id, E->dWalls.p, E->dFacets.p, Ncsr.rowStart.p, E->rowCount.p, Ncsr.rowEntry.p, Ncsr.rowOther.p);
        E->launches += 3;
      }
    }
    const long long nC = nC0 + nOrph;
    nOrphOut = nOrph;
    if (nC > 0) {
      E->idx.ensure((size_t)n + 1, tot);   // the cell-sort scratch is free again: queue of long rows + its counter
      CK(cudaMemsetAsync(E->idx.p + n, 0, sizeof(int), st));
      k_row_sort<<<nb128, 128, 0, st>>>(n, Ncsr.rowStart.p, Ncsr.rowEntry.p, Ncsr.rowOther.p, E->idx.p, E->idx.p + n);
      k_row_sort_long<<<148 * 8, 128, 0, st>>>(E->idx.p, E->idx.p + n, Ncsr.rowStart.p, Ncsr.rowEntry.p, Ncsr.rowOther.p); E->launches += 2;
    }
    Ncsr.valid = true; Ncsr.nRows = n;
  }
  if (E->anyMovingWall) { k_walls_step<<<nblk((long long)E->hWalls.size(), 32), 32, 0, st>>>(E->dWalls.p, (int)E->hWalls.size(), makeUniforms(E), E->dCnt, 1, nullptr); E->launches++; }   // Aabb::nodeLastPos (Collision.cpp:138-141)
  E->nSS = nSS; E->nSSa = nSSa; E->nWS = nWS; E->nFS = nFS; E->nOrph = nOrphOut; E->nC = nSS + nWS + nFS + nOrphOut;
  E->curC = 1 - E->curC; E->curR = 1 - E->curR;
  E->nFullRuns++; E->compactPending = false;
  auto t2 = now();
  if (E->comm) domainFinishRebuild(E);
  if (traceReb) {
    auto t3 = now(); auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    static double acc[3] = {0, 0, 0}; static int cnt = 0; acc[0] += ms(t0, t1); acc[1] += ms(t1, t2); acc[2] += ms(t2, t3); cnt++;
    if (cnt % 25 == 0) { fprintf(stderr, "rebuild trace (rank %d, mean of the last 25 of %d): domainExchange %.3f ms, collider run %.3f ms, finish %.3f ms\n", E->comm ? E->comm->rank : 0, cnt, acc[0] / 25, acc[1] / 25, acc[2] / 25); acc[0] = acc[1] = acc[2] = 0; }
  }
}

// Static facets are binned once on the host (they never move in this engine: Impose/clump motion is out of scope).
void buildFacetGrid(wdem_engine* E) {
  const int nf = (int)E->hFacets.size();
  FacetGrid& g = E->fgrid; memset(&g, 0, sizeof(g));
  if (nf == 0) { E->fgridValid = true; return; }
  const double vd = E->verletDist;
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  std::vector<double> ext; ext.reserve(nf);
  for (auto& f : E->hFacets) { double e = 0; for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], f.bmin[k] - vd); hi[k] = std::max(hi[k], f.bmax[k] + vd); e = std::max(e, f.bmax[k] - f.bmin[k]); } ext.push_back(e); }
  // a rotating group sweeps its bounding sphere: size the (fixed) grid for it; a group that translates out of the grid is still binned correctly (clamped
  // cell coordinates on both sides of the lookup), only less efficiently
  for (int gi = 0; gi < E->groups.n; gi++) {
    const d3 c = E->groups.g[gi].pos; const double R = E->groupRadius[gi] + vd; const double cc3[3] = {c.x, c.y, c.z};
    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], cc3[k] - R); hi[k] = std::max(hi[k], cc3[k] + R); }
  }
  std::nth_element(ext.begin(), ext.begin() + nf / 2, ext.end());
  double cell = std::max(ext[nf / 2] + 2 * vd, 1e-300);
  double vol = std::max(hi[0] - lo[0], cell) * std::max(hi[1] - lo[1], cell) * std::max(hi[2] - lo[2], cell);
  cell = std::max(cell, std::cbrt(vol / (16.0 * nf + 64)));
  g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.inv = 1. / cell;
  g.nx = std::max(1, (int)std::floor((hi[0] - lo[0]) * g.inv) + 1); g.ny = std::max(1, (int)std::floor((hi[1] - lo[1]) * g.inv) + 1); g.nz = std::max(1, (int)std::floor((hi[2] - lo[2]) * g.inv) + 1);
  auto cc = [&](double x, double o, int n) { int c = (int)std::floor((x - o) * g.inv); return c < 0 ? 0 : (c >= n ? n - 1 : c); };
  const size_t ncell = (size_t)g.nx * g.ny * g.nz;
  std::vector<int> start(ncell + 1, 0);
  for (int pass = 0; pass < 2; pass++) {
    std::vector<int> fill;
    std::vector<int> list;
    if (pass == 1) { for (size_t k = 0; k < ncell; k++) start[k + 1] += start[k]; fill.assign(start.begin(), start.end() - 1); list.resize(start[ncell]); }
    for (int f = 0; f < nf; f++) {
      const Facet& F = E->hFacets[f];
      int x0 = cc(F.bmin[0] - vd, g.ox, g.nx), x1 = cc(F.bmax[0] + vd, g.ox, g.nx), y0 = cc(F.bmin[1] - vd, g.oy, g.ny), y1 = cc(F.bmax[1] + vd, g.oy, g.ny), z0 = cc(F.bmin[2] - vd, g.oz, g.nz), z1 = cc(F.bmax[2] + vd, g.oz, g.nz);
      for (int iz = z0; iz <= z1; iz++) for (int iy = y0; iy <= y1; iy++) for (int ix = x0; ix <= x1; ix++) {
        size_t k = ((size_t)iz * g.ny + iy) * g.nx + ix;
        if (pass == 0) start[k + 1]++; else list[fill[k]++] = f;
      }
    }
    if (pass == 1) {
      E->fgCellStart.ensure(ncell + 1, E->bytes); E->fgCellFacet.ensure(std::max<size_t>(list.size(), 1), E->bytes);
      CK(cudaMemcpyAsync(E->fgCellStart.p, start.data(), (ncell + 1) * sizeof(int), cudaMemcpyHostToDevice, E->stream));
      if (!list.empty()) CK(cudaMemcpyAsync(E->fgCellFacet.p, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
      CK(cudaStreamSynchronize(E->stream));
    }
  }
  g.cellStart = E->fgCellStart.p; g.cellFacet = E->fgCellFacet.p;
  E->fgridValid = true;
}

// full run with moving facets: refresh their boxes and re-bin them on the device into the grid laid out by buildFacetGrid (same geometry, new lists)
void rebinFacets(wdem_engine* E) {
  const int nf = (int)E->hFacets.size(); cudaStream_t st = E->stream; FacetGrid& g = E->fgrid;
  const size_t ncell = (size_t)g.nx * g.ny * g.nz;
  k_facets_step<<<nblk(nf, 128), 128, 0, st>>>(E->dFacets.p, nf, E->groups, makeUniforms(E), E->dCnt, 1, E->impDev(0));
  CK(cudaMemsetAsync(E->fgCellStart.p, 0, (ncell + 1) * sizeof(int), st));
  k_fgrid_bin<<<nblk(nf, 128), 128, 0, st>>>(E->dFacets.p, nf, g, E->verletDist, E->fgCellStart.p, nullptr, nullptr, 0);
  exclusiveSum(E, E->fgCellStart.p, E->fgCellStart.p, (int)ncell + 1);
  int total = 0; CK(cudaMemcpyAsync(&total, E->fgCellStart.p + ncell, sizeof(int), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
  E->fgCellFacet.ensure((size_t)std::max(total, 1), E->bytes); E->fgCursor.ensure(ncell + 1, E->bytes);
  g.cellStart = E->fgCellStart.p; g.cellFacet = E->fgCellFacet.p;
  CK(cudaMemsetAsync(E->fgCursor.p, 0, (ncell + 1) * sizeof(int), st));
  k_fgrid_bin<<<nblk(nf, 128), 128, 0, st>>>(E->dFacets.p, nf, g, E->verletDist, nullptr, E->fgCursor.p, E->fgCellFacet.p, 1);
  E->launches += 5;
}

void ensureVerletAndDt(wdem_engine* E) {
  const wdem_config& c = E->cfg;
  syncGlobals(E);
  // AabbCollider::setVerletDist (Collision.cpp:164-184)
  if (std::isnan(E->verletDist)) {
    if (c.verletDist >= 0) E->verletDist = c.verletDist;
    else if (std::isfinite(E->minRadius)) E->verletDist = std::abs(c.verletDist) * E->minRadius;
    else E->verletDist = 0.;
  }
  // Scene::doOneStep (Scene.cpp:380-398): dt = dtSafety * critDt when unset
  if (std::isnan(E->dt) || std::isinf(E->dt) || E->dt < 0) {
    if (std::isinf(E->pwaveDt)) throw EngErr{"Failed to obtain meaningful dt from engines and fields automatically", WDEM_ERR_RUNTIME};
    E->dt = c.dtSafety * E->pwaveDt;
  }
}

void fetchCounters(wdem_engine* E) {
  CK(cudaMemcpyAsync(E->hCnt, E->dCnt, sizeof(Counters), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
}

double dynDtCrit(wdem_engine* E) {
  const int n = E->n; if (n == 0) return INFINITY;
  const int nbS = nblk(n, 256), nbK = nblk(E->clumps.n, 256), nb = nbS + nbK;
  E->blk.ensure((size_t)std::max(nb, 1) * 7, E->bytes);
  CsrStore& R = E->csr[E->curR];
  k_dyndt<<<nbS, 256, 0, E->stream>>>(E->S(), n, E->con[E->curC].view(), R.rowStart.p, R.rowEntry.p, E->cfg.law, E->blk.p); E->launches++;
  if (nbK) { k_clumps_dyndt<<<nbK, 256, 0, E->stream>>>(E->clumps.view(), E->S(), E->id2slot.p, E->con[E->curC].view(), R.rowStart.p, R.rowEntry.p, E->cfg.law, E->blk.p + nbS); E->launches++; }
  std::vector<double> hb(nb);
  CK(cudaMemcpyAsync(hb.data(), E->blk.p, nb * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  double r = INFINITY; for (double v : hb) r = std::min(r, v);
  return std::sqrt(r);
}
double dynDtCritGlobal(wdem_engine* E) {
  double r = dynDtCrit(E);
  if (!E->comm) return r;
  const int W = E->comm->world; int64_t m; memcpy(&m, &r, 8); std::vector<int64_t> all(W);
  E->comm->allgather64(&m, 1, all.data(), E->stream);
  for (int k = 0; k < W; k++) { double v; memcpy(&v, &all[k], 8); r = std::min(r, v); }
  return r;
}
double pown(double x, unsigned n) { unsigned d = n >> 1, r = n & 1; double a = d == 0 ? 1 : pown(x * x, d); double b = r == 0 ? 1 : x; return a * b; }  // Math.hpp:161-176

void harvestPhases(wdem_engine* E, int set) {
  if (!E->evPending[set]) return;
  CK(cudaEventSynchronize(E->evr[set][5]));
  float ms;
  for (int k = 0; k < 5; k++) { CK(cudaEventElapsedTime(&ms, E->evr[set][k], E->evr[set][k + 1])); E->phaseAcc[k] += ms; }
  E->evPending[set] = false;
}

// ContactLoop + force reduction of one step.  `guard` points at the device-side Verlet-dirty flag: kernels launched before the host knows whether a
// collider full run is due return immediately when it is set (they are launched again after the full run).
void launchContactsAndReduce(wdem_engine* E, const StepU& u, const int* guard, cudaEvent_t* ev) {
  const wdem_config& c = E->cfg; cudaStream_t st = E->stream, st2 = E->stream2;
  Contacts C = E->con[E->curC].view();
  static const bool sideOff = getenv("WDEM_NO_SIDE_STREAM") != nullptr;
  const bool side = !sideOff && (E->nWS + E->nFS + E->nOrph) > 0 && E->nSSa > 200000;   // small scenes: the fork/join costs more host time than the overlap wins   // boundary kernels are small and latency-bound: overlap them with the big one
  cudaStream_t sb = side ? st2 : st;
  if (ev) CK(cudaEventRecord(ev[2], st));
  if (side) { CK(cudaEventRecord(E->evFork, st)); CK(cudaStreamWaitEvent(st2, E->evFork, 0)); }
  // forces on walls at rest accumulate (never reset in the reference): summed by the wall-contact kernels themselves (wall_acc_block)
  WallAcc wa; memset(&wa, 0, sizeof(wa)); WallAcc wo = wa;
  if (E->anyStaticWall && c.applyForces && (E->nWS + E->nOrph) > 0) {
    const int nW = std::min((int)E->hWalls.size(), WD_ACC_WALLS), nbW = nblk(E->nWS, 128), nbO = nblk(E->nOrph, 128);
    E->wallPartial.ensure((size_t)(nbW + nbO + 1) * nW * 6, E->bytes);
    wa.nWalls = nW; wa.nBlocksTotal = nbW + nbO; wa.blockBase = 0; wa.foldInKernel = (nbW + nbO) <= 96 ? 1 : 0; wa.partial = E->wallPartial.p; wa.acc = E->wallAcc.p; wa.ticket = reinterpret_cast<unsigned*>(E->wallAcc.p + E->hWalls.size() * 6);
    wo = wa; wo.blockBase = nbW;
  }
  if (E->nWS > 0) {
    switch (c.law + (u.energy ? 2 : 0)) {
      case 0: k_contact_ws<0><<<nblk(E->nWS, 128), 128, 0, sb>>>(C, E->nSS, E->nWS, E->S(), E->dWalls.p, E->dMats.p, u, E->dCnt, guard, wa); break;
      case 1: k_contact_ws<1><<<nblk(E->nWS, 128), 128, 0, sb>>>(C, E->nSS, E->nWS, E->S(), E->dWalls.p, E->dMats.p, u, E->dCnt, guard, wa); break;
      case 2: k_contact_ws<2><<<nblk(E->nWS, 128), 128, 0, sb>>>(C, E->nSS, E->nWS, E->S(), E->dWalls.p, E->dMats.p, u, E->dCnt, guard, wa); break;
      case 3: k_contact_ws<3><<<nblk(E->nWS, 128), 128, 0, sb>>>(C, E->nSS, E->nWS, E->S(), E->dWalls.p, E->dMats.p, u, E->dCnt, guard, wa); break;
    }
    E->launches++;
  }
  if (E->nFS > 0) {
    switch (c.law + (u.energy ? 2 : 0)) {
      case 0: k_contact_fs<0><<<nblk(E->nFS, 128), 128, 0, sb>>>(C, E->nSS + E->nWS, E->nFS, E->S(), E->dFacets.p, E->dMats.p, u, E->dCnt, guard); break;
      case 1: k_contact_fs<1><<<nblk(E->nFS, 128), 128, 0, sb>>>(C, E->nSS + E->nWS, E->nFS, E->S(), E->dFacets.p, E->dMats.p, u, E->dCnt, guard); break;
      case 2: k_contact_fs<2><<<nblk(E->nFS, 128), 128, 0, sb>>>(C, E->nSS + E->nWS, E->nFS, E->S(), E->dFacets.p, E->dMats.p, u, E->dCnt, guard); break;
      case 3: k_contact_fs<3><<<nblk(E->nFS, 128), 128, 0, sb>>>(C, E->nSS + E->nWS, E->nFS, E->S(), E->dFacets.p, E->dMats.p, u, E->dCnt, guard); break;
    }
    E->launches++;
  }
  if (E->nOrph > 0) {
    const long long f0 = E->nSS + E->nWS + E->nFS;
    switch (c.law + (u.energy ? 2 : 0)) {
      case 0: k_contact_orphans<0><<<nblk(E->nOrph, 128), 128, 0, sb>>>(C, f0, E->nOrph, E->S(), E->dWalls.p, E->dFacets.p, E->dMats.p, u, E->dCnt, guard, wo); break;
      case 1: k_contact_orphans<1><<<nblk(E->nOrph, 128), 128, 0, sb>>>(C, f0, E->nOrph, E->S(), E->dWalls.p, E->dFacets.p, E->dMats.p, u, E->dCnt, guard, wo); break;
      case 2: k_contact_orphans<2><<<nblk(E->nOrph, 128), 128, 0, sb>>>(C, f0, E->nOrph, E->S(), E->dWalls.p, E->dFacets.p, E->dMats.p, u, E->dCnt, guard, wo); break;
      case 3: k_contact_orphans<3><<<nblk(E->nOrph, 128), 128, 0, sb>>>(C, f0, E->nOrph, E->S(), E->dWalls.p, E->dFacets.p, E->dMats.p, u, E->dCnt, guard, wo); break;
    }
    E->launches++;
  }
  // the big kernel goes last so that the small boundary kernels already occupy their share of the SMs
  if (E->nSSa > 0) {
    const int gridSS = std::min(nblk(E->nSSa, WD_CONTACT_BLOCK), E->numSMs * WD_CONTACT_CTAS_PER_SM);   // persistent: one CTA per resident slot
    switch (c.law + (u.energy ? 2 : 0)) {
      case 0: k_contact_ss<0><<<gridSS, WD_CONTACT_BLOCK, 0, st>>>(C, 0, E->nSSa, E->S(), E->dMats.p, u, E->dCnt, guard); break;
      case 1: k_contact_ss<1><<<gridSS, WD_CONTACT_BLOCK, 0, st>>>(C, 0, E->nSSa, E->S(), E->dMats.p, u, E->dCnt, guard); break;
      case 2: k_contact_ss<2><<<gridSS, WD_CONTACT_BLOCK, 0, st>>>(C, 0, E->nSSa, E->S(), E->dMats.p, u, E->dCnt, guard); break;
      case 3: k_contact_ss<3><<<gridSS, WD_CONTACT_BLOCK, 0, st>>>(C, 0, E->nSSa, E->S(), E->dMats.p, u, E->dCnt, guard); break;
    }
    E->launches++;
  }
  if (side) { CK(cudaEventRecord(E->evJoin, st2)); CK(cudaStreamWaitEvent(st, E->evJoin, 0)); }
  if (ev) CK(cudaEventRecord(ev[3], st));
  const int nNow = E->n;   // a full run of a decomposed scene changes the local sphere count (migration, ghosts)
  if (nNow > 0 && (c.applyForces || c.reset)) {
    // ContactLoop.applyForces=False: no contact force is applied, but Leapfrog.reset still sets force = gravity*mass, torque = 0 (Leapfrog.cpp:240-245): null rows
    CsrStore& R = E->csr[E->curR];
    k_reduce<<<nblk(nNow, WD_REDUCE_BLOCK), WD_REDUCE_BLOCK, 0, st>>>(E->S(), nNow, C, c.applyForces ? R.rowStart.p : nullptr, R.rowEntry.p, u, guard, E->impDev(u.impSlot));
    E->launches++;
  }
  if (wa.nWalls && !wa.foldInKernel) { k_wall_fold<<<wa.nWalls * 6, 256, 0, st>>>(wa.partial, wa.nBlocksTotal, wa.nWalls * 6, wa.acc, guard); E->launches++; }
  if (ev) CK(cudaEventRecord(ev[4], st));
}

// Leapfrog of the end-to-end path: the host state arrives in NCHUNK pieces on a copy stream; each piece is integrated as soon as it has landed
// (k_leapfrog_io, upload order) and its new pos/vel start their way back on a second copy stream while later pieces are still being uploaded
// and, after the last piece, while the contact kernels run.  PCIe is full duplex, so the step costs about max(H2D, D2H) + contacts.
void leapfrogPipelined(wdem_engine* E, const StepU& u) {
  const long long n = E->e2e.count; cudaStream_t st = E->stream;      // spheres the host sends: all of them, or the subdomain's owned ones (wdem_step_e2e_owned)
  const wdem_engine::E2E& io = E->e2e;
  double* in = E->ioIn.p; double* out = E->ioOut.p;
  LeapIO lio{in, in + 3 * n, in + 6 * n, out, out + 3 * n};
  const int NCH = E->nChunk;
  const long long per = (n + NCH - 1) / NCH;
  // the staging buffers are free: the previous call waited for its copies; order the copy streams behind whatever the main stream did before
  // Deferred tail (wdem_step_e2e without stats): the previous call returned while its contact pass was still running; the staging buffer was last read by that call's
  // k_leapfrog_io launches (evLf), so this step's upload only waits for those and overlaps the contact pass of the previous step.
  if (io.defer && io.ioBusy) CK(cudaStreamWaitEvent(E->sH2D, E->evLf[E->lastLfChunk], 0));
  else { CK(cudaEventRecord(E->evFork, st)); CK(cudaStreamWaitEvent(E->sH2D, E->evFork, 0)); }
  for (int c = 0; c < NCH; c++) {
    const long long first = c * per, cnt = std::min(per, n - first);
    if (cnt <= 0) break;
    CK(cudaMemcpyAsync(in + 3 * first, io.pos + 3 * first, cnt * 24, cudaMemcpyHostToDevice, E->sH2D));
    CK(cudaMemcpyAsync(in + 3 * n + 3 * first, io.vel + 3 * first, cnt * 24, cudaMemcpyHostToDevice, E->sH2D));
    CK(cudaMemcpyAsync(in + 6 * n + 3 * first, io.angVel + 3 * first, cnt * 24, cudaMemcpyHostToDevice, E->sH2D));
    CK(cudaEventRecord(E->evUp[c], E->sH2D));
  }
  for (int c = 0; c < NCH; c++) {
    const long long first = c * per, cnt = std::min(per, n - first);
    if (cnt <= 0) break;
    CK(cudaStreamWaitEvent(st, E->evUp[c], 0));
    if (u.energy) k_leapfrog_io<true><<<nblk(cnt, 256), 256, 0, st>>>(E->S(), u, E->dCnt, E->e2e.slotMap, first, cnt, lio);
    else k_leapfrog_io<false><<<nblk(cnt, 256), 256, 0, st>>>(E->S(), u, E->dCnt, E->e2e.slotMap, first, cnt, lio);
    E->launches++;
    CK(cudaEventRecord(E->evLf[c], st)); E->lastLfChunk = c; E->e2e.ioBusy = true;
    CK(cudaStreamWaitEvent(E->sD2H, E->evLf[c], 0));
    if (io.posOut) CK(cudaMemcpyAsync(io.posOut + 3 * first, out + 3 * first, cnt * 24, cudaMemcpyDeviceToHost, E->sD2H));
    if (io.velOut) CK(cudaMemcpyAsync(io.velOut + 3 * first, out + 3 * n + 3 * first, cnt * 24, cudaMemcpyDeviceToHost, E->sD2H));
  }
  // kinetic energy of the new state (velocities are final after Leapfrog): block sums, summed by the host in a fixed order
  const int nb = nblk(E->n, 256); E->blk.ensure((size_t)nb * 7, E->bytes); E->ekBlocks = nb;
  if ((size_t)nb > E->hEkCap) { if (E->hEk) cudaFreeHost(E->hEk); CK(cudaMallocHost(&E->hEk, (size_t)nb * sizeof(double))); E->hEkCap = nb; }
  k_ekin<<<nb, 256, 0, st>>>(E->S(), E->n, u, E->blk.p); E->launches++;      // owned spheres only (ghosts are skipped)
  CK(cudaEventRecord(E->evEk, st)); CK(cudaStreamWaitEvent(E->sD2H, E->evEk, 0));
  CK(cudaMemcpyAsync(E->hEk, E->blk.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, E->sD2H));
}

template <class T> std::vector<T> d2h(const T* p, size_t n, cudaStream_t st);
// ContactLoop.stress (ContactLoop.cpp:116,218-231,243-244) of the state the last contact pass left: block partial sums on the device, added up in block order
void computeStress(wdem_engine* h, double* stress9) {
  if (!h->cfg.periodic) throw EngErr{"ContactLoop.evalStress needs a periodic scene (ContactLoop.cpp:123)", WDEM_ERR_INVALID};
  for (int k = 0; k < 9; k++) stress9[k] = 0.;
  const long long nC = h->nC;
  if (h->compactPending) throw EngErr{"spheres were removed since the last step: step once before reading contacts, wall forces or the stress (the table still refers to the old slots)", WDEM_ERR_INVALID};
  if (nC > 0) {
    const int nb = nblk(nC, 256); h->blk.ensure((size_t)nb * 9 + 1, h->bytes);
    int* flag = reinterpret_cast<int*>(h->blk.p + (size_t)nb * 9);
    CK(cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
    k_stress<<<nb, 256, 0, h->stream>>>(h->con[h->curC].view(), nC, h->S(), h->dWalls.p, makeUniforms(h), h->blk.p, flag);
    auto hb = d2h(h->blk.p, (size_t)nb * 9 + 1, h->stream);
    int hf; memcpy(&hf, &hb[(size_t)nb * 9], sizeof(int));
    if (hf) throw EngErr{"ContactLoop.trackWork not allowed with multi-nodal particles in contact (facets; ContactLoop.cpp:220)", WDEM_ERR_UNSUPPORTED};
    for (int b = 0; b < nb; b++) for (int k = 0; k < 9; k++) stress9[k] += hb[9 * (size_t)b + k];
  }
  const double vol = mDet(h->cell.hSize);
  for (int k = 0; k < 9; k++) stress9[k] /= vol;
}
// gradV work (ContactLoop.cpp:243-254), once per step after the contact pass when Scene.trackEnergy and ContactLoop.evalStress are both on
void gradVWorkStep(wdem_engine* E) {
  if (!(E->cfg.periodic && E->cfg.evalStress)) return;
  if (!E->trackEnergy) { E->prevVol = NAN; return; }      // doStress without trackEnergy: the reference still updates prevStress/prevVol; restarting the mean is the same to O(dt)
  double s9[9]; computeStress(E, s9);
  const double vol = mDet(E->cell.hSize);
  const double midVol = !std::isnan(E->prevVol) ? .5 * (E->prevVol + vol) : vol;
  m33 mid; for (int k = 0; k < 9; k++) mid.m[k] = .5 * (s9[k] + E->prevStress[k]);
  const m33 gs = mMul(E->cell.gradV, mid);
  E->gradVWork += -((gs.m[0] + gs.m[4]) + gs.m[8]) * E->dt * midVol;
  E->prevVol = vol; for (int k = 0; k < 9; k++) E->prevStress[k] = s9[k];
}

// ---- one step in pieces, so that several steps can be enqueued before the host looks at the Verlet flags (Counters::gflag) ----
struct StepSnap { HostCell cell; GroupsU groups; long long step; double time; StepU u; std::vector<double> wallVel; };

// linearInterpolate (supp/smoothing/LinearInterpolate.hpp:13-52), `w` values per time point
void linInterp(double t, const std::vector<double>& tt, const std::vector<double>& val, int w, size_t& _pos, double* out) {
  size_t pos = _pos; const size_t n = tt.size();
  const double *A, *B; double rel;
  if (t <= tt[0]) { _pos = 0; A = &val[0]; B = &val[(n > 1 ? 1 : 0) * w]; rel = 0.; }
  else if (t >= tt[n - 1]) { _pos = n - 2; A = n > 1 ? &val[pos * w] : &val[(n - 1) * w]; B = &val[(n - 1) * w]; rel = 1.; }
  else {
    pos = std::min(pos, n - 2);
    while ((tt[pos] > t) || (tt[pos + 1] < t)) { if (tt[pos] > t) pos--; else pos++; }
    _pos = pos; A = &val[pos * w]; B = &val[(pos + 1) * w]; rel = (t - tt[pos]) / (tt[pos + 1] - tt[pos]);
  }
  for (int k = 0; k < w; k++) out[k] = (rel == 0.) ? A[k] : (rel == 1.) ? B[k] : A[k] + (B[k] - A[k]) * rel;
}
// linearInterpolateRel (LinearInterpolate.hpp:13-28): indices of the two end points and the relative position between them; the `A` of the t >= last branch is taken
// at the cookie's OLD value, as there
void linRel(double t, const std::vector<double>& tt, size_t& _pos, size_t& iA, size_t& iB, double& rel) {
  size_t pos = _pos; const size_t n = tt.size();
  if (t <= tt[0]) { _pos = 0; iA = 0; iB = n > 1 ? 1 : 0; rel = 0.; return; }
  if (t >= tt[n - 1]) { _pos = n - 2; iA = n > 1 ? pos : n - 1; iB = n - 1; rel = 1.; return; }
  pos = std::min(pos, n - 2);
  while ((tt[pos] > t) || (tt[pos + 1] < t)) { if (tt[pos] > t) pos--; else pos++; }
  _pos = pos; iA = pos; iB = pos + 1; rel = (t - tt[pos]) / (tt[pos + 1] - tt[pos]);
}
// Quaternionr::slerp (Eigen/src/Geometry/Quaternion.h); coefficients x,y,z,w; the 4-term dot in the order of Eigen's Packet2d reduction: (xx'+zz') + (yy'+ww')
void quatSlerp(const double* a, const double* b, double t, double* out) {
  const double one = 1. - std::numeric_limits<double>::epsilon();
  const double d = (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + a[3] * b[3]), absD = std::abs(d);
  double s0, s1;
  if (absD >= one) { s0 = 1. - t; s1 = t; }
  else { const double theta = std::acos(absD), sinTheta = std::sin(theta); s0 = std::sin((1. - t) * theta) / sinTheta; s1 = std::sin(t * theta) / sinTheta; }
  if (d < 0.) s1 = -s1;
  for (int k = 0; k < 4; k++) out[k] = s0 * a[k] + s1 * b[k];
}
// the impositions of the coming step: everything that depends on Scene.time / Scene.dt is evaluated here (libm), walls and facet-group masters are moved on
void imposePrepareStep(wdem_engine* E, int slot) {
  ImpSlot& H = E->hImpRing[slot]; memset(&H, 0, offsetof(ImpSlot, acc));
  const double time = E->time, dt = E->dt; const double NaNv = std::numeric_limits<double>::quiet_NaN();
  for (size_t i = 0; i < E->imposes.size(); i++) {
    wdem_engine::ImposeHost& d = E->imposes[i]; ImpOp& o = H.op[i];
    for (int k = 0; k < 4; k++) o.ori[k] = d.ori[k];
    switch (d.kind) {
      case 1: { const double omega = 2 * M_PI * d.v[0]; o.kind = IOP_DIR; o.flags = (d.flags & 1) ? 0 : 1; o.a[0] = d.v[2]; o.a[1] = d.v[3]; o.a[2] = d.v[4]; o.a[3] = d.v[1] * omega * std::cos(omega * (time - d.v[5])); break; }
      case 2: o.kind = IOP_AXES_SET; for (int ax = 0; ax < 3; ax++) { if (std::isnan(d.v[ax]) || std::isnan(d.v[3 + ax])) { o.a[ax] = NaNv; continue; } const double omega = 2 * M_PI * d.v[ax]; o.a[ax] = d.v[3 + ax] * omega * std::cos(omega * time); } break;
      case 3: o.kind = IOP_FADD; for (int k = 0; k < 3; k++) o.a[k] = d.v[k]; break;
      case 4: { o.kind = IOP_FRADIAL; const d3 A = mk3(d.v[0], d.v[1], d.v[2]), B = mk3(d.v[3], d.v[4], d.v[5]); const d3 ab = B - A; const d3 ax = ab / nrm(ab);
                o.a[0] = A.x; o.a[1] = A.y; o.a[2] = A.z; o.a[3] = ax.x; o.a[4] = ax.y; o.a[5] = ax.z; o.a[6] = d.v[6]; break; }
      case 5: o.kind = IOP_LOCAL6; for (int k = 0; k < 6; k++) { o.a[k] = d.v[k]; o.m[k] = (int)d.v[6 + k]; } break;
      case 6: o.kind = IOP_DIR; o.flags = ((d.flags & 1) ? 1 : 0) | 2 | ((d.flags & 2) ? 4 : 0); for (int k = 0; k < 4; k++) o.a[k] = d.v[k]; break;
      case 7: o.kind = IOP_READFT; o.flags = d.flags & 1; for (int k = 0; k < 3; k++) o.a[k] = d.v[k]; break;
      case 8: {
        double t = time + .5 * dt - d.v[0];
        if ((d.flags & 2) && (t < 0 || t > d.times.back())) { const double norm = t / d.times.back(); t = (norm - std::floor(norm)) * d.times.back(); }
        o.kind = IOP_VAR3D; const bool rot = !(d.ori[3] == 1 && d.ori[0] == 0 && d.ori[1] == 0 && d.ori[2] == 0);
        o.flags = (d.flags & 1) | (rot ? 2 : 0) | (d.vals2.empty() ? 0 : 4);
        linInterp(t, d.times, d.vals, 3, d.ipos, o.a);
        if (!d.vals2.empty()) linInterp(t, d.times, d.vals2, 3, d.ipos2, o.a + 3);
        break;
      }
      case 12: {      // InterpolatedMotion: where the node has to be at the end of this step (Impose.cpp:121-125); one lookup cookie for both tables, as there
        const double nextT = time + dt - d.v[0];
        o.kind = IOP_INTERP; o.a[3] = dt;
        linInterp(nextT, d.times, d.vals, 3, d.ipos, o.a);
        size_t iA, iB; double rel; linRel(nextT, d.times, d.ipos, iA, iB, rel);
        quatSlerp(&d.vals2[4 * iA], &d.vals2[4 * iB], rel, o.ori);
        break;
      }
      case 10: case 11: o.kind = d.kind == 10 ? IOP_ORBIT : IOP_SORBIT; o.flags = d.flags & 1; for (int k = 0; k < 3; k++) o.a[k] = d.v[k]; o.a[3] = d.v[3]; o.a[4] = d.v[4]; o.a[5] = dt; break;
      case 9: { double t = time; if ((d.flags & 1) && t > d.times.back()) t = std::fmod(t, d.times.back()); o.kind = IOP_ANGAXIS; o.m[0] = (int)d.v[0]; linInterp(t, d.times, d.vals, 1, d.ipos, o.a); break; }
      default: break;
    }
  }
  // walls: all DOFs blocked, so only the imposition changes their velocity: the host follows it (Impose::velocity at the end of the node's step, Leapfrog.cpp:248)
  ImposeDev I; I.slot = &H; I.setStart = E->impSetStart.data(); I.setOps = E->impSetOps.data();
  for (size_t w = 0; w < E->wallImpSet.size() && w < 32; w++) {
    if (E->wallImpSet[w] < 0) continue;
    d3 v = mk3(E->impWallVel[3 * w], E->impWallVel[3 * w + 1], E->impWallVel[3 * w + 2]), dummy = mk3(0, 0, 0);
    impose_velocity(I, E->wallImpSet[w], v, dummy, mk3(0, 0, 0));
    E->impWallVel[3 * w] = v.x; E->impWallVel[3 * w + 1] = v.y; E->impWallVel[3 * w + 2] = v.z;
    H.wallHas[w] = 1; H.wallVel[3 * w] = v.x; H.wallVel[3 * w + 1] = v.y; H.wallVel[3 * w + 2] = v.z;
  }
  CK(cudaMemsetAsync(E->dImpRing.p[slot].acc, 0, sizeof(H.acc), E->stream));
  CK(cudaMemcpyAsync(E->dImpRing.p + slot, &H, offsetof(ImpSlot, acc), cudaMemcpyHostToDevice, E->stream));
}

// host prologue (Scene::doOneStep up to the engines, Scene.cpp:414-419): cell integration, per-step uniforms, master nodes of kinematic facet groups
StepU stepPrologue(wdem_engine* E, int k) {
  const wdem_config& c = E->cfg;
  if (c.periodic) E->cell.integrateAndUpdate(E->dt);
  StepU u = makeUniforms(E);
  u.gIn = E->dCnt->gflag + k; u.gOut = E->dCnt->gflag + k + 1; u.impSlot = k;
  if (!E->imposes.empty()) imposePrepareStep(E, k);
  if (u.energy) {   // energy->resetResettables() (Scene.cpp:416): "kinetic" and "elast" restart from zero every step
    CK(cudaMemsetAsync(u.energy + EN_KINETIC, 0, sizeof(double), E->stream)); CK(cudaMemsetAsync(u.energy + EN_ELAST, 0, sizeof(double), E->stream));
  }
  if (E->groups.n > 0 && !E->hFacets.empty()) {
    if (c.periodic && (int)E->hFacets.size() > WD_MAX_PERI_FACETS) throw EngErr{"more than " + std::to_string(WD_MAX_PERI_FACETS) + " facets in a periodic cell are not supported by the GPU engine", WDEM_ERR_UNSUPPORTED};
    // master nodes: all DOFs blocked => Leapfrog's fixed-particle branch with the prescribed velocities: leapfrogTranslate + leapfrogSphericalRotate
    // (Leapfrog.cpp:195-199,275-297) on the host (a handful of nodes; libm sin/cos), members on the device (ClumpData::applyToMembers, Clump.cpp:293-307)
    for (int gi = 0; gi < E->groups.n; gi++) {
      GroupU& G = E->groups.g[gi]; const double dt = E->dt;
      if (c.periodic) {      // a fully blocked node in a deforming cell gets the cell's velocity / position corrections (Leapfrog.cpp:206-215,258-273), no acceleration
        if (u.homoDeform == 4) G.vel = mulMV(u.ImLL4hInv, mulMV(u.LmL, G.pos) + mulMV(u.IpLL4h, G.vel));
        else if (u.homoDeform == 2 || u.homoDeform == 3) {
          if (u.homoDeform == 3) G.vel = G.vel + mulMV(u.midGradV, G.vel - (dt / 2.) * mk3(0, 0, 0)) * dt;
          G.vel = G.vel + mulMV(u.dGradV, G.pos);
        } else if (u.homoDeform == 1) G.pos = G.pos + mulMV(u.nextGradV, G.pos) * dt;
      }
      G.pos = G.pos + G.vel * dt;
      d3 axis = G.angVel;
      if (axis.x != 0. || axis.y != 0. || axis.z != 0.) {
        const double angle = nrm(axis); axis = axis / angle;
        const double ha = 0.5 * (angle * dt), sn = std::sin(ha), qw = std::cos(ha), qx = sn * axis.x, qy = sn * axis.y, qz = sn * axis.z;
        const q4 o = G.ori;
        G.ori.w = qw * o.w - qx * o.x - qy * o.y - qz * o.z; G.ori.x = qw * o.x + qx * o.w + qy * o.z - qz * o.y;
        G.ori.y = qw * o.y + qy * o.w + qz * o.x - qx * o.z; G.ori.z = qw * o.z + qz * o.w + qx * o.y - qy * o.x;
      }
      const double nn = std::sqrt(((G.ori.x * G.ori.x + G.ori.y * G.ori.y) + G.ori.z * G.ori.z) + G.ori.w * G.ori.w);
      G.ori.w /= nn; G.ori.x /= nn; G.ori.y /= nn; G.ori.z /= nn;
      if (gi < (int)E->groupImpSet.size() && E->groupImpSet[gi] >= 0) {      // Impose::velocity before applyToMembers (Leapfrog.cpp:248-252)
        ImposeDev I; I.slot = &E->hImpRing[k]; I.setStart = E->impSetStart.data(); I.setOps = E->impSetOps.data();
        impose_velocity(I, E->groupImpSet[gi], G.vel, G.angVel, G.pos, &G.ori);
      }
    }
  }
  return u;
}

// Leapfrog (+ boundary nodes) and, in decomposed runs, the exchange of (dirty, error) and of the halo payload
void enqueueIntegrate(wdem_engine* E, const StepU& u, cudaEvent_t* ev) {
  const wdem_config& c = E->cfg; cudaStream_t st = E->stream; const int n = E->n;
  if (ev) CK(cudaEventRecord(ev[0], st));
  if (n > 0 && E->e2e.active) leapfrogPipelined(E, u);
  else if (n > 0) {
    if (E->impSpheres) {
      if (E->comm && E->impRead) throw EngErr{"Impose: the read-force kinds (VelocityAndReadForce, ReadForce) are not available in decomposed runs", WDEM_ERR_UNSUPPORTED};
      if (u.energy) k_leapfrog_imp<true><<<nblk(n, 256), 256, 0, st>>>(E->S(), n, u, E->dCnt, E->impDev(u.impSlot));
      else k_leapfrog_imp<false><<<nblk(n, 256), 256, 0, st>>>(E->S(), n, u, E->dCnt, E->impDev(u.impSlot));
    }
    else if (u.energy) k_leapfrog<true><<<nblk(n, 256), 256, 0, st>>>(E->S(), n, u, E->dCnt);
    else k_leapfrog<false><<<nblk(n, 256), 256, 0, st>>>(E->S(), n, u, E->dCnt);
    E->launches++;
  }
  if (!E->imposes.empty() && E->e2e.active) throw EngErr{"Impose is not available with wdem_step_e2e", WDEM_ERR_UNSUPPORTED};
  if (E->nClumpedUploaded > 0) throw EngErr{std::to_string(E->nClumpedUploaded) + " spheres carry the CLUMPED flag but no clump was uploaded (wdem_upload_clumps)", WDEM_ERR_INVALID};
  if (E->clumps.n > 0) {
    if (E->comm || E->e2e.active || E->trackEnergy) throw EngErr{"clumps of spheres: decomposed runs, wdem_step_e2e and Scene.trackEnergy are not supported by the GPU engine", WDEM_ERR_UNSUPPORTED};
    k_clumps_step<<<nblk(E->clumps.n, 128), 128, 0, st>>>(E->clumps.view(), E->S(), E->id2slot.p, u, E->dCnt, E->impDev(u.impSlot)); E->launches++;
  }
  if ((E->groups.n > 0 || E->impFacetNodes) && !E->hFacets.empty()) {
    k_facets_step<<<nblk((long long)E->hFacets.size(), 128), 128, 0, st>>>(E->dFacets.p, (int)E->hFacets.size(), E->groups, u, E->dCnt, 0, E->impDev(u.impSlot)); E->launches++;
  }
  if (E->anyMovingWall) {
    k_walls_step<<<nblk((long long)E->hWalls.size(), 32), 32, 0, st>>>(E->dWalls.p, (int)E->hWalls.size(), u, E->dCnt, 0, E->wallImpSet.empty() ? nullptr : E->dImpRing.p + u.impSlot); E->launches++;
  }
  if (ev) CK(cudaEventRecord(ev[1], st));
  // (dirty, error): every subdomain rebuilds / fails in the same step; over peer memory together with the halo payload, else NCCL
  if (E->comm && E->dom.peer.active) peerExchange(E, u.gOut);
  else if (E->comm) E->comm->allreduceMaxInt(&E->dCnt->error, WD_KMAX + 2, st);
}
// ghosts refreshed, then ContactLoop + force reduction, all void on the device if this step's flag is up
void enqueueContacts(wdem_engine* E, const StepU& u, cudaEvent_t* ev) {
  if (E->comm && E->dom.peer.active) peerUnpack(E);
  else if (E->comm) haloExchange(E);
  launchContactsAndReduce(E, u, u.gOut, ev);
}
// Scene::doOneStep's tail (Scene.cpp:428-432)
void stepEpilogue(wdem_engine* E) {
  if (E->cfg.periodic) E->cell.setNextGradV();
  E->step++; E->time += E->dt; E->stepsSinceFullRun++;
  if (!std::isnan(E->nextDt)) { E->dt = E->nextDt; E->nextDt = NAN; }
}
// DynDt (PeriodicEngine::isActivated, Engine.cpp:170-195 with stepModulo/initRun defaults; DynDt::run DynDt.cpp:90-113)
bool dynDtDue(const wdem_engine* E, long long step) {
  const int P = E->cfg.dynDtPeriod;
  if (P <= 0) return false;
  const long long last = E->dynStepLast < 0 ? E->step : E->dynStepLast;     // the first check initialises the counter with the current step
  return (step - last >= P) || (step % P == 0);
}
void dynDtIfDue(wdem_engine* E) {
  const wdem_config& c = E->cfg;
  if (c.dynDtPeriod <= 0) return;
  if (E->dynStepLast < 0) E->dynStepLast = E->step;
  if ((E->step - E->dynStepLast >= c.dynDtPeriod) || (E->step % c.dynDtPeriod == 0)) {
    E->dynStepPrev = E->dynStepLast; E->dynStepLast = E->step;
    double crDt = dynDtCritGlobal(E);
    if (!std::isinf(crDt)) {
      int nSteps = (int)(E->step - E->dynStepPrev);
      double maxNewDt = E->dt * pown(1. + c.maxRelInc, (unsigned)nSteps);
      double nDt = std::min(crDt * c.dtSafety, maxNewDt);
      if (nDt != E->dt) E->nextDt = nDt;
    }
  }
}

// `nsteps` steps of Scene::doOneStep.  Steps are enqueued in batches of K without a host wait in between; on the device every kernel of a batch is guarded by the
// Verlet flags of the steps before it (Counters::gflag), so the first step that needs a collider full run stops everything behind it.  The host then rolls its own
// per-step state (cell, facet-group masters, step counter) back to that step, does the full run, repeats the contact pass of that step and carries on.  K follows
// the observed number of steps between full runs (1 while every step triggers one, e.g. in free fall); it is 1 for the end-to-end path, with Scene.trackEnergy
// (resettable energies are zeroed per step) and always ends at a step in which DynDt is due.
void runSteps(wdem_engine* E, int64_t nsteps, wdem_stats* stats) {
  cudaStream_t st = E->stream;
  std::vector<StepSnap> snaps;
  int64_t done = 0;
  static const bool trace = getenv("WDEM_STEP_TRACE") != nullptr;      // host-side timers of the stepping loop (diagnostics)
  double tEnq = 0, tWait = 0, tReb = 0; long nBatch = 0, nVoid = 0, nReb = 0; auto clk = [] { return std::chrono::steady_clock::now(); };
  auto msSince = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(clk() - a).count(); };
  while (done < nsteps) {
    auto t0 = clk();
    ensureVerletAndDt(E);
    const bool full0 = E->conDirty || E->verletDist == 0;     // a full run is certain: do not enqueue the contact pass speculatively
    if (E->cfg.dynDtPeriod > 0 && E->dynStepLast < 0) E->dynStepLast = E->step;      // PeriodicEngine's first activation check
    const long long since0 = E->stepsSinceFullRun;
    int K = 1;
    if (!full0 && !E->e2e.active && !E->trackEnergy && E->n > 0) {
      K = (int)std::min<int64_t>(std::min(E->batchK, E->batchMax), nsteps - done);
      for (int k = 0; k < K; k++) if (dynDtDue(E, E->step + k)) { K = k + 1; break; }
    }
    snaps.resize(K);
    for (int k = 0; k < K; k++) {
      const StepU u = stepPrologue(E, k);
      snaps[k] = StepSnap{E->cell, E->groups, E->step, E->time, u, E->impWallVel};
      cudaEvent_t* ev = nullptr;
      if (E->profiling) { const int set = (int)(E->step % wdem_engine::NEVSET); harvestPhases(E, set); ev = E->evr[set]; }
      enqueueIntegrate(E, u, ev);
      if (E->e2e.active && E->e2e.defer) {      // the verdict of this step's Verlet test travels before the contact pass is enqueued: the host does not wait for the contacts
        CK(cudaMemcpyAsync(E->hCnt, E->dCnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(E->evFlag, st));
      }
      if (!full0) enqueueContacts(E, u, ev);
      if (ev) { CK(cudaEventRecord(ev[5], st)); E->evPending[E->step % wdem_engine::NEVSET] = true; }
      if (k < K - 1) stepEpilogue(E);
    }
    if (!(E->e2e.active && E->e2e.defer)) {
      CK(cudaMemcpyAsync(E->hCnt, E->dCnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(E->evFlag, st));
    }
    if (trace) { tEnq += msSince(t0); t0 = clk(); }
    CK(cudaEventSynchronize(E->evFlag));
    if (trace) { tWait += msSince(t0); t0 = clk(); nBatch++; }
    checkDeviceError(E);
    int kd = full0 ? 0 : -1;
    if (!full0) for (int k = 0; k < K; k++) if (E->hCnt->gflag[k + 1]) { kd = k; break; }
    int executed = K;
    if (kd >= 0) {
      // back to step kd, after its prologue and its Leapfrog
      E->cell = snaps[kd].cell; E->groups = snaps[kd].groups; E->step = snaps[kd].step; E->time = snaps[kd].time; E->impWallVel = snaps[kd].wallVel;
      const long long interval = since0 + kd + 1;
      rebuild(E);
      E->conDirty = false; E->stepsSinceFullRun = -1;      // the epilogue below counts this step
      E->batchK = (int)std::max<long long>(1, std::min<long long>(E->batchMax, interval / 2));
      CK(cudaMemsetAsync(E->dCnt->gflag, 0, sizeof(int) * (WD_KMAX + 1), st));
      if (stats) stats->n_rebuilds++;
      StepU u = snaps[kd].u; u.gIn = E->dCnt->gflag; u.gOut = E->dCnt->gflag;      // slot 0 is always 0: the repeated pass is not speculative
      cudaEvent_t* ev = E->profiling ? E->evr[E->step % wdem_engine::NEVSET] : nullptr;
      launchContactsAndReduce(E, u, u.gOut, ev);
      if (ev) CK(cudaEventRecord(ev[5], st));
      executed = kd + 1;
      if (trace) { CK(cudaStreamSynchronize(st)); tReb += msSince(t0); nReb++; nVoid += K - executed; }
    } else if (E->stepsSinceFullRun + K > 2 * (long long)E->batchK) E->batchK = std::min(E->batchMax, E->batchK + 1);      // long quiet stretch: larger batches
    if (!E->imposes.empty()) {      // VelocityAndReadForce.dist: += vel*dt in every executed step in which some node carried it (Impose.cpp:166-173)
      for (auto& d : E->imposes) if (d.kind == 6 && d.nNodes > 0) for (int k = 0; k < executed; k++) d.dist += d.v[3] * snaps[k].u.dt;
      for (auto& d : E->imposes) if ((d.kind == 10 || d.kind == 11) && d.nNodes > 0) for (int k = 0; k < executed; k++) d.dist += d.v[3] * snaps[k].u.dt;      // .angle += omega*dt (Impose.cpp:36,47)
      E->lastImpSlot = executed - 1;
    }
    gradVWorkStep(E);
    dynDtIfDue(E);
    stepEpilogue(E);
    done += executed;
  }
  if (trace && nsteps > 1) fprintf(stderr, "runSteps trace: %lld steps in %ld batches (batchK now %d), %ld full runs, %ld void steps | host ms: enqueue %.2f, wait %.2f, full runs + repeated pass %.2f\n",
                                   (long long)nsteps, nBatch, E->batchK, nReb, nVoid, tEnq, tWait, tReb);
}

template <class F>
int guarded(wdem_handle h, F&& f) {
  if (!h) return WDEM_ERR_INVALID;
  try {
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) { h->err = "cudaSetDevice failed (no usable CUDA device; this engine has no CPU fallback)"; return WDEM_ERR_CUDA; }
    g_allocStream = h->stream;
    f();
    return WDEM_OK;
  } catch (CudaErr& e) { h->err = e.msg; if (h->comm) h->comm->abort(); return e.code; }
  catch (EngErr& e) { h->err = e.msg; if (h->comm) h->comm->abort(); return e.code; }
  catch (CommErr& e) { h->err = e.msg; if (h->comm) h->comm->abort(); return WDEM_ERR_RUNTIME; }
  catch (std::exception& e) { h->err = e.what(); if (h->comm) h->comm->abort(); return WDEM_ERR_RUNTIME; }
}

void countContacts(wdem_engine* E, long long& nReal, long long& nAlive) {
  nReal = nAlive = 0;
  if (E->nC == 0) return;
  E->cntScratch.ensure(2, E->bytes);
  CK(cudaMemsetAsync(E->cntScratch.p, 0, 2 * sizeof(unsigned long long), E->stream));
  if (E->comm) k_count_contacts_owned<<<nblk(E->nC, 256), 256, 0, E->stream>>>(E->con[E->curC].view(), E->nC, E->S().flags, E->cntScratch.p);
  else k_count_flags<<<nblk(E->nC, 256), 256, 0, E->stream>>>(E->con[E->curC].flags.p, E->nC, E->cntScratch.p);
  unsigned long long h[2];
  CK(cudaMemcpyAsync(h, E->cntScratch.p, sizeof(h), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  nReal = (long long)h[0]; nAlive = (long long)h[1];
}

void ensureUp2Slot(wdem_engine* E) {
  if (E->up2slotValid || E->n == 0) return;
  E->up2slot.ensure(E->n, E->bytes);
  k_up2slot<<<nblk(E->n, 256), 256, 0, E->stream>>>(E->S().up, E->n, E->up2slot.p);
  E->up2slotValid = true;
}

template <class T>
std::vector<T> d2h(const T* p, size_t n, cudaStream_t st) {
  std::vector<T> v(n);
  if (n) { CK(cudaMemcpyAsync(v.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
  return v;
}
}  // namespace

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

int wdem_default_config(wdem_config* c) {
  if (!c) return WDEM_ERR_INVALID;
  memset(c, 0, sizeof(*c));
  c->dt = NAN; c->dtSafety = 0.9; c->distFactor = -1.; c->damping = 0.2; c->verletDist = -0.05; c->poisson = 0.2; c->en = NAN; c->maxRelInc = 1e-4;
  c->loneMask = 2; c->reset = 1; c->iniLensTouch = 1; c->trsfRenorm = 100; c->applyForces = 1; c->law = WDEM_LAW_IDEALELPL; c->noAttraction = 1; c->dynDtPeriod = 100;
  return WDEM_OK;
}

const char* wdem_last_error(wdem_handle h) { return h ? h->err.c_str() : g_createError.c_str(); }

int wdem_create(const wdem_config* cfg, wdem_handle* out) {
  if (!cfg || !out) return WDEM_ERR_INVALID;
  *out = nullptr;
  wdem_engine* E = nullptr;
  try {
    validateConfig(*cfg);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw CudaErr{"no CUDA device available: the B200 DEM engine has no CPU fallback", WDEM_ERR_CUDA};
    if (cfg->device < 0 || cfg->device >= ndev) throw EngErr{"invalid device ordinal", WDEM_ERR_INVALID};
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) throw CudaErr{std::string("device '") + prop.name + "' is not sm_100 (this library carries sm_100a code only)", WDEM_ERR_CUDA};
    E = new wdem_engine(); E->cfg = *cfg; E->dt = cfg->dt; E->numSMs = prop.multiProcessorCount;
    if (const char* g = getenv("WDEM_E2E_CHUNKS")) E->nChunk = std::max(1, std::min((int)wdem_engine::NCHUNK, atoi(g)));
    if (const char* g = getenv("WDEM_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));   // experiment knob: 32/64/128
    CK(cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&E->stream2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&E->sH2D, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&E->sD2H, cudaStreamNonBlocking));
    const unsigned evf = getenv("WDEM_E2E_TRACE") ? cudaEventDefault : cudaEventDisableTiming;
    for (auto& e : E->evUp) CK(cudaEventCreateWithFlags(&e, evf));
    for (auto& e : E->evLf) CK(cudaEventCreateWithFlags(&e, evf));
    CK(cudaEventCreateWithFlags(&E->evEk, evf));
    CK(cudaEventCreate(&E->ev0)); CK(cudaEventCreate(&E->ev1));
    CK(cudaEventCreateWithFlags(&E->evFlag, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&E->evFork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&E->evJoin, cudaEventDisableTiming));
    for (auto& set : E->evr) for (auto& e : set) CK(cudaEventCreate(&e));
    {   // freed blocks stay in the device's pool instead of going back to the driver at the next synchronisation
      cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, cfg->device) == cudaSuccess) { unsigned long long keepAll = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keepAll); } else cudaGetLastError();
    }
    CK(cudaMalloc(&E->dCnt, sizeof(Counters))); CK(cudaMemset(E->dCnt, 0, sizeof(Counters)));
    CK(cudaMallocHost(&E->hCnt, sizeof(Counters))); memset(E->hCnt, 0, sizeof(Counters));
    CK(cudaMallocHost(&E->hScratch, 256 * sizeof(int))); memset(E->hScratch, 0, 256 * sizeof(int));
    CK(cudaMallocHost(&E->hBounds, 16 * sizeof(double))); memset(E->hBounds, 0, 16 * sizeof(double));
    CK(cudaEventCreateWithFlags(&E->evOrph, cudaEventDisableTiming));
    if (const char* g = getenv("WDEM_BATCH")) E->batchMax = std::max(1, std::min(WD_KMAX, atoi(g)));
    *out = E;
    return WDEM_OK;
  } catch (CudaErr& e) { g_createError = e.msg; delete E; return e.code; }
  catch (EngErr& e) { g_createError = e.msg; delete E; return e.code; }
}

int wdem_destroy(wdem_handle h) {
  if (!h) return WDEM_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& s : h->sph) s.release(); for (auto& c : h->con) c.release(); for (auto& r : h->csr) r.release(); h->clumps.release();
  h->dImpSetStart.release(); h->dImpSetOps.release(); h->dImpOfId.release(); h->dImpRing.release(); if (h->hImpRing) cudaFreeHost(h->hImpRing);
  h->facetBoxes.release(); h->ownedSlot.release(); h->ownedIds.release();
  h->id2slot.release(); h->up2slot.release(); h->dMats.release(); h->dWalls.release(); h->dFacets.release(); h->fgCellStart.release(); h->fgCellFacet.release(); h->fgCursor.release(); h->keys.release(); h->keysAlt.release(); h->idx.release(); h->idxAlt.release();
  h->orphCnt.release(); h->rowCount.release(); h->srcOld.release(); h->candBuf.release(); h->wallBits.release(); h->ioBuf.release(); h->cellStart.release(); h->counts.release(); h->offsets.release(); h->cubTemp.release(); h->blk.release(); h->ckeys.release(); h->ckeysAlt.release(); h->cntScratch.release(); h->cvals.release(); h->cvalsAlt.release(); h->realFlag.release(); h->realPos.release(); for (auto& b : h->tmpCd) b.release();
  peerTeardown(h, false);
  h->dom.release(); delete h->comm; h->comm = nullptr;
  if (h->dCnt) cudaFree(h->dCnt); if (h->hCnt) cudaFreeHost(h->hCnt); if (h->hScratch) cudaFreeHost(h->hScratch); if (h->hBounds) cudaFreeHost(h->hBounds); if (h->evOrph) cudaEventDestroy(h->evOrph);
  if (h->ev0) cudaEventDestroy(h->ev0); if (h->ev1) cudaEventDestroy(h->ev1); for (auto& set : h->evr) for (auto& e : set) if (e) cudaEventDestroy(e);
  if (h->evFlag) cudaEventDestroy(h->evFlag); if (h->evFork) cudaEventDestroy(h->evFork); if (h->evJoin) cudaEventDestroy(h->evJoin);
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  for (cudaStream_t cs : {h->sH2D, h->sD2H}) if (cs) { cudaStreamSynchronize(cs); cudaStreamDestroy(cs); }
  for (auto& e : h->evUp) if (e) cudaEventDestroy(e); for (auto& e : h->evLf) if (e) cudaEventDestroy(e); if (h->evEk) cudaEventDestroy(h->evEk);
  h->boundsBuf.release(); h->wallAcc.release(); h->wallPartial.release(); h->dEnergy.release(); h->ioIn.release(); h->ioOut.release(); if (h->hEk) cudaFreeHost(h->hEk);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return WDEM_OK;
}

int wdem_set_config(wdem_handle h, const wdem_config* cfg) {
  return guarded(h, [&] {
    if (!cfg) throw EngErr{"null config", WDEM_ERR_INVALID};
    validateConfig(*cfg);
    if (cfg->law != h->cfg.law && h->nC > 0) throw EngErr{"cannot change the contact law while contacts exist", WDEM_ERR_UNSUPPORTED};
    int dev = h->cfg.device; bool newDt = !(cfg->dt == h->cfg.dt || (std::isnan(cfg->dt) && std::isnan(h->cfg.dt)));
    h->cfg = *cfg; h->cfg.device = dev;
    if (newDt) h->dt = cfg->dt;
  });
}
int wdem_set_profiling(wdem_handle h, int on) { return guarded(h, [&] { h->profiling = on != 0; }); }

int wdem_upload_materials(wdem_handle h, int n, const int32_t* kind, const double* density, const double* young, const double* ktDivKn, const double* tanPhi,
                          const double* surfEnergy, const double* alpha) {
  return guarded(h, [&] {
    if (n <= 0 || !kind || !density || !young || !ktDivKn || !tanPhi) throw EngErr{"wdem_upload_materials: bad arguments", WDEM_ERR_INVALID};
    h->hMats.resize(n);
    for (int i = 0; i < n; i++) { Material m; m.kind = kind[i]; m.pad = 0; m.density = density[i]; m.young = young[i]; m.ktDivKn = ktDivKn[i]; m.tanPhi = tanPhi[i]; m.surfEnergy = surfEnergy ? surfEnergy[i] : 0.; m.alpha = alpha ? alpha[i] : 0.; h->hMats[i] = m; }
    h->dMats.ensure(n, h->bytes);
    CK(cudaMemcpyAsync(h->dMats.p, h->hMats.data(), n * sizeof(Material), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  });
}

int wdem_upload_spheres(wdem_handle h, int64_t n, const int32_t* ids, const double* pos, const double* ori, const double* vel, const double* angVel, const double* mass,
                        const double* inertia, const double* radius, const uint32_t* flags, const uint32_t* mask, const int32_t* matId) {
  return guarded(h, [&] {
    if (n < 0 || n > (1ll << 30) || (n > 0 && (!pos || !mass || !inertia || !radius))) throw EngErr{"wdem_upload_spheres: bad arguments", WDEM_ERR_INVALID};
    if (h->hMats.empty()) throw EngErr{"wdem_upload_spheres: upload materials first", WDEM_ERR_INVALID};
    // DemField::selfTest-like checks (Particle.cpp:566-625)
    std::vector<double4> posr(n), velm(n), angi(n), q(n); std::vector<uint32_t> fl(n), mk(n); std::vector<int32_t> mt(n), id(n), up(n);
    int maxId = h->maxId; double minR = INFINITY, pw = INFINITY;
    for (auto& w : h->hWalls) maxId = std::max(maxId, w.id);
    for (auto& f : h->hFacets) maxId = std::max(maxId, f.id);
    for (int64_t i = 0; i < n; i++) {
      id[i] = ids ? ids[i] : (int32_t)i; up[i] = (int32_t)i;
      if (id[i] < 0) throw EngErr{"negative particle id", WDEM_ERR_INVALID};
      maxId = std::max(maxId, id[i]);
      fl[i] = flags ? flags[i] : 0u; mk[i] = mask ? mask[i] : 5u; mt[i] = matId ? matId[i] : 0;
      if (mt[i] < 0 || mt[i] >= (int)h->hMats.size()) throw EngErr{"sphere material index out of range", WDEM_ERR_INVALID};
      if (fl[i] & 128u) throw EngErr{"sphere #" + std::to_string(id[i]) + ": DemData.flags has the CLUMP bit (a master node is not a particle: wdem_upload_clumps)", WDEM_ERR_INVALID};
      if (!(radius[i] > 0)) throw EngErr{"sphere #" + std::to_string(id[i]) + ": radius must be positive", WDEM_ERR_INVALID};
      if ((fl[i] & 7u) != 7u && !(mass[i] > 0)) throw EngErr{"DemField.selfTest: particle #" + std::to_string(id[i]) + " has non-positive mass with free translational DOFs", WDEM_ERR_INVALID};
      if ((fl[i] & 56u) != 56u && !(inertia[i] > 0)) throw EngErr{"DemField.selfTest: particle #" + std::to_string(id[i]) + " has non-positive inertia with free rotational DOFs", WDEM_ERR_INVALID};
      posr[i] = make_double4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], radius[i]);
      velm[i] = make_double4(vel ? vel[3 * i] : 0, vel ? vel[3 * i + 1] : 0, vel ? vel[3 * i + 2] : 0, mass[i]);
      angi[i] = make_double4(angVel ? angVel[3 * i] : 0, angVel ? angVel[3 * i + 1] : 0, angVel ? angVel[3 * i + 2] : 0, inertia[i]);
      q[i] = ori ? make_double4(ori[4 * i], ori[4 * i + 1], ori[4 * i + 2], ori[4 * i + 3]) : make_double4(0, 0, 0, 1);
      minR = std::min(minR, radius[i]);
      const Material& m = h->hMats[mt[i]];
      pw = std::min(pw, radius[i] / (1. * std::sqrt(m.young / m.density)));  // DemFuncs::pWaveDt (Funcs.cpp:151-175)
    }
    {   // Particle.id is the key of id2slot and of the contact history: duplicates would silently corrupt both
      std::vector<int32_t> srt(id); std::sort(srt.begin(), srt.end());
      auto dup = std::adjacent_find(srt.begin(), srt.end());
      if (dup != srt.end()) throw EngErr{"wdem_upload_spheres: duplicate particle id " + std::to_string(*dup), WDEM_ERR_INVALID};
      for (auto& w : h->hWalls) if (std::binary_search(srt.begin(), srt.end(), w.id)) throw EngErr{"wdem_upload_spheres: particle id " + std::to_string(w.id) + " is already used by a wall", WDEM_ERR_INVALID};
      for (auto& f : h->hFacets) if (std::binary_search(srt.begin(), srt.end(), f.id)) throw EngErr{"wdem_upload_spheres: particle id " + std::to_string(f.id) + " is already used by a facet", WDEM_ERR_INVALID};
    }
    h->n = (int)n; h->minRadius = minR; h->pwaveDt = pw; h->maxId = maxId;
    h->hIds.assign(id.begin(), id.end()); h->pendingRows.clear(); h->compactPending = false;
    h->cur = 0;
    SphereStore& s = h->sph[0];
    s.ensure(std::max<int64_t>(n, 1), h->bytes);
    cudaStream_t st = h->stream;
    auto H2D = [&](void* d, const void* src, size_t b) { if (b) CK(cudaMemcpyAsync(d, src, b, cudaMemcpyHostToDevice, st)); };
    H2D(s.posr.p, posr.data(), n * sizeof(double4)); H2D(s.velm.p, velm.data(), n * sizeof(double4)); H2D(s.angi.p, angi.data(), n * sizeof(double4)); H2D(s.ori.p, q.data(), n * sizeof(double4));
    for (auto& b : s.f) if (n) CK(cudaMemsetAsync(b.p, 0, n * sizeof(double), st));
    for (auto& b : s.l) if (n) CK(cudaMemsetAsync(b.p, 0, n * sizeof(double), st));
    H2D(s.flags.p, fl.data(), n * 4); H2D(s.mask.p, mk.data(), n * 4); H2D(s.mat.p, mt.data(), n * 4); H2D(s.id.p, id.data(), n * 4); H2D(s.up.p, up.data(), n * 4);
    const int idCap = std::max(maxId, h->dom.globalMaxId);
    h->id2slot.ensure((size_t)idCap + 2, h->bytes);
    CK(cudaMemsetAsync(h->id2slot.p, 0xFF, ((size_t)idCap + 2) * sizeof(int), st));
    h->dom.nOwn = (int)n; h->dom.globalsValid = false; h->dom.rmaxGlobal = 0; for (int64_t i = 0; i < n; i++) h->dom.rmaxGlobal = std::max(h->dom.rmaxGlobal, radius[i]);
    h->dom.nSendTot = h->dom.nRecvTot = 0; h->dom.nImm = 0;
    CK(cudaStreamSynchronize(st));
    h->conDirty = true; h->nC = h->nSS = h->nSSa = h->nWS = h->nFS = h->nOrph = 0; h->csr[0].valid = h->csr[1].valid = false; h->up2slotValid = false;
    h->verletDist = NAN; h->fgridValid = false; h->boundsValid = false; h->batchK = 1;
    h->imposes.clear(); h->impSetStart.clear(); h->impSetOps.clear(); h->wallImpSet.clear(); h->groupImpSet.clear(); h->impSpheres = false; h->impRead = false; h->impFacetNodes = false;
    h->clumps.release(); h->nClumpedUploaded = 0; for (int64_t i = 0; i < n; i++) if (fl[i] & 64u) h->nClumpedUploaded++;
  });
}

// ClumpData of the scene's clumps of spheres (Clump.hpp).  relPos/relOri NULL: derived from the members' uploaded state as ClumpData::makeClump does
// (Clump.cpp:252-253).  angMom NULL or NaN: initialised by the first step (Leapfrog.cpp:303-305).
int wdem_upload_clumps(wdem_handle h, int32_t nClumps, const int32_t* memberStart, const int32_t* memberIds, const double* pos, const double* ori, const double* vel, const double* angVel,
                       const double* angMom, const double* mass, const double* inertia, const uint32_t* flags, const double* relPos, const double* relOri) {
  return guarded(h, [&] {
    if (nClumps < 0 || (nClumps > 0 && (!memberStart || !memberIds || !pos || !ori || !mass || !inertia))) throw EngErr{"wdem_upload_clumps: bad arguments", WDEM_ERR_INVALID};
    if (!h->conDirty || h->cur != 0 || h->nC != 0) throw EngErr{"wdem_upload_clumps: call right after wdem_upload_spheres", WDEM_ERR_INVALID};
    if (h->comm) throw EngErr{"wdem_upload_clumps: not available in decomposed runs", WDEM_ERR_UNSUPPORTED};
    ClumpStore& K = h->clumps; K.release();
    if (nClumps == 0) return;
    const int n = h->n, nMem = memberStart[nClumps];
    if (memberStart[0] != 0 || nMem < 0) throw EngErr{"wdem_upload_clumps: memberStart must start at 0 and ascend", WDEM_ERR_INVALID};
    cudaStream_t st = h->stream;
    // the members' uploaded state (upload order: no collider run has re-ordered the slots yet)
    std::vector<double4> sp(n), so(n); std::vector<uint32_t> sfl(n);
    SphereStore& s = h->sph[0];
    CK(cudaMemcpyAsync(sp.data(), s.posr.p, n * sizeof(double4), cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(so.data(), s.ori.p, n * sizeof(double4), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sfl.data(), s.flags.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    std::unordered_map<int32_t, int> upOf; upOf.reserve(n * 2); for (int i = 0; i < n; i++) upOf[h->hIds[i]] = i;
    std::vector<int32_t> clumpOfId((size_t)std::max(h->maxId, 0) + 2, -1);
    std::vector<double4> cp(nClumps), co(nClumps), cv(nClumps), cw(nClumps), cl(nClumps), cf(nClumps), ct(nClumps), cI(nClumps), rp(nMem), rq(nMem); std::vector<uint32_t> cfl(nClumps); std::vector<int> ms(memberStart, memberStart + nClumps + 1);
    int nClumped = 0;
    for (int c = 0; c < nClumps; c++) {
      if (ms[c + 1] <= ms[c]) throw EngErr{"ClumpData::makeClump: 0 nodes.", WDEM_ERR_INVALID};
      cfl[c] = flags ? flags[c] : 0u;
      if ((cfl[c] & 7u) != 7u && !(mass[c] > 0)) throw EngErr{"DemField.selfTest: clump " + std::to_string(c) + " has non-positive mass with free translational DOFs", WDEM_ERR_INVALID};
      cp[c] = make_double4(pos[3 * c], pos[3 * c + 1], pos[3 * c + 2], mass[c]); co[c] = make_double4(ori[4 * c], ori[4 * c + 1], ori[4 * c + 2], ori[4 * c + 3]);
      cv[c] = vel ? make_double4(vel[3 * c], vel[3 * c + 1], vel[3 * c + 2], 0) : make_double4(0, 0, 0, 0); cw[c] = angVel ? make_double4(angVel[3 * c], angVel[3 * c + 1], angVel[3 * c + 2], 0) : make_double4(0, 0, 0, 0);
      cl[c] = angMom ? make_double4(angMom[3 * c], angMom[3 * c + 1], angMom[3 * c + 2], 0) : make_double4(NAN, NAN, NAN, 0);
      cf[c] = ct[c] = make_double4(0, 0, 0, 0); cI[c] = make_double4(inertia[3 * c], inertia[3 * c + 1], inertia[3 * c + 2], 0);
      q4 q; q.x = co[c].x; q.y = co[c].y; q.z = co[c].z; q.w = co[c].w; const q4 qc = quatConj(q);
      for (int m = ms[c]; m < ms[c + 1]; m++) {
        auto it = upOf.find(memberIds[m]);
        if (it == upOf.end()) throw EngErr{"clump: member #" + std::to_string(memberIds[m]) + " is not a sphere", WDEM_ERR_INVALID};
        const int i = it->second;
        if (clumpOfId[memberIds[m]] >= 0) throw EngErr{"Node of particle #" + std::to_string(memberIds[m]) + ": already clumped.", WDEM_ERR_INVALID};
        if (!(sfl[i] & 64u)) throw EngErr{"clump member #" + std::to_string(memberIds[m]) + ": DemData.flags lacks the CLUMPED bit", WDEM_ERR_INVALID};
        clumpOfId[memberIds[m]] = c; nClumped++;
        if (relPos) rp[m] = make_double4(relPos[3 * m], relPos[3 * m + 1], relPos[3 * m + 2], 0);
        else { const d3 r = quatRotate(qc, xyz(sp[i]) - xyz(cp[c])); rp[m] = make_double4(r.x, r.y, r.z, 0); }
        if (relOri) rq[m] = make_double4(relOri[4 * m], relOri[4 * m + 1], relOri[4 * m + 2], relOri[4 * m + 3]);
        else { q4 mo; mo.x = so[i].x; mo.y = so[i].y; mo.z = so[i].z; mo.w = so[i].w; const q4 r = quatMul(qc, mo); rq[m] = make_double4(r.x, r.y, r.z, r.w); }
      }
    }
    if (nClumped != h->nClumpedUploaded) throw EngErr{"wdem_upload_clumps: " + std::to_string(h->nClumpedUploaded) + " spheres carry the CLUMPED flag but " + std::to_string(nClumped) + " are members of the uploaded clumps", WDEM_ERR_INVALID};
    K.n = nClumps; K.nMem = nMem;
    auto up4 = [&](DBuf<double4>& b, const std::vector<double4>& v) { b.ensure(v.size(), h->bytes); CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(double4), cudaMemcpyHostToDevice, st)); };
    up4(K.pos, cp); up4(K.ori, co); up4(K.vel, cv); up4(K.angVel, cw); up4(K.angMom, cl); up4(K.force, cf); up4(K.torque, ct); up4(K.inertia, cI); up4(K.relPos, rp); up4(K.relOri, rq);
    K.flags.ensure(nClumps, h->bytes); CK(cudaMemcpyAsync(K.flags.p, cfl.data(), nClumps * 4, cudaMemcpyHostToDevice, st));
    K.memStart.ensure(nClumps + 1, h->bytes); CK(cudaMemcpyAsync(K.memStart.p, ms.data(), (nClumps + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    K.memId.ensure(nMem, h->bytes); CK(cudaMemcpyAsync(K.memId.p, memberIds, nMem * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    K.clumpOfId.ensure(clumpOfId.size(), h->bytes); CK(cudaMemcpyAsync(K.clumpOfId.p, clumpOfId.data(), clumpOfId.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // Leapfrog precedes the first collider run: the members must be addressable by id from the first step on
    k_id2slot_fill<<<nblk(n, 256), 256, 0, st>>>(s.id.p, n, h->id2slot.p); h->launches++;
    CK(cudaStreamSynchronize(st));
    h->nClumpedUploaded = 0;      // all claimed
  });
}
int wdem_download_clumps(wdem_handle h, double* pos, double* ori, double* vel, double* angVel, double* angMom, double* force, double* torque) {
  return guarded(h, [&] {
    ClumpStore& K = h->clumps; const int n = K.n; if (n == 0) return;
    std::vector<double4> b(n);
    auto get = [&](DBuf<double4>& src, double* dst, int w) {
      if (!dst) return;
      CK(cudaMemcpyAsync(b.data(), src.p, n * sizeof(double4), cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream));
      for (int i = 0; i < n; i++) { dst[w * i] = b[i].x; dst[w * i + 1] = b[i].y; dst[w * i + 2] = b[i].z; if (w == 4) dst[w * i + 3] = b[i].w; }
    };
    get(K.pos, pos, 3); get(K.ori, ori, 4); get(K.vel, vel, 3); get(K.angVel, angVel, 3); get(K.angMom, angMom, 3); get(K.force, force, 3); get(K.torque, torque, 3);
  });
}
int64_t wdem_num_clumps(wdem_handle h) { return h ? h->clumps.n : 0; }

int wdem_upload_imposes(wdem_handle h, int32_t nImp, const wdem_impose* imp, int32_t nSets, const int32_t* setStart, const int32_t* setOps, int64_t nSphereIds,
                        const int32_t* sphereIds, const int32_t* sphereSet, const int32_t* wallSet, const int32_t* groupSet) {
  return guarded(h, [&] {
    if (nImp < 0 || nImp > WD_MAX_IMPOSE) throw EngErr{"wdem_upload_imposes: at most " + std::to_string(WD_MAX_IMPOSE) + " Impose objects", WDEM_ERR_UNSUPPORTED};
    if (nImp > 0 && (!imp || nSets <= 0 || !setStart || !setOps)) throw EngErr{"wdem_upload_imposes: bad arguments", WDEM_ERR_INVALID};
    if (!h->conDirty || h->nC != 0 || h->cur != 0) throw EngErr{"wdem_upload_imposes: call after the particles are uploaded and before the first step", WDEM_ERR_INVALID};
    if (h->cfg.periodic && wallSet) throw EngErr{"wdem_upload_imposes: impositions on walls in a periodic cell are not supported", WDEM_ERR_UNSUPPORTED};
    h->imposes.clear(); h->impSetStart.clear(); h->impSetOps.clear(); h->wallImpSet.clear(); h->groupImpSet.clear(); h->impSpheres = false; h->impRead = false; h->impFacetNodes = false;
    if (nImp == 0) return;
    cudaStream_t st = h->stream;
    for (int i = 0; i < nImp; i++) {
      const wdem_impose& s = imp[i]; wdem_engine::ImposeHost d;
      if (s.kind < 1 || s.kind > 12) throw EngErr{"wdem_upload_imposes: unknown kind " + std::to_string(s.kind), WDEM_ERR_UNSUPPORTED};
      if ((s.kind == 10 || s.kind == 11) && !(s.v[3] == s.v[3])) throw EngErr{"CircularOrbit / StableCircularOrbit: omega must be given", WDEM_ERR_INVALID};
      if (s.kind == 11 && !(s.v[4] == s.v[4])) throw EngErr{"StableCircularOrbit: radius must be given", WDEM_ERR_INVALID};
      d.kind = s.kind; d.flags = s.flags; memcpy(d.v, s.v, sizeof(d.v)); memcpy(d.ori, s.ori, sizeof(d.ori));
      if (s.kind == 12) {
        if (s.nTimes <= 0 || !s.times || !s.vals || !s.vals2) throw EngErr{"InterpolatedMotion: poss, oris, times must have the same (non-zero) length", WDEM_ERR_INVALID};
        d.times.assign(s.times, s.times + s.nTimes); d.vals.assign(s.vals, s.vals + (size_t)3 * s.nTimes); d.vals2.assign(s.vals2, s.vals2 + (size_t)4 * s.nTimes);
        for (int k = 0; k + 1 < s.nTimes; k++) if (!(d.times[k + 1] > d.times[k])) throw EngErr{"InterpolatedMotion: times must be increasing", WDEM_ERR_INVALID};
      }
      if (s.kind == 8 || s.kind == 9) {
        if (s.nTimes <= 0 || !s.times || !s.vals) throw EngErr{s.kind == 8 ? "VariableVelocity3d: times must not be empty." : "VariableAlignedRotation: timeAngVel must not be empty.", WDEM_ERR_INVALID};
        const int w = s.kind == 8 ? 3 : 1;
        d.times.assign(s.times, s.times + s.nTimes); d.vals.assign(s.vals, s.vals + (size_t)w * s.nTimes); if (s.vals2) d.vals2.assign(s.vals2, s.vals2 + (size_t)w * s.nTimes);
        for (int k = 0; k + 1 < s.nTimes; k++) if (!(d.times[k + 1] >= d.times[k])) throw EngErr{"Impose: time values must be non-decreasing", WDEM_ERR_INVALID};
      }
      if (s.kind == 5) for (int k = 0; k < 6; k++) { const int w = (int)s.v[6 + k]; if (w != 0 && w != 1 && w != 2) throw EngErr{"Local6Dofs.whats components must be 0, 2 or 1", WDEM_ERR_INVALID}; }
      if (s.kind == 6) d.dist = s.v[4];
      if (s.kind == 10 || s.kind == 11) d.dist = s.v[5];      // .angle at upload
      if (s.kind == 6 || s.kind == 7) h->impRead = true;
      h->imposes.push_back(d);
    }
    h->impSetStart.assign(setStart, setStart + nSets + 1); h->impSetOps.assign(setOps, setOps + setStart[nSets]);
    for (int v : h->impSetOps) if (v < 0 || v >= nImp) throw EngErr{"wdem_upload_imposes: set refers to an unknown Impose object", WDEM_ERR_INVALID};
    bool orbit = false, interp = false;
    auto setKinds = [&](int set, bool& vel, bool& other) { orbit = false; interp = false; for (int e = h->impSetStart[set]; e < h->impSetStart[set + 1]; e++) { const int k = h->imposes[h->impSetOps[e]].kind; if (k == 10 || k == 11) orbit = true; if (k == 12) interp = true; if (k == 1 || k == 2 || k == 8 || k == 9 || k == 10 || k == 11 || k == 12) vel = true; else if (k == 5) { vel = true; for (int q = 0; q < 6; q++) if ((int)h->imposes[h->impSetOps[e]].v[6 + q] == 2) other = true; } else if (k == 6) { vel = true; other = true; } else other = true; } };
    auto touch = [&](int set) { if (set < 0) return; if (set >= nSets) throw EngErr{"wdem_upload_imposes: set index out of range", WDEM_ERR_INVALID}; for (int e = h->impSetStart[set]; e < h->impSetStart[set + 1]; e++) h->imposes[h->impSetOps[e]].nNodes++; };
    // spheres: table by particle id
    std::vector<int32_t> ofId((size_t)std::max(h->maxId, 0) + 2, -1);
    std::unordered_set<int32_t> known(h->hIds.begin(), h->hIds.end());
    for (int64_t k = 0; k < nSphereIds; k++) {
      if (sphereSet[k] < 0) continue;
      if (!known.count(sphereIds[k])) throw EngErr{"wdem_upload_imposes: particle #" + std::to_string(sphereIds[k]) + " is not an uploaded sphere", WDEM_ERR_INVALID};
      touch(sphereSet[k]); ofId[sphereIds[k]] = sphereSet[k]; h->impSpheres = true;
      { bool v_ = false, o_ = false; setKinds(sphereSet[k], v_, o_); if (interp) throw EngErr{"wdem_upload_imposes: InterpolatedMotion acts on the master nodes of facet groups only", WDEM_ERR_UNSUPPORTED}; }
    }
    h->dImpOfId.ensure(ofId.size(), h->bytes); CK(cudaMemcpyAsync(h->dImpOfId.p, ofId.data(), ofId.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (wallSet) {
      if (h->hWalls.size() > 32) throw EngErr{"wdem_upload_imposes: impositions on walls need at most 32 walls", WDEM_ERR_UNSUPPORTED};
      h->wallImpSet.assign(wallSet, wallSet + h->hWalls.size()); h->impWallVel.assign(3 * h->hWalls.size(), 0.);
      bool any = false;
      for (size_t w = 0; w < h->hWalls.size(); w++) {
        for (int k = 0; k < 3; k++) h->impWallVel[3 * w + k] = h->hWalls[w].vel[k];
        if (wallSet[w] < 0) continue;
        bool vel = false, other = false; setKinds(wallSet[w], vel, other);
        if (other) throw EngErr{"wdem_upload_imposes: walls (all DOFs blocked) take velocity impositions only", WDEM_ERR_UNSUPPORTED};
        if (orbit) throw EngErr{"wdem_upload_imposes: CircularOrbit / StableCircularOrbit act on spheres and on facet nodes (wdem_upload_facet_imposes), not on walls", WDEM_ERR_UNSUPPORTED};
        if (interp) throw EngErr{"wdem_upload_imposes: InterpolatedMotion acts on the master nodes of facet groups only", WDEM_ERR_UNSUPPORTED};
        touch(wallSet[w]); h->hWalls[w].moving = 1; any = true;      // DemData::guessMoving: ... || impose (Particle.cpp:638-640)
      }
      if (any) { h->anyMovingWall = true; h->anyStaticWall = false; for (auto& w : h->hWalls) if (!w.moving) h->anyStaticWall = true; CK(cudaMemcpyAsync(h->dWalls.p, h->hWalls.data(), h->hWalls.size() * sizeof(Wall), cudaMemcpyHostToDevice, st)); }
      else h->wallImpSet.clear();
    }
    if (groupSet) {
      h->groupImpSet.assign(groupSet, groupSet + h->groups.n);
      for (int g = 0; g < h->groups.n; g++) { if (groupSet[g] < 0) continue; bool vel = false, other = false; setKinds(groupSet[g], vel, other); if (other) throw EngErr{"wdem_upload_imposes: facet groups (all DOFs blocked) take velocity impositions only", WDEM_ERR_UNSUPPORTED}; if (orbit) throw EngErr{"wdem_upload_imposes: CircularOrbit / StableCircularOrbit act on spheres and on facet nodes (wdem_upload_facet_imposes), not on facet groups", WDEM_ERR_UNSUPPORTED}; touch(groupSet[g]); }
    }
    h->dImpSetStart.ensure(h->impSetStart.size(), h->bytes); CK(cudaMemcpyAsync(h->dImpSetStart.p, h->impSetStart.data(), h->impSetStart.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    h->dImpSetOps.ensure(std::max<size_t>(h->impSetOps.size(), 1), h->bytes); if (!h->impSetOps.empty()) CK(cudaMemcpyAsync(h->dImpSetOps.p, h->impSetOps.data(), h->impSetOps.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    h->dImpRing.ensure(WD_KMAX + 1, h->bytes); CK(cudaMemsetAsync(h->dImpRing.p, 0, (WD_KMAX + 1) * sizeof(ImpSlot), st));
    if (!h->hImpRing) CK(cudaMallocHost(&h->hImpRing, (WD_KMAX + 1) * sizeof(ImpSlot)));
    h->lastImpSlot = 0;
    CK(cudaStreamSynchronize(st));
  });
}
int wdem_download_boundary_velocities(wdem_handle h, double* wallVel, double* groupVel, double* groupAngVel) {
  return guarded(h, [&] {
    if (wallVel) for (size_t w = 0; w < h->hWalls.size(); w++) for (int k = 0; k < 3; k++)
      wallVel[3 * w + k] = (w < h->wallImpSet.size() && h->wallImpSet[w] >= 0) ? h->impWallVel[3 * w + k] : h->hWalls[w].vel[k];
    for (int g = 0; g < h->groups.n; g++) {
      const GroupU& G = h->groups.g[g];
      if (groupVel) { groupVel[3 * g] = G.vel.x; groupVel[3 * g + 1] = G.vel.y; groupVel[3 * g + 2] = G.vel.z; }
      if (groupAngVel) { groupAngVel[3 * g] = G.angVel.x; groupAngVel[3 * g + 1] = G.angVel.y; groupAngVel[3 * g + 2] = G.angVel.z; }
    }
  });
}
int wdem_upload_facet_imposes(wdem_handle h, const int32_t* facetSet) {
  return guarded(h, [&] {
    if (h->imposes.empty() || !facetSet) throw EngErr{"wdem_upload_facet_imposes: call after wdem_upload_imposes", WDEM_ERR_INVALID};
    if (!h->conDirty || h->nC != 0 || h->cur != 0) throw EngErr{"wdem_upload_facet_imposes: call before the first step", WDEM_ERR_INVALID};
    const int nSets = (int)h->impSetStart.size() - 1; bool any = false;
    for (size_t f = 0; f < h->hFacets.size(); f++) {
      const int set = facetSet[f]; h->hFacets[f].nodeImp1 = 0;
      if (set < 0) continue;
      if (set >= nSets) throw EngErr{"wdem_upload_facet_imposes: set index out of range", WDEM_ERR_INVALID};
      if (h->hFacets[f].group1 > 0) throw EngErr{"wdem_upload_facet_imposes: a facet of a kinematic group follows its master node and cannot carry an imposition of its own", WDEM_ERR_INVALID};
      for (int e = h->impSetStart[set]; e < h->impSetStart[set + 1]; e++) {
        wdem_engine::ImposeHost& d = h->imposes[h->impSetOps[e]];
        const bool velOnly = d.kind == 1 || d.kind == 2 || d.kind == 10 || d.kind == 11 || (d.kind == 8 && d.vals2.empty());
        if (!velOnly) throw EngErr{"wdem_upload_facet_imposes: facet nodes (all DOFs blocked) take velocity impositions only (HarmonicOscillation, AlignedHarmonicOscillations, VariableVelocity3d without rotation, CircularOrbit, StableCircularOrbit)", WDEM_ERR_UNSUPPORTED};
        d.nNodes += 3;
      }
      h->hFacets[f].nodeImp1 = set + 1; any = true;
    }
    h->impFacetNodes = any;
    if (!h->hFacets.empty()) { CK(cudaMemcpyAsync(h->dFacets.p, h->hFacets.data(), h->hFacets.size() * sizeof(Facet), cudaMemcpyHostToDevice, h->stream)); CK(cudaStreamSynchronize(h->stream)); }
    h->fgridValid = false;
  });
}
int wdem_upload_clump_imposes(wdem_handle h, const int32_t* clumpSet) {
  return guarded(h, [&] {
    if (h->imposes.empty() || !clumpSet || h->clumps.n == 0) throw EngErr{"wdem_upload_clump_imposes: call after wdem_upload_clumps and wdem_upload_imposes", WDEM_ERR_INVALID};
    if (!h->conDirty || h->nC != 0 || h->cur != 0) throw EngErr{"wdem_upload_clump_imposes: call before the first step", WDEM_ERR_INVALID};
    const int nSets = (int)h->impSetStart.size() - 1; bool any = false;
    for (int c = 0; c < h->clumps.n; c++) {
      const int set = clumpSet[c];
      if (set < 0) continue;
      if (set >= nSets) throw EngErr{"wdem_upload_clump_imposes: set index out of range", WDEM_ERR_INVALID};
      for (int e = h->impSetStart[set]; e < h->impSetStart[set + 1]; e++) {
        wdem_engine::ImposeHost& d = h->imposes[h->impSetOps[e]];
        if (d.kind == 6 || d.kind == 7) throw EngErr{"wdem_upload_clump_imposes: the read-force kinds (VelocityAndReadForce, ReadForce) are not available on clump master nodes", WDEM_ERR_UNSUPPORTED};
        if (d.kind == 12) throw EngErr{"wdem_upload_imposes: InterpolatedMotion acts on the master nodes of facet groups only", WDEM_ERR_UNSUPPORTED};
        d.nNodes++;
      }
      any = true;
    }
    h->clumps.impSet.ensure((size_t)h->clumps.n, h->bytes);
    CK(cudaMemcpyAsync(h->clumps.impSet.p, clumpSet, (size_t)h->clumps.n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream)); CK(cudaStreamSynchronize(h->stream));
    h->clumps.anyImp = any;
  });
}
int wdem_get_impose_state(wdem_handle h, int32_t ix, double* out) {
  return guarded(h, [&] {
    if (ix < 0 || ix >= (int)h->imposes.size() || !out) throw EngErr{"wdem_get_impose_state: bad arguments", WDEM_ERR_INVALID};
    double acc[8];
    CK(cudaMemcpyAsync(acc, h->dImpRing.p[h->lastImpSlot].acc + 8 * ix, sizeof(acc), cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream));
    out[0] = acc[0]; out[1] = h->imposes[ix].dist; for (int k = 0; k < 6; k++) out[2 + k] = acc[1 + k];
  });
}
int wdem_upload_clump_forces(wdem_handle h, const double* force, const double* torque) {
  return guarded(h, [&] {
    ClumpStore& K = h->clumps; const int n = K.n;
    if (!h->conDirty || h->cur != 0 || h->nC != 0) throw EngErr{"wdem_upload_clump_forces: call right after wdem_upload_clumps", WDEM_ERR_INVALID};
    std::vector<double4> b(n);
    for (int w = 0; w < 2; w++) {
      const double* src = w ? torque : force; if (!src || n == 0) continue;
      for (int i = 0; i < n; i++) b[i] = make_double4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0);
      CK(cudaMemcpyAsync((w ? K.torque : K.force).p, b.data(), n * sizeof(double4), cudaMemcpyHostToDevice, h->stream)); CK(cudaStreamSynchronize(h->stream));
    }
  });
}

int wdem_upload_forces(wdem_handle h, const double* force, const double* torque) {
  return guarded(h, [&] {
    if (!h->conDirty || h->cur != 0 || h->comm) throw EngErr{"wdem_upload_forces: call right after wdem_upload_spheres (single-domain runs)", WDEM_ERR_INVALID};
    const size_t n = (size_t)h->n; if (!n) return;
    std::vector<double> soa(6 * n, 0.);
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) { if (force) soa[k * n + i] = force[3 * i + k]; if (torque) soa[(3 + k) * n + i] = torque[3 * i + k]; }
    for (int k = 0; k < 6; k++) CK(cudaMemcpyAsync(h->sph[0].f[k].p, soa.data() + k * n, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  });
}

int wdem_add_spheres(wdem_handle h, int64_t m, const int32_t* ids, const double* pos, const double* ori, const double* vel, const double* angVel, const double* mass,
                     const double* inertia, const double* radius, const uint32_t* flags, const uint32_t* mask, const int32_t* matId) {
  return guarded(h, [&] {
    if (m < 0 || (m > 0 && (!ids || !pos || !mass || !inertia || !radius))) throw EngErr{"wdem_add_spheres: bad arguments", WDEM_ERR_INVALID};
    // Decomposed runs: every rank is called with the SAME list and keeps the spheres it is responsible for — the slab of the centre, or simply every world-th sphere in a
    // periodic cell; the collider full run that the insertion triggers sends each one home (migration) and exports the ghosts.  The id check covers what this subdomain
    // holds (owned spheres and ghost copies) plus the boundaries; global uniqueness of the ids is the caller's business, as in ParticleContainer::insertAt.
    const bool dec = h->comm != nullptr;
    if (h->hMats.empty() || (!dec && h->hIds.size() != (size_t)h->n)) throw EngErr{"wdem_add_spheres: upload materials and spheres first", WDEM_ERR_INVALID};
    if (dec && h->clumps.n) throw EngErr{"wdem_add_spheres: clumps are not supported in decomposed runs", WDEM_ERR_UNSUPPORTED};
    if (m == 0) return;
    const int64_t n0 = h->n;
    if (n0 + m > (1ll << 30)) throw EngErr{"too many spheres", WDEM_ERR_INVALID};
    cudaStream_t st = h->stream;
    std::unordered_set<int32_t> used;
    if (!dec) used.insert(h->hIds.begin(), h->hIds.end());
    for (auto& w : h->hWalls) used.insert(w.id);
    for (auto& f : h->hFacets) used.insert(f.id);
    std::vector<char> take(m, 1);
    if (dec) {
      const int W = h->comm->world, R = h->comm->rank; const DomainState& M = h->dom;
      if ((int)M.cuts.size() != W - 1) throw EngErr{"wdem_add_spheres: call wdem_set_domain first", WDEM_ERR_INVALID};
      for (int64_t i = 0; i < m; i++) {
        int r = (int)(i % W);
        if (!h->cfg.periodic) { const double x = pos[3 * i + M.axis]; r = 0; while (r + 1 < W && x >= M.cuts[r]) r++; }
        take[i] = (r == R);
      }
      if (h->id2slot.p && h->n > 0) {      // ids held here
        h->idx.ensure((size_t)m + 1, h->bytes);
        CK(cudaMemcpyAsync(h->idx.p, ids, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, st)); CK(cudaMemsetAsync(h->idx.p + m, 0, sizeof(int), st));
        k_dom_ids_in_use<<<nblk(m, 256), 256, 0, st>>>(h->idx.p, (int)m, h->id2slot.p, (long long)h->id2slot.cap, h->n, h->idx.p + m); h->launches++;
        int inUse = 0; CK(cudaMemcpyAsync(&inUse, h->idx.p + m, sizeof(int), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        if (inUse) throw EngErr{"wdem_add_spheres: particle id " + std::to_string(inUse - 1) + " is negative or already in use", WDEM_ERR_INVALID};
      }
    }
    std::vector<double4> posr, velm, angi, q; std::vector<uint32_t> fl, mk; std::vector<int32_t> mt, id, up;
    int maxId = std::max(h->maxId, dec ? h->dom.globalMaxId : -1); double minR = h->minRadius, pw = h->pwaveDt, rmax = h->dom.rmaxGlobal;
    for (int64_t i0 = 0; i0 < m; i0++) {
      const int32_t idNew = ids[i0];
      if (idNew < 0 || !used.insert(idNew).second) throw EngErr{"wdem_add_spheres: particle id " + std::to_string(idNew) + " is negative or already in use", WDEM_ERR_INVALID};
      maxId = std::max(maxId, idNew);
      const uint32_t f_ = flags ? flags[i0] : 0u; const int32_t mt_ = matId ? matId[i0] : 0;
      if (mt_ < 0 || mt_ >= (int)h->hMats.size()) throw EngErr{"sphere material index out of range", WDEM_ERR_INVALID};
      if (f_ & (64u | 128u)) throw EngErr{"wdem_add_spheres: clump members cannot be added to a running scene", WDEM_ERR_UNSUPPORTED};
      if (!(radius[i0] > 0)) throw EngErr{"sphere #" + std::to_string(idNew) + ": radius must be positive", WDEM_ERR_INVALID};
      if ((f_ & 7u) != 7u && !(mass[i0] > 0)) throw EngErr{"DemField.selfTest: particle #" + std::to_string(idNew) + " has non-positive mass with free translational DOFs", WDEM_ERR_INVALID};
      if ((f_ & 56u) != 56u && !(inertia[i0] > 0)) throw EngErr{"DemField.selfTest: particle #" + std::to_string(idNew) + " has non-positive inertia with free rotational DOFs", WDEM_ERR_INVALID};
      // scene-wide quantities follow the whole list on every rank (they stay consistent without another exchange)
      minR = std::min(minR, radius[i0]); rmax = std::max(rmax, radius[i0]);
      const Material& mm = h->hMats[mt_];
      pw = std::min(pw, radius[i0] / (1. * std::sqrt(mm.young / mm.density)));
      if (!take[i0]) continue;
      id.push_back(idNew); up.push_back((int32_t)(n0 + (int64_t)up.size())); fl.push_back(f_); mk.push_back(mask ? mask[i0] : 5u); mt.push_back(mt_);
      posr.push_back(make_double4(pos[3 * i0], pos[3 * i0 + 1], pos[3 * i0 + 2], radius[i0]));
      velm.push_back(make_double4(vel ? vel[3 * i0] : 0, vel ? vel[3 * i0 + 1] : 0, vel ? vel[3 * i0 + 2] : 0, mass[i0]));
      angi.push_back(make_double4(angVel ? angVel[3 * i0] : 0, angVel ? angVel[3 * i0 + 1] : 0, angVel ? angVel[3 * i0 + 2] : 0, inertia[i0]));
      q.push_back(ori ? make_double4(ori[4 * i0], ori[4 * i0 + 1], ori[4 * i0 + 2], ori[4 * i0 + 3]) : make_double4(0, 0, 0, 1));
    }
    const int64_t k = (int64_t)id.size(), n1 = n0 + k;
    SphereStore& s = h->sph[h->cur];
    s.ensure((size_t)std::max<int64_t>(n1, 1), h->bytes, /*keep*/ true, st);
    auto H2D = [&](void* d, const void* src, size_t b) { if (b) CK(cudaMemcpyAsync(d, src, b, cudaMemcpyHostToDevice, st)); };
    H2D(s.posr.p + n0, posr.data(), k * sizeof(double4)); H2D(s.velm.p + n0, velm.data(), k * sizeof(double4)); H2D(s.angi.p + n0, angi.data(), k * sizeof(double4)); H2D(s.ori.p + n0, q.data(), k * sizeof(double4));
    if (k) for (auto& b : s.f) CK(cudaMemsetAsync(b.p + n0, 0, k * sizeof(double), st));     // a new particle sees zero force in its first Leapfrog
    if (k) for (auto& b : s.l) CK(cudaMemsetAsync(b.p + n0, 0, k * sizeof(double), st));
    H2D(s.flags.p + n0, fl.data(), k * 4); H2D(s.mask.p + n0, mk.data(), k * 4); H2D(s.mat.p + n0, mt.data(), k * 4); H2D(s.id.p + n0, id.data(), k * 4); H2D(s.up.p + n0, up.data(), k * 4);
    const size_t oldCap = h->id2slot.cap;
    h->id2slot.ensure((size_t)maxId + 2, h->bytes, /*keep*/ true, st);
    if (h->id2slot.cap > oldCap) CK(cudaMemsetAsync(h->id2slot.p + oldCap, 0xFF, (h->id2slot.cap - oldCap) * sizeof(int), st));
    if (dec && k) { k_dom_map_appended<<<nblk(k, 256), 256, 0, st>>>(s.id.p, (int)n0, (int)k, h->id2slot.p); h->launches++; }
    CK(cudaStreamSynchronize(st));
    if (h->compactPending && k) {   // the appended slots have no counterpart in the generation the incidence rows were built for
      h->dom.origSlot.ensure((size_t)n1, h->bytes, /*keep*/ true, st);
      CK(cudaMemsetAsync(h->dom.origSlot.p + n0, 0xFF, (size_t)k * sizeof(int), st)); CK(cudaStreamSynchronize(st));
    }
    h->n = (int)n1; h->minRadius = minR; h->pwaveDt = pw; h->maxId = maxId; h->dom.rmaxGlobal = rmax;
    if (dec) { h->dom.nOwn += (int)k; h->dom.globalMaxId = std::max(h->dom.globalMaxId, maxId); }
    else { h->dom.nOwn = (int)n1; h->hIds.insert(h->hIds.end(), id.begin(), id.end()); }
    h->conDirty = true; h->up2slotValid = false;     // ParticleContainer::add marks the contacts dirty: the collider re-initialises at the next step (ParticleContainer.cpp:23-73)
  });
}

int wdem_remove_spheres(wdem_handle h, int64_t m, const int32_t* ids) {
  return guarded(h, [&] {
    if (m < 0 || (m > 0 && !ids)) throw EngErr{"wdem_remove_spheres: bad arguments", WDEM_ERR_INVALID};
    if (h->clumps.n) throw EngErr{"wdem_remove_spheres: not available in scenes with clumps", WDEM_ERR_UNSUPPORTED};
    if (m == 0) return;
    const int n0 = h->n; cudaStream_t st = h->stream; int64_t& tot = h->bytes;
    if (h->comm) {
      // Decomposed runs: every rank is called with the SAME list and lets go of whatever it holds of it: an owned sphere stops being owned (flagged like a ghost of
      // the previous generation, which the full run that follows drops), a ghost copy is dropped anyway; in both cases the sphere's contacts go with it and its id
      // is free again.  Ids this subdomain does not hold are some other subdomain's business (no global "unknown id" check).
      std::unordered_set<int32_t> gone(ids, ids + m);
      if ((int64_t)gone.size() != m) throw EngErr{"wdem_remove_spheres: duplicate ids", WDEM_ERR_INVALID};
      if (n0 == 0 || !h->id2slot.p) return;
      h->idx.ensure((size_t)m + 1, tot); h->counts.ensure((size_t)n0 + 1, tot);
      CK(cudaMemcpyAsync(h->idx.p, ids, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, st)); CK(cudaMemsetAsync(h->idx.p + m, 0, sizeof(int), st));
      k_fill_int<<<nblk(n0, 256), 256, 0, st>>>(h->counts.p, n0, 1);
      SphereStore& A = h->sph[h->cur];
      k_dom_mark_removed<<<nblk(m, 256), 256, 0, st>>>(h->idx.p, (int)m, h->id2slot.p, (long long)h->id2slot.cap, n0, A.flags.p, h->counts.p, h->idx.p + m);
      if (h->csr[h->curR].valid && h->nC > 0) {
        CsrStore& R = h->csr[h->curR]; ContactStore& cs = h->con[h->curC];
        k_kill_rows<<<nblk(n0, 256), 256, 0, st>>>(A.id.p, n0, h->counts.p, h->id2slot.p, R.rowStart.p, R.rowEntry.p, R.nRows, cs.flags.p, cs.realBits.p); h->launches++;
      }
      k_dom_forget_ids<<<nblk(m, 256), 256, 0, st>>>(h->idx.p, (int)m, h->id2slot.p, (long long)h->id2slot.cap);
      int nOwnedGone = 0; CK(cudaMemcpyAsync(&nOwnedGone, h->idx.p + m, sizeof(int), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
      h->launches += 3;
      h->dom.nOwn -= nOwnedGone;
      h->conDirty = true; h->up2slotValid = false;
      return;
    }
    if (h->hIds.size() != (size_t)n0) throw EngErr{"wdem_remove_spheres: no spheres uploaded", WDEM_ERR_INVALID};
    std::unordered_set<int32_t> gone(ids, ids + m);
    if ((int64_t)gone.size() != m) throw EngErr{"wdem_remove_spheres: duplicate ids", WDEM_ERR_INVALID};
    std::vector<int> newUpOf(n0); std::vector<int32_t> keptIds; keptIds.reserve(n0); size_t found = 0;
    for (int u = 0; u < n0; u++) { if (gone.count(h->hIds[u])) { newUpOf[u] = -1; found++; } else { newUpOf[u] = (int)keptIds.size(); keptIds.push_back(h->hIds[u]); } }
    if (found != (size_t)m) throw EngErr{"wdem_remove_spheres: unknown sphere id", WDEM_ERR_INVALID};
    const int n1 = n0 - (int)m;
    SphereStore& A = h->sph[h->cur]; SphereStore& B = h->sph[1 - h->cur];
    B.ensure((size_t)std::max(n1, 1), tot);
    h->idx.ensure((size_t)n0 + 1, tot); h->counts.ensure((size_t)n0 + 1, tot); h->offsets.ensure((size_t)n0 + 1, tot); h->dom.origSlot.ensure((size_t)std::max(n1, 1), tot);
    CK(cudaMemcpyAsync(h->idx.p, newUpOf.data(), (size_t)n0 * sizeof(int), cudaMemcpyHostToDevice, st));
    if (!h->compactPending && h->csr[h->curR].valid) {   // ids by slot of the generation the rows / the contact table refer to (k_orphans)
      h->dom.oldIds.ensure((size_t)n0, tot);
      CK(cudaMemcpyAsync(h->dom.oldIds.p, A.id.p, (size_t)n0 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    k_keep_flag<<<nblk(n0, 256), 256, 0, st>>>(A.up.p, n0, h->idx.p, h->counts.p);
    if (h->csr[h->curR].valid && h->nC > 0) {   // the removed particles' contacts go with them (before id2slot forgets where their rows are)
      CsrStore& R = h->csr[h->curR]; ContactStore& cs = h->con[h->curC];
      k_kill_rows<<<nblk(n0, 256), 256, 0, st>>>(A.id.p, n0, h->counts.p, h->id2slot.p, R.rowStart.p, R.rowEntry.p, R.nRows, cs.flags.p, cs.realBits.p); h->launches++;
    }
    exclusiveSum(h, h->counts.p, h->offsets.p, n0);
    k_compact_spheres<<<nblk(n0, 256), 256, 0, st>>>(B.view(), A.view(), n0, h->counts.p, h->offsets.p, h->idx.p, h->id2slot.p, h->dom.origSlot.p);
    CK(cudaStreamSynchronize(st));   // newUpOf is pageable host memory
    h->launches += 4;
    h->cur = 1 - h->cur; h->n = n1; h->dom.nOwn = n1; h->hIds.swap(keptIds);
    h->compactPending = h->csr[h->curR].valid;   // before the first full run there is no history to look up
    h->conDirty = true; h->up2slotValid = false;  // ParticleContainer::remove marks the contacts dirty (ParticleContainer.cpp:121-157)
  });
}

int wdem_upload_walls(wdem_handle h, int n, const int32_t* ids, const double* pos, const int32_t* axis, const int32_t* sense, const double* vel, const uint32_t* mask,
                      const int32_t* matId) {
  return guarded(h, [&] {
    if (n < 0 || (n > 0 && (!ids || !pos || !axis || !sense))) throw EngErr{"wdem_upload_walls: bad arguments", WDEM_ERR_INVALID};
    h->hWalls.resize(n);
    for (int i = 0; i < n; i++) {
      Wall w; memset(&w, 0, sizeof(w));
      for (int k = 0; k < 3; k++) { w.pos[k] = pos[3 * i + k]; w.vel[k] = vel ? vel[3 * i + k] : 0.; }
      w.axis = axis[i]; w.sense = sense[i]; w.mask = mask ? mask[i] : 3u; w.mat = matId ? matId[i] : 0; w.id = ids[i];
      w.moving = (w.vel[0] != 0. || w.vel[1] != 0. || w.vel[2] != 0.) ? 1 : 0;
      for (int k = 0; k < 3; k++) w.lastPos[k] = w.pos[k];
      if (w.axis < 0 || w.axis > 2 || w.sense < -1 || w.sense > 1) throw EngErr{"Wall.axis must be 0..2 and Wall.sense -1,0,1", WDEM_ERR_INVALID};
      if (w.mat < 0 || w.mat >= (int)h->hMats.size()) throw EngErr{"wall material index out of range", WDEM_ERR_INVALID};
      h->hWalls[i] = w;
    }
    h->anyMovingWall = h->anyStaticWall = false; for (auto& w : h->hWalls) { if (w.moving) h->anyMovingWall = true; else h->anyStaticWall = true; }
    h->wallAcc.ensure((size_t)std::max(n, 1) * 6 + 1, h->bytes);      // + the ticket word
    CK(cudaMemsetAsync(h->wallAcc.p, 0, ((size_t)std::max(n, 1) * 6 + 1) * sizeof(double), h->stream));
    h->dWalls.ensure(std::max(n, 1), h->bytes);
    if (n) CK(cudaMemcpyAsync(h->dWalls.p, h->hWalls.data(), n * sizeof(Wall), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->conDirty = true;
  });
}

int wdem_upload_facets(wdem_handle h, int n, const int32_t* ids, const double* verts, const double* vvel, const double* fakeVel, const double* halfThick, const double* n21lim,
                       const uint32_t* mask, const int32_t* matId) {
  return guarded(h, [&] {
    if (n < 0 || (n > 0 && (!ids || !verts))) throw EngErr{"wdem_upload_facets: bad arguments", WDEM_ERR_INVALID};
    h->hFacets.resize(n);
    for (int i = 0; i < n; i++) {
      Facet f; memset(&f, 0, sizeof(f));
      for (int k = 0; k < 9; k++) { f.v[k] = verts[9 * i + k]; f.vvel[k] = vvel ? vvel[9 * i + k] : 0.; }
      for (int k = 0; k < 3; k++) { f.fakeVel[k] = fakeVel ? fakeVel[3 * i + k] : 0.; f.n21lim[k] = n21lim ? n21lim[3 * i + k] : NAN; }
      f.halfThick = halfThick ? halfThick[i] : 0.; f.mask = mask ? mask[i] : 3u; f.mat = matId ? matId[i] : 0; f.id = ids[i];
      if (f.mat < 0 || f.mat >= (int)h->hMats.size()) throw EngErr{"facet material index out of range", WDEM_ERR_INVALID};
      // Bo1_Facet_Aabb::go (Facet.cpp:275-298), aperiodic
      for (int k = 0; k < 3; k++) {
        f.bmin[k] = f.v[k] - f.halfThick; f.bmax[k] = f.v[k] + f.halfThick;
        for (int j = 1; j < 3; j++) { f.bmin[k] = std::min(f.bmin[k], f.v[3 * j + k] - f.halfThick); f.bmax[k] = std::max(f.bmax[k], f.v[3 * j + k] + f.halfThick); }
      }
      h->hFacets[i] = f;
    }
    h->dFacets.ensure(std::max(n, 1), h->bytes);
    if (n) CK(cudaMemcpyAsync(h->dFacets.p, h->hFacets.data(), n * sizeof(Facet), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->conDirty = true; h->fgridValid = false; h->groups.n = 0; h->groupRadius.clear();
  });
}

int wdem_upload_facet_groups(wdem_handle h, int nGroups, const int32_t* groupOfFacet, const double* pos, const double* ori, const double* vel, const double* angVel) {
  return guarded(h, [&] {
    const int nf = (int)h->hFacets.size();
    if (nGroups < 0 || nGroups > WD_MAX_GROUPS) throw EngErr{"wdem_upload_facet_groups: at most " + std::to_string(WD_MAX_GROUPS) + " groups", WDEM_ERR_UNSUPPORTED};
    if (nGroups > 0 && (!groupOfFacet || !pos || !ori || !vel || !angVel)) throw EngErr{"wdem_upload_facet_groups: bad arguments", WDEM_ERR_INVALID};
    if (!h->conDirty || h->nC != 0) throw EngErr{"wdem_upload_facet_groups: call after wdem_upload_facets and before the first step", WDEM_ERR_INVALID};
    h->groups.n = nGroups; h->groupRadius.assign(nGroups, 0.);
    for (int g = 0; g < nGroups; g++) {
      GroupU& G = h->groups.g[g];
      G.pos = mk3(pos[3 * g], pos[3 * g + 1], pos[3 * g + 2]); G.ori.x = ori[4 * g]; G.ori.y = ori[4 * g + 1]; G.ori.z = ori[4 * g + 2]; G.ori.w = ori[4 * g + 3];
      G.vel = mk3(vel[3 * g], vel[3 * g + 1], vel[3 * g + 2]); G.angVel = mk3(angVel[3 * g], angVel[3 * g + 1], angVel[3 * g + 2]);
    }
    for (int f = 0; f < nf; f++) {
      Facet& F = h->hFacets[f];
      const int g = groupOfFacet[f];
      if (g >= nGroups) throw EngErr{"wdem_upload_facet_groups: group index out of range", WDEM_ERR_INVALID};
      F.group1 = g < 0 ? 0 : g + 1;
      for (int k = 0; k < 9; k++) F.lastv[k] = F.v[k];
      if (g < 0) continue;
      const GroupU& G = h->groups.g[g]; q4 qc = G.ori; qc.x = -qc.x; qc.y = -qc.y; qc.z = -qc.z;      // conjugate
      for (int i = 0; i < 3; i++) {
        const d3 d = mk3(F.v[3 * i], F.v[3 * i + 1], F.v[3 * i + 2]) - G.pos;
        const d3 r = quatRotate(qc, d);                                                               // ClumpData.relPos (Clump.cpp:252)
        F.rel[3 * i] = r.x; F.rel[3 * i + 1] = r.y; F.rel[3 * i + 2] = r.z;
        const d3 vv = G.vel + cross(G.angVel, d);
        F.vvel[3 * i] = vv.x; F.vvel[3 * i + 1] = vv.y; F.vvel[3 * i + 2] = vv.z;
        h->groupRadius[g] = std::max(h->groupRadius[g], nrm(d));
      }
    }
    if (nf) CK(cudaMemcpyAsync(h->dFacets.p, h->hFacets.data(), nf * sizeof(Facet), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->fgridValid = false;
  });
}
int wdem_download_facets(wdem_handle h, double* verts, double* groupPos, double* groupOri) {
  return guarded(h, [&] {
    const size_t nf = h->hFacets.size();
    if (verts && nf) { auto F = d2h(h->dFacets.p, nf, h->stream); for (size_t f = 0; f < nf; f++) for (int k = 0; k < 9; k++) verts[9 * f + k] = F[f].v[k]; }
    for (int g = 0; g < h->groups.n; g++) {
      const GroupU& G = h->groups.g[g];
      if (groupPos) { groupPos[3 * g] = G.pos.x; groupPos[3 * g + 1] = G.pos.y; groupPos[3 * g + 2] = G.pos.z; }
      if (groupOri) { groupOri[4 * g] = G.ori.x; groupOri[4 * g + 1] = G.ori.y; groupOri[4 * g + 2] = G.ori.z; groupOri[4 * g + 3] = G.ori.w; }
    }
  });
}

int wdem_set_cell(wdem_handle h, const double* hSize, const double* gradV, const double* nextGradV, int homoDeform) {
  return guarded(h, [&] {
    if (!hSize) throw EngErr{"wdem_set_cell: hSize required", WDEM_ERR_INVALID};
    HostCell& c = h->cell;
    for (int i = 0; i < 9; i++) { c.hSize.m[i] = hSize[i]; c.gradV.m[i] = gradV ? gradV[i] : 0.; c.nextGradV.m[i] = nextGradV ? nextGradV[i] : c.gradV.m[i]; }
    c.pprevHsize = c.hSize; c.trsf = mIdent(); c.homoDeform = homoDeform;
    if (homoDeform < 0 || homoDeform > 4) throw EngErr{"Cell.homoDeform must be in {0,1,2,3,4}.", WDEM_ERR_INVALID};
    c.integrateAndUpdate(0);
  });
}
int wdem_set_next_gradv(wdem_handle h, const double* g) { return guarded(h, [&] { for (int i = 0; i < 9; i++) h->cell.nextGradV.m[i] = g[i]; }); }
int wdem_get_cell(wdem_handle h, double* hSize) { return guarded(h, [&] { for (int i = 0; i < 9; i++) hSize[i] = h->cell.hSize.m[i]; }); }

static void stepN(wdem_handle h, int64_t nsteps, wdem_stats* stats, bool cheapStats) {
  {
    wdem_stats local; memset(&local, 0, sizeof(local));
    for (double& p : h->phaseAcc) p = 0.;
    long long l0 = h->launches;
    CK(cudaMemsetAsync(h->dCnt, 0, 4 * sizeof(unsigned long long), h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    runSteps(h, nsteps, &local);
    CK(cudaEventRecord(h->ev1, h->stream));
    if (h->e2e.active && h->e2e.defer) return;      // deferred tail: the contact pass is still running; its errors (sticky on the device) surface with the next call
    if (h->comm) h->comm->allreduceMaxInt(&h->dCnt->error, 1, h->stream);
    fetchCounters(h);
    CK(cudaEventSynchronize(h->ev1));
    for (int k = 0; k < wdem_engine::NEVSET; k++) harvestPhases(h, k);
    checkDeviceError(h);
    const double* phase = h->phaseAcc;
    float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (stats) {
      local.device_ms = ms; local.n_spheres = h->comm ? h->dom.nOwn : h->n; local.n_full_runs_total = h->nFullRuns; local.n_launches = h->launches - l0; local.step = h->step; local.dt = h->dt; local.time = h->time;
      local.ms_integrate = phase[0]; local.ms_rebuild = phase[1]; local.ms_contact = phase[2]; local.ms_reduce = phase[3]; local.ms_dyndt = phase[4];
      // cheapStats (end-to-end path, one step per call): no extra counting kernel + sync; the step's own counters give the real contacts exactly
      // and the table size stands in for the number of live candidates
      long long nr, na;
      if (cheapStats) { nr = (long long)h->hCnt->nReal; na = h->nC; } else countContacts(h, nr, na);
      local.n_contacts = na; local.n_real = nr;
      local.sum_real = (int64_t)h->hCnt->nReal; local.sum_potential_evaluated = (int64_t)h->hCnt->nPotEval - (int64_t)h->hCnt->nReal; local.sum_contacts = (int64_t)h->hCnt->nPotEval;
      *stats = local;
    }
  }
}
int wdem_step(wdem_handle h, int64_t nsteps, wdem_stats* stats) { return guarded(h, [&] { stepN(h, nsteps, stats, false); }); }

// End-to-end step of one subdomain (or of a single-domain engine addressed by id): the host sends the state of the spheres it owns BY ID, in any order, and gets the new
// pos / vel back in the same order.  Same pipeline as wdem_step_e2e (chunked upload -> integrate -> download, contact pass overlapping the download and, with
// stats == NULL, the next call's upload); the ids are turned into slots on the device first.  *owned_changed != 0: spheres migrated in this step's collider run - the
// outputs are still those of the spheres sent, but the next call must send the new owned set (wdem_download_owned).
int wdem_step_e2e_owned(wdem_handle h, int64_t m, const int32_t* ids, const double* pos_in, const double* vel_in, const double* angVel_in, double* pos_out, double* vel_out,
                        double* ek_out, int32_t* owned_changed, wdem_stats* stats) {
  return guarded(h, [&] {
    if (m <= 0 || !ids || !pos_in || !vel_in || !angVel_in) throw EngErr{"wdem_step_e2e_owned: bad arguments", WDEM_ERR_INVALID};
    if (h->conDirty) throw EngErr{"wdem_step_e2e_owned: step at least once first (ownership is settled by the first collider run)", WDEM_ERR_INVALID};
    const long long nOwn0 = h->comm ? h->dom.nOwn : h->n;
    if (m != nOwn0) throw EngErr{"wdem_step_e2e_owned: " + std::to_string(m) + " spheres sent, " + std::to_string(nOwn0) + " owned", WDEM_ERR_INVALID};
    if (!h->imposes.empty() || h->clumps.n) throw EngErr{"wdem_step_e2e_owned: Impose and clumps are not available on the end-to-end path", WDEM_ERR_UNSUPPORTED};
    cudaStream_t st = h->stream;
    const long long mig0 = h->dom.migratedOut + h->dom.migratedIn;
    h->ownedIds.ensure((size_t)m, h->bytes); h->ownedSlot.ensure((size_t)m, h->bytes); h->orphCnt.ensure(1, h->bytes);
    { const double *i0 = h->ioIn.p, *o0 = h->ioOut.p; h->ioIn.ensure(9 * (size_t)m, h->bytes); h->ioOut.ensure(6 * (size_t)m, h->bytes);
      if (h->ioIn.p != i0 || h->ioOut.p != o0) h->e2e.ioBusy = false; }      // new staging buffers: the copy streams must be ordered behind their (stream-ordered) allocation
    // the main stream has to wait for nothing here but its own past; ids -> slots (+ are they all owned?) before the first chunk is integrated
    CK(cudaMemcpyAsync(h->ownedIds.p, ids, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(h->orphCnt.p, 0, sizeof(int), st));
    k_owned_slots<<<nblk(m, 256), 256, 0, st>>>(h->ownedIds.p, (int)m, h->id2slot.p, (long long)h->id2slot.cap, h->n, h->S().flags, h->ownedSlot.p, h->orphCnt.p); h->launches++;
    int* hMiss = h->hScratch + 12; *hMiss = 0;
    CK(cudaMemcpyAsync(hMiss, h->orphCnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    h->e2e.defer = !stats && !h->profiling && !getenv("WDEM_E2E_NO_DEFER");
    h->e2e.count = m; h->e2e.slotMap = h->ownedSlot.p;
    h->e2e.active = true; h->e2e.pos = pos_in; h->e2e.vel = vel_in; h->e2e.angVel = angVel_in; h->e2e.posOut = pos_out; h->e2e.velOut = vel_out;
    try { stepN(h, 1, stats, true); } catch (...) { h->e2e.active = false; cudaStreamSynchronize(h->sH2D); cudaStreamSynchronize(h->sD2H); throw; }
    h->e2e.active = false;
    CK(cudaEventRecord(h->evEk, h->sD2H));
    CK(cudaStreamSynchronize(h->sD2H));
    if (*hMiss) throw EngErr{"wdem_step_e2e_owned: " + std::to_string(*hMiss) + " of the ids sent are not owned by this subdomain (the state is not usable any more)", WDEM_ERR_INVALID};
    if (ek_out) { double e = 0; for (int b = 0; b < h->ekBlocks; b++) e += h->hEk[b]; *ek_out = e; }
    if (owned_changed) *owned_changed = (h->dom.migratedOut + h->dom.migratedIn != mig0 || (h->comm ? h->dom.nOwn : h->n) != nOwn0) ? 1 : 0;
  });
}

int wdem_set_state(wdem_handle h, int64_t first, int64_t cnt, const double* pos, const double* vel, const double* angVel) {
  return guarded(h, [&] {
    if (h->comm) throw EngErr{"wdem_set_state: not available in decomposed runs (spheres migrate between subdomains)", WDEM_ERR_UNSUPPORTED};
    if (first < 0 || cnt < 0 || first + cnt > h->n) throw EngErr{"wdem_set_state: range out of bounds", WDEM_ERR_INVALID};
    if (cnt == 0) return;
    ensureUp2Slot(h);
    h->ioBuf.ensure((size_t)h->n * 19, h->bytes);
    double *dp = nullptr, *dv = nullptr, *dw = nullptr;
    if (pos) { dp = h->ioBuf.p; CK(cudaMemcpyAsync(dp, pos, cnt * 24, cudaMemcpyHostToDevice, h->stream)); }
    if (vel) { dv = h->ioBuf.p + 3 * cnt; CK(cudaMemcpyAsync(dv, vel, cnt * 24, cudaMemcpyHostToDevice, h->stream)); }
    if (angVel) { dw = h->ioBuf.p + 6 * cnt; CK(cudaMemcpyAsync(dw, angVel, cnt * 24, cudaMemcpyHostToDevice, h->stream)); }
    k_scatter_state<<<nblk(cnt, 256), 256, 0, h->stream>>>(h->S(), h->n, h->up2slot.p, first, cnt, dp, dv, dw);
    CK(cudaStreamSynchronize(h->stream));
  });
}

static void downloadInto(wdem_handle h, double* pos, double* ori, double* vel, double* angVel, double* force, double* torque) {
  if (h->comm) throw EngErr{"decomposed run: use wdem_download_owned (upload order is not defined across subdomains)", WDEM_ERR_UNSUPPORTED};
  const int n = h->n; if (n == 0) return;
  h->ioBuf.ensure((size_t)n * 19, h->bytes);
  double* d = h->ioBuf.p;
  double *dpos = pos ? d : nullptr, *dori = ori ? d + 3 * (size_t)n : nullptr, *dvel = vel ? d + 7 * (size_t)n : nullptr, *dw = angVel ? d + 10 * (size_t)n : nullptr,
         *df = force ? d + 13 * (size_t)n : nullptr, *dt_ = torque ? d + 16 * (size_t)n : nullptr;
  k_gather_state<<<nblk(n, 256), 256, 0, h->stream>>>(h->S(), n, dpos, dori, dvel, dw, df, dt_);
  auto D2H = [&](double* dst, double* src, size_t cnt) { if (dst) CK(cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); };
  D2H(pos, dpos, 3 * (size_t)n); D2H(ori, dori, 4 * (size_t)n); D2H(vel, dvel, 3 * (size_t)n); D2H(angVel, dw, 3 * (size_t)n); D2H(force, df, 3 * (size_t)n); D2H(torque, dt_, 3 * (size_t)n);
  CK(cudaStreamSynchronize(h->stream));
}
int wdem_download_state(wdem_handle h, double* pos, double* ori, double* vel, double* angVel) { return guarded(h, [&] { downloadInto(h, pos, ori, vel, angVel, nullptr, nullptr); }); }
int wdem_download_forces(wdem_handle h, double* force, double* torque) { return guarded(h, [&] { downloadInto(h, nullptr, nullptr, nullptr, nullptr, force, torque); }); }

int wdem_kinetic_energy(wdem_handle h, double* ek) {
  return guarded(h, [&] {
    const int n = h->n; double e = 0;
    if (n) {
      const int nb = nblk(n, 256); h->blk.ensure((size_t)nb * 7, h->bytes);
      k_ekin<<<nb, 256, 0, h->stream>>>(h->S(), n, makeUniforms(h), h->blk.p);
      auto hb = d2h(h->blk.p, nb, h->stream);
      for (double v : hb) e += v;
    }
    *ek = e;
  });
}

int wdem_download_walls(wdem_handle h, double* pos, double* force, double* torque) {
  return guarded(h, [&] {
    const size_t nW = h->hWalls.size(); if (!nW) return;
    cudaStream_t st = h->stream;
    auto W = d2h(h->dWalls.p, nW, st);
    for (size_t w = 0; w < nW; w++) for (int k = 0; k < 3; k++) { if (pos) pos[3 * w + k] = W[w].pos[k]; if (force) force[3 * w + k] = 0.; if (torque) torque[3 * w + k] = 0.; }
    if (!force && !torque) return;
    if (h->nC == 0 || (!h->anyMovingWall && !h->compactPending)) {   // no wall in DemField.nodes: only the accumulated values of the walls at rest
      auto acc = d2h(h->wallAcc.p, nW * 6, st);
      for (size_t w = 0; w < nW; w++) for (int k = 0; k < 3; k++) { if (force) force[3 * w + k] = acc[6 * w + k]; if (torque) torque[3 * w + k] = acc[6 * w + 3 + k]; }
      return;
    }
    if (h->compactPending) throw EngErr{"spheres were removed since the last step: step once before reading contacts, wall forces or the stress (the table still refers to the old slots)", WDEM_ERR_INVALID};

    // wall-sphere contacts live in [nSS, nSS+nWS) and among the orphans; forces are added per wall in ascending id of the sphere, the order of
    // the per-particle contact map the reference walks (Particle.hpp:36)
    const long long nC = h->nC; ContactStore& cs = h->con[h->curC];
    auto flags = d2h(cs.flags.p, nC, st); auto pair = d2h(cs.pair.p, nC, st); auto rCv = d2h(cs.r[2].p, nC, st), rDv = d2h(cs.r[3].p, nC, st);
    auto Fg = [&](long long c) { return rDv[c]; }; auto cp = [&](long long c) { return rCv[c]; };
    auto sid = d2h(h->S().id, h->n, st); auto sfl = d2h(h->S().flags, h->n, st);
    const bool rolling = (h->cfg.relRollStiff > 0. && h->cfg.rollTanPhi > 0.);
    std::vector<double4> Tg; if (rolling) Tg = d2h(cs.Tg.p, nC, st);
    struct Row { int wall, sid; long long c; };
    std::vector<Row> rows;
    for (long long c = 0; c < nC; c++) {
      if (!(flags[c] & CF_REAL) || (flags[c] & CF_DEAD) || cf_kind(flags[c]) != KIND_WS) continue;
      if (sfl[pair[c].y] & F_GHOST) continue;
      rows.push_back(Row{pair[c].x, sid[pair[c].y], c});
    }
    std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.wall != b.wall ? a.wall < b.wall : a.sid < b.sid; });
    std::vector<d3> F(nW, mk3(0, 0, 0)), T(nW, mk3(0, 0, 0));
    for (const Row& r : rows) {
      const d3 Fc = xyz(Fg(r.c));                                     // wall = pA: sign +1 (Contact.cpp:28-31)
      const d3 xc = xyz(cp(r.c)) - mk3(W[r.wall].pos[0], W[r.wall].pos[1], W[r.wall].pos[2]);
      d3 Tc = cross(xc, Fc);
      if (rolling) Tc = Tc + xyz(Tg[r.c]);
      F[r.wall] = F[r.wall] + Fc; T[r.wall] = T[r.wall] + Tc;
    }
    // walls at rest: the value accumulated since the upload (k_wall_sum / k_wall_acc); moving walls (reset by Leapfrog every step): this step's contacts
    auto acc = d2h(h->wallAcc.p, nW * 6, st);
    for (size_t w = 0; w < nW; w++) {
      const bool mv = W[w].moving != 0;
      if (force) for (int k = 0; k < 3; k++) force[3 * w + k] = mv ? comp(F[w], k) : acc[6 * w + k];
      if (torque) for (int k = 0; k < 3; k++) torque[3 * w + k] = mv ? comp(T[w], k) : acc[6 * w + 3 + k];
    }
  });
}

int wdem_track_energy(wdem_handle h, int on) {
  return guarded(h, [&] {
    if (on && !h->trackEnergy) { h->dEnergy.ensure(EN_COUNT, h->bytes); CK(cudaMemsetAsync(h->dEnergy.p, 0, EN_COUNT * sizeof(double), h->stream)); CK(cudaStreamSynchronize(h->stream)); }
    h->trackEnergy = on != 0;
  });
}
int wdem_get_energy(wdem_handle h, double* out8) {
  return guarded(h, [&] {
    if (!out8) throw EngErr{"wdem_get_energy: null output", WDEM_ERR_INVALID};
    for (int k = 0; k < EN_COUNT; k++) out8[k] = 0.;
    if (!h->trackEnergy) throw EngErr{"Scene.trackEnergy is off (wdem_track_energy)", WDEM_ERR_INVALID};
    auto v = d2h(h->dEnergy.p, EN_COUNT, h->stream);
    for (int k = 0; k < EN_COUNT; k++) out8[k] = v[k];
  });
}

int wdem_get_stress(wdem_handle h, double* stress9) {
  return guarded(h, [&] {
    if (!stress9) throw EngErr{"wdem_get_stress: null output", WDEM_ERR_INVALID};
    computeStress(h, stress9);
  });
}
int wdem_get_gradv_work(wdem_handle h, double* work) { return guarded(h, [&] { if (!work) throw EngErr{"wdem_get_gradv_work: null output", WDEM_ERR_INVALID}; *work = h->gradVWork; }); }

int wdem_step_e2e(wdem_handle h, int64_t nsteps, const double* pos_in, const double* vel_in, const double* angVel_in, double* pos_out, double* vel_out, double* ek_out,
                  wdem_stats* stats) {
  if (h && !h->comm && nsteps == 1 && h->n > 0 && pos_in && vel_in && angVel_in && !getenv("WDEM_E2E_SEQUENTIAL")) {
    // pipelined: copies overlap with the integrator chunk by chunk (leapfrogPipelined); same results as the sequential path below
    return guarded(h, [&] {
      const size_t n = (size_t)h->n;
      ensureUp2Slot(h);
      { const double *i0 = h->ioIn.p, *o0 = h->ioOut.p; h->ioIn.ensure(9 * n, h->bytes); h->ioOut.ensure(6 * n, h->bytes);
        if (h->ioIn.p != i0 || h->ioOut.p != o0) h->e2e.ioBusy = false; }      // new staging buffers: the copy streams must be ordered behind their (stream-ordered) allocation
      h->e2e.defer = !stats && !h->profiling && !getenv("WDEM_E2E_NO_DEFER");
      h->e2e.count = (long long)n; h->e2e.slotMap = h->up2slot.p;
      h->e2e.active = true; h->e2e.pos = pos_in; h->e2e.vel = vel_in; h->e2e.angVel = angVel_in; h->e2e.posOut = pos_out; h->e2e.velOut = vel_out;
      try { stepN(h, 1, stats, true); } catch (...) { h->e2e.active = false; cudaStreamSynchronize(h->sH2D); cudaStreamSynchronize(h->sD2H); throw; }
      h->e2e.active = false;
      CK(cudaEventRecord(h->evEk, h->sD2H));
      CK(cudaStreamSynchronize(h->sD2H));
      if (getenv("WDEM_E2E_TRACE") && !h->e2e.defer) {   // device timeline of the call relative to its start (ms): upload / integrate per chunk, end of step, last byte on the host
        float a, b; std::string ln = "e2e trace:";
        for (int c = 0; c < h->nChunk; c++) { cudaEventElapsedTime(&a, h->ev0, h->evUp[c]); cudaEventElapsedTime(&b, h->ev0, h->evLf[c]); char t[64]; snprintf(t, sizeof(t), " [%.2f %.2f]", a, b); ln += t; }
        cudaEventElapsedTime(&a, h->ev0, h->ev1); cudaEventElapsedTime(&b, h->ev0, h->evEk);
        fprintf(stderr, "%s step_end %.2f d2h_end %.2f\n", ln.c_str(), a, b);
      }
      if (ek_out) { double e = 0; for (int b = 0; b < h->ekBlocks; b++) e += h->hEk[b]; *ek_out = e; }
    });
  }
  int rc = wdem_set_state(h, 0, h ? h->n : 0, pos_in, vel_in, angVel_in);
  if (rc) return rc;
  rc = wdem_step(h, nsteps, stats);
  if (rc) return rc;
  rc = wdem_download_state(h, pos_out, nullptr, vel_out, nullptr);
  if (rc) return rc;
  if (ek_out) rc = wdem_kinetic_energy(h, ek_out);
  return rc;
}

int64_t wdem_num_contacts(wdem_handle h, int realOnly) {
  long long nr = 0, na = 0;
  int rc = guarded(h, [&] { countContacts(h, nr, na); });
  if (rc) return -1;
  return realOnly ? nr : na;
}

int64_t wdem_download_contacts(wdem_handle h, int realOnly, int64_t cap, int32_t* ids, int32_t* cellDist, int32_t* real, double* uN, double* forceLocal, double* torqueLocal,
                               double* trsf, double* contPt, double* phys, int32_t* stepCreated, double* physRaw) {
  int64_t nOut = 0;
  int rc = guarded(h, [&] {
    const long long nC = h->nC; if (nC == 0) return;
    if (h->compactPending) throw EngErr{"spheres were removed since the last step: step once before reading contacts, wall forces or the stress (the table still refers to the old slots)", WDEM_ERR_INVALID};
    cudaStream_t st = h->stream; ContactStore& cs = h->con[h->curC];
    auto pair = d2h(cs.pair.p, nC, st); auto flags = d2h(cs.flags.p, nC, st); auto sid = d2h(h->S().id, h->n, st);
    std::vector<uint32_t> sfl; if (h->comm) sfl = d2h(h->S().flags, h->n, st);
    const bool hertz = h->cfg.law == WDEM_LAW_HERTZ_DMT, rolling = (h->cfg.relRollStiff > 0. && h->cfg.rollTanPhi > 0.);
    std::vector<double4> rv[4]; for (int k = 0; k < 4; k++) rv[k] = d2h(cs.r[k].p, nC, st);
    auto php = d2h(cs.ph.p, nC, st);
    std::vector<double> ph4v; if (hertz) ph4v = d2h(cs.ph4.p, nC, st);
    std::vector<double4> tqv; if (rolling) tqv = d2h(cs.tq.p, nC, st);
    auto sc = d2h(cs.stepCreated.p, nC, st);
    // views of the four state planes (wdem_kernels.cuh: struct Contacts)
    auto tcol = [&](int k, long long c) { const double4& v = rv[k / 3][c]; return (k % 3) == 0 ? v.x : ((k % 3) == 1 ? v.y : v.z); };
    auto fl = [&](int k, long long c) { return k == 0 ? rv[3][c].w : (k == 1 ? rv[0][c].w : rv[1][c].w); };
    auto phk = [&](int k, long long c) { const double4& v = php[c]; return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : (k == 3 ? v.w : ph4v[c]))); };
    auto tqk = [&](int k, long long c) { const double4& v = tqv[c]; return k == 0 ? v.x : (k == 1 ? v.y : v.z); };
    auto cp = [&](long long c) { return rv[2][c]; };
    std::vector<int32_t> cdv[3];
    if (h->cfg.periodic) for (int k = 0; k < 3; k++) cdv[k] = d2h(cs.cd[k].p, nC, st);
    struct Row { uint64_t key; long long c; };
    std::vector<Row> rows; rows.reserve(nC);
    for (long long c = 0; c < nC; c++) {
      if (flags[c] & CF_DEAD) continue;
      if (realOnly && !(flags[c] & CF_REAL)) continue;
      const int kd = cf_kind(flags[c]);
      if (h->comm && (sfl[kd == KIND_SS ? pair[c].x : pair[c].y] & F_GHOST)) continue;   // reported by the subdomain owning pA (sphere-sphere) / the sphere
      int a = (kd == KIND_SS) ? sid[pair[c].x] : (kd == KIND_WS ? h->hWalls[pair[c].x].id : h->hFacets[pair[c].x].id), b = sid[pair[c].y];
      rows.push_back(Row{((uint64_t)(uint32_t)a << 32) | (uint32_t)b, c});
    }
    std::sort(rows.begin(), rows.end(), [](const Row& x, const Row& y) { return x.key < y.key; });
    nOut = (int64_t)rows.size();
    if (cap < nOut) { if (cap > 0) throw EngErr{"wdem_download_contacts: capacity too small", WDEM_ERR_INVALID}; return; }
    for (int64_t i = 0; i < nOut; i++) {
      const long long c = rows[i].c; const bool isReal = flags[c] & CF_REAL;
      if (ids) { ids[2 * i] = (int32_t)(rows[i].key >> 32); ids[2 * i + 1] = (int32_t)(rows[i].key & 0xFFFFFFFFu); }
      if (cellDist) for (int k = 0; k < 3; k++) cellDist[3 * i + k] = h->cfg.periodic ? cdv[k][c] : 0;
      if (real) real[i] = isReal;
      if (uN) uN[i] = isReal ? cp(c).w : NAN;
      if (forceLocal) for (int k = 0; k < 3; k++) forceLocal[3 * i + k] = isReal ? fl(k, c) : 0.;
      if (torqueLocal) for (int k = 0; k < 3; k++) torqueLocal[3 * i + k] = (isReal && rolling) ? tqk(k, c) : 0.;
      if (trsf) {
        d3 c0 = mk3(tcol(0, c), tcol(1, c), tcol(2, c)), c1 = mk3(tcol(3, c), tcol(4, c), tcol(5, c)); d3 c2 = cross(c0, c1);
        double* o = trsf + 9 * i;
        if (isReal) { o[0] = c0.x; o[1] = c0.y; o[2] = c0.z; o[3] = c1.x; o[4] = c1.y; o[5] = c1.z; o[6] = c2.x; o[7] = c2.y; o[8] = c2.z; }
        else for (int k = 0; k < 9; k++) o[k] = (k % 4 == 0) ? 1. : 0.;
      }
      if (contPt) { contPt[3 * i] = isReal ? cp(c).x : NAN; contPt[3 * i + 1] = isReal ? cp(c).y : NAN; contPt[3 * i + 2] = isReal ? cp(c).z : NAN; }
      if (phys) {
        if (!isReal) { phys[3 * i] = phys[3 * i + 1] = phys[3 * i + 2] = NAN; }
        else if (!hertz) { phys[3 * i] = phk(0, c); phys[3 * i + 1] = phk(1, c); phys[3 * i + 2] = phk(2, c); }
        else { double u_ = cp(c).w; phys[3 * i] = (3 / 2.) * phk(0, c) * std::sqrt(-u_); phys[3 * i + 1] = phk(1, c) * std::sqrt(u_ < 0 ? -u_ : 0); phys[3 * i + 2] = phk(2, c); }
      }
      if (stepCreated) stepCreated[i] = sc[c];
      if (physRaw) for (int k = 0; k < 5; k++) physRaw[5 * i + k] = (isReal && k < (hertz ? 5 : 3)) ? phk(k, c) : 0.;
    }
  });
  return rc ? -1 : nOut;
}

int wdem_upload_contacts(wdem_handle h, int64_t n, const int32_t* ids, const int32_t* cellDist, const int32_t* stepCreated, const double* uN, const double* forceLocal,
                         const double* torqueLocal, const double* trsf, const double* contPt, const double* physRaw) {
  return guarded(h, [&] {
    if (n < 0 || (n > 0 && (!ids || !uN || !forceLocal || !trsf || !contPt || !physRaw))) throw EngErr{"wdem_upload_contacts: bad arguments", WDEM_ERR_INVALID};
    if (h->comm) throw EngErr{"wdem_upload_contacts: not available in decomposed runs", WDEM_ERR_UNSUPPORTED};
    if (!h->conDirty || h->cur != 0 || h->nC != 0) throw EngErr{"wdem_upload_contacts: call after wdem_upload_spheres/walls/facets and before the first step", WDEM_ERR_INVALID};
    const bool hertz = h->cfg.law == WDEM_LAW_HERTZ_DMT, rolling = (h->cfg.relRollStiff > 0. && h->cfg.rollTanPhi > 0.);
    std::unordered_map<int32_t, int> sph, wal, fac;
    for (size_t i = 0; i < h->hIds.size(); i++) sph[h->hIds[i]] = (int)i;
    for (size_t i = 0; i < h->hWalls.size(); i++) wal[h->hWalls[i].id] = (int)i;
    for (size_t i = 0; i < h->hFacets.size(); i++) fac[h->hFacets[i].id] = (int)i;
    std::vector<CRec> rows((size_t)n);
    for (int64_t i = 0; i < n; i++) {
      CRec r; memset(&r, 0, sizeof(r));
      const int32_t a = ids[2 * i], b = ids[2 * i + 1];
      auto sb = sph.find(b);
      if (sb == sph.end()) throw EngErr{"wdem_upload_contacts: contact ##" + std::to_string(a) + "+" + std::to_string(b) + ": id2 must be a sphere (canonical order: boundary first, lower sphere id first)", WDEM_ERR_INVALID};
      int kind, ia;
      auto sa = sph.find(a);
      if (sa != sph.end()) { if (!(a < b)) throw EngErr{"wdem_upload_contacts: sphere-sphere contacts must have id1<id2 (##" + std::to_string(a) + "+" + std::to_string(b) + ")", WDEM_ERR_INVALID}; kind = KIND_SS; ia = sa->second; }
      else if (wal.count(a)) { kind = KIND_WS; ia = wal[a]; }
      else if (fac.count(a)) { kind = KIND_FS; ia = fac[a]; }
      else throw EngErr{"wdem_upload_contacts: unknown particle id " + std::to_string(a), WDEM_ERR_INVALID};
      r.idA = a; r.idB = b; r.flags = CF_REAL | ((uint32_t)kind << CF_KIND_SHIFT); r.stepCreated = stepCreated ? stepCreated[i] : 0;
      for (int k = 0; k < 3; k++) r.cd[k] = cellDist ? cellDist[3 * i + k] : 0;
      r.pad = ia; r.pairY = sb->second;
      for (int k = 0; k < 6; k++) r.t[k] = trsf[9 * i + k];                       // columns 0 and 1 (column-major on the wire); column 2 = c0 x c1
      r.cp = make_double4(contPt[3 * i], contPt[3 * i + 1], contPt[3 * i + 2], uN[i]);
      for (int k = 0; k < 3; k++) { r.fl[k] = forceLocal[3 * i + k]; r.tq[k] = (torqueLocal && rolling) ? torqueLocal[3 * i + k] : 0.; }
      for (int k = 0; k < (hertz ? 5 : 3); k++) r.ph[k] = physRaw[5 * i + k];
      r.Fg = make_double4(0, 0, 0, 0);                                            // rewritten by the contact kernel before anybody reads it
      rows[(size_t)i] = r;
    }
    h->pendingRows.swap(rows);
  });
}

int wdem_set_scene(wdem_handle h, int64_t step, double time) {
  return guarded(h, [&] { if (step < 0) throw EngErr{"wdem_set_scene: negative step", WDEM_ERR_INVALID}; h->step = step; h->time = time; h->dynStepLast = h->dynStepPrev = -1; });
}

int wdem_crit_dt(wdem_handle h, double* dt) { return guarded(h, [&] { *dt = h->pwaveDt; }); }
int wdem_set_periodic_state(wdem_handle h, int64_t dynDtStepLast, int64_t dynDtStepPrev) {
  return guarded(h, [&] { if (dynDtStepLast < -1 || dynDtStepPrev < -1) throw EngErr{"wdem_set_periodic_state: bad step", WDEM_ERR_INVALID}; h->dynStepLast = dynDtStepLast; h->dynStepPrev = dynDtStepPrev; });
}
int wdem_get_periodic_state(wdem_handle h, int64_t* dynDtStepLast, int64_t* dynDtStepPrev) {
  return guarded(h, [&] { if (dynDtStepLast) *dynDtStepLast = h->dynStepLast; if (dynDtStepPrev) *dynDtStepPrev = h->dynStepPrev; });
}
int wdem_get_scalars(wdem_handle h, double* dt, double* time, int64_t* step, double* verletDist, int64_t* nFullRuns) {
  return guarded(h, [&] { if (dt) *dt = h->dt; if (time) *time = h->time; if (step) *step = h->step; if (verletDist) *verletDist = h->verletDist; if (nFullRuns) *nFullRuns = h->nFullRuns; });
}
// ---- slab decomposition
int wdem_comm_unique_id(char* out128) {
  if (!out128) return WDEM_ERR_INVALID;
  ncclUniqueId id; static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (ncclGetUniqueId(&id) != ncclSuccess) return WDEM_ERR_RUNTIME;
  memcpy(out128, &id, 128); return WDEM_OK;
}
int wdem_comm_init_nccl(wdem_handle h, int rank, int world, const char* id128) {
  return guarded(h, [&] {
    if (!id128 || world < 1 || world > WDEM_MAX_RANKS || rank < 0 || rank >= world) throw EngErr{"wdem_comm_init_nccl: bad arguments", WDEM_ERR_INVALID};
    if (h->comm) throw EngErr{"communicator already set", WDEM_ERR_INVALID};
    ncclUniqueId id; memcpy(&id, id128, 128);
    h->comm = new NcclTransport(rank, world, id);
  });
}
void* wdem_local_group_create(int world) { return (world >= 1 && world <= WDEM_MAX_RANKS) ? new LocalGroup(world) : nullptr; }
int wdem_local_group_destroy(void* g) { delete (LocalGroup*)g; return WDEM_OK; }
int wdem_comm_init_local(wdem_handle h, void* group, int rank) {
  return guarded(h, [&] {
    LocalGroup* G = (LocalGroup*)group;
    if (!G || rank < 0 || rank >= G->world) throw EngErr{"wdem_comm_init_local: bad arguments", WDEM_ERR_INVALID};
    if (h->comm) throw EngErr{"communicator already set", WDEM_ERR_INVALID};
    h->comm = new LocalTransport(G, rank);
  });
}
int wdem_set_domain(wdem_handle h, int axis, int ncuts, const double* cuts, int64_t globalMaxId) {
  return guarded(h, [&] {
    if (!h->comm) throw EngErr{"wdem_set_domain: initialise the communicator first", WDEM_ERR_INVALID};
    if (axis < 0 || axis > 2 || ncuts != h->comm->world - 1 || (ncuts > 0 && !cuts)) throw EngErr{"wdem_set_domain: need world-1 cuts along axis 0..2", WDEM_ERR_INVALID};
    for (int k = 0; k < ncuts; k++) if (!(cuts[k] == cuts[k]) || (k > 0 && cuts[k] < cuts[k - 1])) throw EngErr{"wdem_set_domain: cuts must be ascending", WDEM_ERR_INVALID};
    if (h->cfg.noBreak) throw EngErr{"Law2.noBreak is not supported in decomposed runs (separated real contacts are not exchanged)", WDEM_ERR_UNSUPPORTED};
    h->dom.axis = axis; h->dom.cuts.assign(cuts, cuts + ncuts); h->dom.globalMaxId = (int)globalMaxId; h->conDirty = true;
  });
}
int64_t wdem_num_owned(wdem_handle h) { return h ? (h->comm ? h->dom.nOwn : h->n) : -1; }
int wdem_set_rebalance(wdem_handle h, double tolerance, int minFullRuns) {
  return guarded(h, [&] {
    if (!h->comm) throw EngErr{"wdem_set_rebalance: initialise the communicator first", WDEM_ERR_INVALID};
    if (!(tolerance >= 0.) || minFullRuns < 1) throw EngErr{"wdem_set_rebalance: tolerance >= 0 (0 = off), minFullRuns >= 1", WDEM_ERR_INVALID};
    h->dom.rebalTol = tolerance; h->dom.rebalMinRuns = minFullRuns;
  });
}
int wdem_get_domain(wdem_handle h, double* cuts, int64_t* nRebalances) {
  return guarded(h, [&] {
    if (!h->comm) throw EngErr{"wdem_get_domain: not a decomposed run", WDEM_ERR_INVALID};
    if (cuts) for (size_t k = 0; k < h->dom.cuts.size(); k++) cuts[k] = h->dom.cuts[k];
    if (nRebalances) *nRebalances = h->dom.nRebalances;
  });
}
int wdem_download_owned(wdem_handle h, int32_t* ids, double* pos, double* ori, double* vel, double* angVel, double* force, double* torque) {
  return guarded(h, [&] {
    const int n = h->n; if (n == 0 || !ids) return;
    cudaStream_t st = h->stream;
    const int nOwn = h->comm ? h->dom.nOwn : n;
    h->ioBuf.ensure((size_t)n * 20 + 8, h->bytes); h->counts.ensure((size_t)n + 1, h->bytes); h->offsets.ensure((size_t)n + 1, h->bytes);
    k_owned_flag<<<nblk(n, 256), 256, 0, st>>>(h->S().flags, n, h->counts.p);
    exclusiveSum(h, h->counts.p, h->offsets.p, n);
    double* d = h->ioBuf.p; const size_t m = (size_t)nOwn;
    double *dpos = pos ? d : nullptr, *dori = ori ? d + 3 * m : nullptr, *dvel = vel ? d + 7 * m : nullptr, *dw = angVel ? d + 10 * m : nullptr, *df = force ? d + 13 * m : nullptr,
           *dt_ = torque ? d + 16 * m : nullptr;
    int32_t* dids = (int32_t*)(d + 19 * m);
    k_gather_owned<<<nblk(n, 256), 256, 0, st>>>(h->S(), n, h->offsets.p, dids, dpos, dori, dvel, dw, df, dt_);
    auto D2H = [&](void* dst, void* src, size_t bytes) { if (dst) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st)); };
    D2H(ids, dids, m * 4); D2H(pos, dpos, 24 * m); D2H(ori, dori, 32 * m); D2H(vel, dvel, 24 * m); D2H(angVel, dw, 24 * m); D2H(force, df, 24 * m); D2H(torque, dt_, 24 * m);
    CK(cudaStreamSynchronize(st));
  });
}
int wdem_set_owned(wdem_handle h, int64_t n, const int32_t* ids, const double* pos, const double* vel, const double* angVel) {
  return guarded(h, [&] {
    if (n < 0 || (n > 0 && !ids)) throw EngErr{"wdem_set_owned: bad arguments", WDEM_ERR_INVALID};
    if (n == 0) return;
    if (h->conDirty) throw EngErr{"wdem_set_owned: step at least once first (ownership is settled by the first collider run)", WDEM_ERR_INVALID};
    cudaStream_t st = h->stream;
    h->ioBuf.ensure((size_t)n * 10 + 8, h->bytes); h->orphCnt.ensure(1, h->bytes);
    double* d = h->ioBuf.p; int32_t* dids = (int32_t*)(d + 9 * (size_t)n);
    double *dp = nullptr, *dv = nullptr, *dw = nullptr;
    CK(cudaMemcpyAsync(dids, ids, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (pos) { dp = d; CK(cudaMemcpyAsync(dp, pos, (size_t)n * 24, cudaMemcpyHostToDevice, st)); }
    if (vel) { dv = d + 3 * (size_t)n; CK(cudaMemcpyAsync(dv, vel, (size_t)n * 24, cudaMemcpyHostToDevice, st)); }
    if (angVel) { dw = d + 6 * (size_t)n; CK(cudaMemcpyAsync(dw, angVel, (size_t)n * 24, cudaMemcpyHostToDevice, st)); }
    CK(cudaMemsetAsync(h->orphCnt.p, 0, sizeof(int), st));
    k_scatter_owned<<<nblk(n, 256), 256, 0, st>>>(h->S(), h->id2slot.p, dids, n, dp, dv, dw, h->orphCnt.p);
    int miss = 0; CK(cudaMemcpyAsync(&miss, h->orphCnt.p, sizeof(int), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    if (miss) throw EngErr{"wdem_set_owned: " + std::to_string(miss) + " ids are not owned by this subdomain", WDEM_ERR_INVALID};
  });
}
int wdem_domain_stats(wdem_handle h, int64_t* nOwned, int64_t* nGhosts, int64_t* migratedOut, int64_t* migratedIn, int64_t* rowsIn) {
  return guarded(h, [&] {
    if (nOwned) *nOwned = h->comm ? h->dom.nOwn : h->n; if (nGhosts) *nGhosts = h->comm ? h->n - h->dom.nOwn : 0;
    if (migratedOut) *migratedOut = h->dom.migratedOut; if (migratedIn) *migratedIn = h->dom.migratedIn; if (rowsIn) *rowsIn = h->dom.rowsIn;
  });
}

int wdem_peer_active(wdem_handle h) { return h ? ((h->comm && h->dom.peer.active) ? 1 : 0) : -1; }

int wdem_device_info(wdem_handle h, int64_t* bytes, void** stream) { return guarded(h, [&] { if (bytes) *bytes = h->bytes; if (stream) *stream = (void*)h->stream; }); }

}  // extern "C"
