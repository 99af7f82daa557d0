// woo_plugin.cpp — host side of the drop-in boundary, C++/pybind11 like the reference's own host code.
//
// * GpuDemLoop: one Engine that takes the slot of [Leapfrog, InsertionSortCollider, ContactLoop(, DynDt)] in Scene.engines
//   (reference: py/utils.py:122-129, core/Engine.hpp:19-104, Scene::doOneStep core/Scene.cpp:366-432).  It reads the attributes of the
//   reference's engine/functor objects it is constructed from, uploads the DemField through the C-ABI of include/wdem.h and steps on the
//   GPU.  C-ABI errors become std::runtime_error, which is what the reference's engines throw out of run().
// * A minimal object model with the reference's class and attribute names (Scene, Cell, DemField, Particle, Node, DemData, Sphere, Wall,
//   Facet, FrictMat, HertzMat, Contact, L6Geom, FrictPhys and the attribute-carrying engine/functor classes), enough for the reference's
//   scripts/tests of this path to run by renaming the import.  Inside real Woo the same GpuDemLoop body is registered with
//   WOO_PLUGIN(dem,(GpuDemLoop)) against Woo's own classes instead.
// There is no CPU implementation behind any of these classes: running Leapfrog/InsertionSortCollider/ContactLoop/DynDt on their own throws.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <list>
#include <set>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/wdem.h"

namespace py = pybind11;
using std::shared_ptr; using std::make_shared; using std::string; using std::vector;
typedef std::array<double, 3> V3; typedef std::array<int, 3> V3i; typedef std::array<double, 4> Q4; typedef std::array<double, 9> M3;
static const double NaN = std::numeric_limits<double>::quiet_NaN();

static py::array_t<double> np3(const V3& v) { py::array_t<double> a(3); std::memcpy(a.mutable_data(), v.data(), 24); return a; }
static py::array_t<double> np4(const Q4& v) { py::array_t<double> a(4); std::memcpy(a.mutable_data(), v.data(), 32); return a; }
static py::array_t<double> np33(const M3& m) { py::array_t<double> a({3, 3}); std::memcpy(a.mutable_data(), m.data(), 72); return a; }
static py::array_t<int> np3i(const V3i& v) { py::array_t<int> a(3); std::memcpy(a.mutable_data(), v.data(), 12); return a; }
template <size_t N> static std::array<double, N> seqN(const py::object& o, const char* what) {
  py::array_t<double, py::array::c_style | py::array::forcecast> a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(o);
  if (!a || (size_t)a.size() != N) throw std::invalid_argument(string(what) + ": expected " + std::to_string(N) + " numbers");
  std::array<double, N> r; std::memcpy(r.data(), a.data(), N * 8); return r;
}

struct DemField; struct Scene;

// ---- materials (dem/FrictMat.hpp:20-31, dem/Hertz.hpp:9-16)
struct Material { double density = 1e3; int id = -1; virtual ~Material() {} virtual int kind() const { return -1; } };
struct ElastMat : Material { double young = 1e9; };
struct FrictMat : ElastMat { double tanPhi = .5, ktDivKn = .2; int kind() const override { return WDEM_MAT_FRICT; } };
struct HertzMat : FrictMat { double surfEnergy = 0., alpha = 0.; int kind() const override { return WDEM_MAT_HERTZ; } };

// ---- nodes and particles (core/Field.hpp:102-103, dem/Particle.hpp:91-97,198-292)
enum { DOF_ALL = 63, GRAVITY_SKIP = 512, DAMPING_SKIP = 2048 };
struct Node;
// ---- Impose (dem/Particle.hpp:130-160, dem/Impose.hpp): attribute carriers; the device evaluates them (wdem_upload_imposes).  kind() = wdem_impose.kind
struct Impose : std::enable_shared_from_this<Impose> {
  virtual ~Impose() {} virtual int kind() const { return 0; }
  DemField* field = nullptr;                       // set when uploaded: attribute changes after that force a re-upload
  double sumF = 0, dist = 0; V3 F{{0, 0, 0}}, T{{0, 0, 0}};      // results read back by DemField::pull (VelocityAndReadForce.sumF/.dist, ReadForce.F/.T)
  void touch();
};
struct HarmonicOscillation : Impose { double freq = NaN, amp = NaN, t0 = 0; V3 dir{{1, 0, 0}}; bool perpFree = false; int kind() const override { return 1; } };
struct AlignedHarmonicOscillations : Impose { V3 freqs{{NaN, NaN, NaN}}, amps{{0, 0, 0}}; int kind() const override { return 2; } };
struct ConstantForce : Impose { V3 Fc{{0, 0, 0}}; int kind() const override { return 3; } };
struct RadialForce : Impose { shared_ptr<Node> nodeA, nodeB; double Fmag = 0; int kind() const override { return 4; } };
struct Local6Dofs : Impose { Q4 ori{{0, 0, 0, 1}}; std::array<double, 6> values{{0, 0, 0, 0, 0, 0}}; std::array<int, 6> whats{{0, 0, 0, 0, 0, 0}}; int kind() const override { return 5; } };
struct VelocityAndReadForce : Impose { V3 dir{{1, 0, 0}}; double vel = 0; bool latBlock = true, invF = false; string energyName; int kind() const override { return 6; } };
struct ReadForce : Impose { shared_ptr<Node> node; int kind() const override { return 7; } };
struct VariableVelocity3d : Impose { vector<double> times; vector<V3> vels, angVels; Q4 ori{{0, 0, 0, 1}}; vector<string> hooks; bool diff = false, wrap = false; double t0 = 0; int kind() const override { return 8; } };
struct VariableAlignedRotation : Impose { int axis = 0; vector<std::array<double, 2>> timeAngVel; bool wrap = false; int kind() const override { return 9; } };
struct CircularOrbit : Impose { shared_ptr<Node> node; bool rotate = false; double omega = NaN, angle = 0; int kind() const override { return 10; } };      // Impose.hpp (CircularOrbit)
struct StableCircularOrbit : CircularOrbit { double radius = NaN; int kind() const override { return 11; } };
struct InterpolatedMotion : Impose { vector<V3> poss; vector<Q4> oris; vector<double> times; double t0 = 0; int kind() const override { return 12; } };
struct CombinedImpose : Impose { vector<shared_ptr<Impose>> imps; };
struct DemData {
  virtual ~DemData() {}
  shared_ptr<Impose> impose;
  std::weak_ptr<Node> master;      // the clump's master node of a clumped node (DemData::master, Particle.hpp:233)
  V3 vel{{0, 0, 0}}, angVel{{0, 0, 0}}, force{{0, 0, 0}}, torque{{0, 0, 0}}, inertia{{0, 0, 0}}; double mass = 0; unsigned flags = 0;
  DemField* field = nullptr;
  // DemData::guessMoving (Particle.cpp:638-640): does the node need motion integration?
  bool guessMoving() const { return (mass != 0 && (flags & 63u) != 63u) || vel[0] != 0 || vel[1] != 0 || vel[2] != 0 || angVel[0] != 0 || angVel[1] != 0 || angVel[2] != 0 || bool(impose); }
  string blockedGet() const { string r; const char* n = "xyzXYZ"; for (int i = 0; i < 6; i++) if (flags & (1u << i)) r += n[i]; return r; }
  void blockedSet(const string& s) { flags &= ~63u; for (char c : s) { const char* n = "xyzXYZ"; const char* p = strchr(n, c); if (!p) throw std::invalid_argument("blocked: only xyzXYZ allowed"); flags |= 1u << (p - n); } }
};
struct Node { V3 pos{{0, 0, 0}}; Q4 ori{{0, 0, 0, 1}}; shared_ptr<DemData> dem; };
// master node data of a clump of spheres with mass (dem/Clump.hpp:60-100)
enum { CLUMP_CLUMPED = 64, CLUMP_CLUMP = 128 };
struct ClumpData : DemData { vector<shared_ptr<Node>> nodes; vector<V3> relPos; vector<Q4> relOri; V3 angMom{{NaN, NaN, NaN}}; double equivRad = NaN; };
struct Shape { vector<shared_ptr<Node>> nodes; virtual ~Shape() {} virtual int numNodes() const = 0; };
struct Sphere : Shape { double radius = NaN; int numNodes() const override { return 1; } };
struct Wall : Shape { int sense = 0, axis = 0; int numNodes() const override { return 1; } };
struct Facet : Shape { V3 fakeVel{{0, 0, 0}}; double halfThick = 0.; V3 n21lim{{NaN, NaN, NaN}}; int numNodes() const override { return 3; } };
struct Particle { int id = -1; unsigned mask = 1; shared_ptr<Shape> shape; shared_ptr<Material> material; DemField* field = nullptr; };

// ---- contacts (dem/Contact.hpp:21-106, dem/L6Geom.hpp:16-21, dem/FrictMat.hpp:33-35)
struct CGeomNode { V3 pos; };
struct L6Geom { double uN = NaN; M3 trsf; CGeomNode node; };
struct FrictPhys { V3 force, torque; double kn = NaN, kt = NaN, tanPhi = NaN; double raw[5] = {0, 0, 0, 0, 0}; /* the constants as the device stores them (wdem.h: physRaw) */ };
struct Contact { int id1 = -1, id2 = -1; V3i cellDist{{0, 0, 0}}; int stepCreated = -1; shared_ptr<L6Geom> geom; shared_ptr<FrictPhys> phys; bool isReal() const { return geom && phys; } };

struct GpuDemLoop;
struct ParticleContainer {
  DemField* dem = nullptr; vector<shared_ptr<Particle>> v; std::list<int> freeIds;
  int findFreeId() {      // ParticleContainer::findFreeId (ParticleContainer.cpp:23-36): freed ids are handed out again, oldest first
    const long size = (long)v.size();
    auto it = freeIds.begin();
    while (!freeIds.empty()) {
      const int id = *it; it = freeIds.erase(it);
      if (id <= size) { if (id < size && v[id]) throw std::logic_error("ParticleContainer::findFreeId: freeIds contained " + std::to_string(id) + ", but it is occupied?!"); return id; }
    }
    return (int)size;
  }
};
// clump of boundary particles moved rigidly by its master node (ParticleContainer::addClumped -> ClumpData::makeClump, Clump.cpp:187-261)
struct ClumpGroup { shared_ptr<Node> master; vector<int> members; /* Particle ids */ };
struct ContactContainer { DemField* dem = nullptr; vector<shared_ptr<Contact>> v; std::map<std::pair<int, int>, size_t> ix; };

// ---- the field (dem/Particle.hpp:339-361)
struct DemField {
  ParticleContainer par; ContactContainer con; vector<shared_ptr<Node>> nodes; vector<ClumpGroup> clumps;
  vector<shared_ptr<Node>> sphereClumps;      // master nodes carrying ClumpData (clumps of spheres with mass, integrated on the device)
  vector<int> pendingRemoves;                 // ids of spheres removed while the scene lives on the device (wdem_remove_spheres)
  vector<shared_ptr<Particle>> pendingAdds;   // spheres added while the scene lives on the device: appended there (wdem_add_spheres) instead of re-uploading everything
  V3 gravity{{0, 0, 0}}; unsigned loneMask = 2; double distFactor = -1.;
  // coherence with the device copy (SURVEY §7 "Host<->device coherence"): upload once, download lazily
  GpuDemLoop* owner = nullptr; bool topoDirty = true, stateDirty = false, deviceNewer = false, contactsStale = true;
  DemField() { par.dem = this; con.dem = this; }
  vector<shared_ptr<Impose>> imposeOf;        // Impose objects in upload order
  void pull(); void pullContacts(); void pushContacts(wdem_handle h);
  void touchState() { stateDirty = true; }
};

// ---- periodic cell (core/Cell.hpp)
struct Cell { M3 hSize{{1, 0, 0, 0, 1, 0, 0, 0, 1}}, gradV{{0, 0, 0, 0, 0, 0, 0, 0, 0}}, nextGradV{{0, 0, 0, 0, 0, 0, 0, 0, 0}}; int homoDeform = WDEM_HOMO_GRADV2; bool dirty = true; };

// ---- engines (core/Engine.hpp:19-104)
struct Engine {
  bool dead = false; string label; Scene* scene = nullptr; shared_ptr<DemField> field;
  virtual ~Engine() {}
  virtual void run() { throw std::runtime_error(string(cls()) + ".run(): woo_b200 has no CPU implementation of this engine; put it into GpuDemLoop(leapfrog, collider, contactLoop, dynDt)"); }
  virtual double critDt() { return std::numeric_limits<double>::infinity(); }
  virtual const char* cls() const { return "Engine"; }
};
struct Leapfrog : Engine { double damping = .2; bool reset = false, kinSplit = false, dontCollect = false; const char* cls() const override { return "Leapfrog"; } };
struct BoundFunctor { virtual ~BoundFunctor() {} };
struct Bo1_Sphere_Aabb : BoundFunctor {}; struct Bo1_Facet_Aabb : BoundFunctor {}; struct Bo1_Wall_Aabb : BoundFunctor { double halfThick = 0.; };
struct InsertionSortCollider : Engine { vector<shared_ptr<BoundFunctor>> boundDispatcher; double verletDist = -.05; long nFullRuns = 0; const char* cls() const override { return "InsertionSortCollider"; } };
struct CGeomFunctor { bool noRatch = false, iniLensTouch = true; int trsfRenorm = 100, approxMask = 0; virtual ~CGeomFunctor() {} };
struct Cg2_Sphere_Sphere_L6Geom : CGeomFunctor {}; struct Cg2_Wall_Sphere_L6Geom : CGeomFunctor {}; struct Cg2_Facet_Sphere_L6Geom : CGeomFunctor {};
struct CPhysFunctor { string label; virtual ~CPhysFunctor() {} };
struct Cp2_FrictMat_FrictPhys : CPhysFunctor { py::object tanPhi = py::none(); };
struct Cp2_HertzMat_HertzPhys : CPhysFunctor { double poisson = .2, en = NaN; };
struct LawFunctor { string label; virtual ~LawFunctor() {} };
struct Law2_L6Geom_FrictPhys_IdealElPl : LawFunctor { bool noSlip = false, noBreak = false, noFrict = false, iniEqlb = false; double relRollStiff = 0., relTwistStiff = 0., rollTanPhi = 0.; };
struct Law2_L6Geom_HertzPhys_DMT : LawFunctor { bool noAttraction = true; };
struct ContactLoop : Engine {
  vector<shared_ptr<CGeomFunctor>> geoDisp; vector<shared_ptr<CPhysFunctor>> phyDisp; vector<shared_ptr<LawFunctor>> lawDisp;
  bool applyForces = true, evalStress = false, dist00 = true; int updatePhys = 0, reorderEvery = 1000; py::object hook = py::none();
  M3 stress{{0, 0, 0, 0, 0, 0, 0, 0, 0}};   // ContactLoop.hpp:116-127: filled after every step when evalStress and the scene is periodic
  const char* cls() const override { return "ContactLoop"; }
};
struct DynDt : Engine { int stepPeriod = 100; double maxRelInc = 1e-4; bool dryRun = false; long stepLast = -1, stepPrev = -1;   // PeriodicEngine counters (Engine.hpp:156-157)
  const char* cls() const override { return "DynDt"; } };

// ---- Scene (core/Scene.hpp:191-205, Scene.cpp:366-432)
struct Scene {
  double dt = NaN, nextDt = NaN, dtSafety = .9, time = 0; long step = 0; bool periodic = false, trackEnergy = false, deterministic = false;
  shared_ptr<Cell> cell = make_shared<Cell>(); vector<py::object> engines; vector<shared_ptr<DemField>> fields;
  std::map<string, double> energy;   // Scene.energy (core/EnergyTracker.hpp) as far as this path feeds it; filled after every step when trackEnergy
  shared_ptr<DemField> dem() const { if (fields.empty()) throw std::runtime_error("Scene has no DemField"); return fields[0]; }
  void one(); void run(long n);
};

// ---- the B200 engine group
static void chk(wdem_handle h, int rc) { if (rc != WDEM_OK) throw std::runtime_error(string("GpuDemLoop: ") + wdem_last_error(h)); }

struct GpuDemLoop : Engine {
  shared_ptr<Leapfrog> leapfrog; shared_ptr<InsertionSortCollider> collider; shared_ptr<ContactLoop> contactLoop; shared_ptr<DynDt> dynDt;
  int device = 0; wdem_handle h = nullptr; wdem_config cfg; bool haveCfg = false; long nSpheres = 0; wdem_stats last;
  vector<int> sphereIds; vector<shared_ptr<Particle>> sphereOf, wallOf, facetOf; bool trackingOn = false, facetNodesMove = false;
  std::map<Material*, int> matIndex;   // materials as uploaded (by object identity)
  const char* cls() const override { return "GpuDemLoop"; }
  ~GpuDemLoop() override { release(); }
  void release() { if (h) { wdem_destroy(h); h = nullptr; } if (field && field->owner == this) field->owner = nullptr; }

  // attributes of the wrapped engines -> wdem_config (the reference engines re-read theirs at every run())
  wdem_config makeConfig() {
    if (!leapfrog || !collider || !contactLoop) throw std::runtime_error("GpuDemLoop needs leapfrog, collider and contactLoop");
    if (!scene || !field) throw std::runtime_error("GpuDemLoop is not attached to a Scene with a DemField");
    wdem_config c; wdem_default_config(&c);
    c.device = device; c.dt = scene->dt; c.dtSafety = scene->dtSafety; c.periodic = scene->periodic ? 1 : 0;
    for (int k = 0; k < 3; k++) c.gravity[k] = field->gravity[k];
    c.loneMask = field->loneMask; c.distFactor = field->distFactor;
    c.damping = leapfrog->damping; c.reset = leapfrog->reset ? 1 : 0;
    c.verletDist = collider->verletDist;
    for (auto& b : collider->boundDispatcher) if (auto w = dynamic_cast<Bo1_Wall_Aabb*>(b.get())) c.wallHalfThick = w->halfThick;
    if (contactLoop->geoDisp.empty() || contactLoop->phyDisp.size() != 1 || contactLoop->lawDisp.size() != 1) throw std::runtime_error("GpuDemLoop: ContactLoop needs Cg2 functors, exactly one Cp2 and one Law2 functor");
    const CGeomFunctor& g0 = *contactLoop->geoDisp[0];
    for (auto& g : contactLoop->geoDisp) {
      if (g->approxMask != 0) throw std::runtime_error("GpuDemLoop: Cg2_*.approxMask!=0 is not supported");
      if (g->noRatch != g0.noRatch || g->iniLensTouch != g0.iniLensTouch || g->trsfRenorm != g0.trsfRenorm) throw std::runtime_error("GpuDemLoop: all Cg2 functors must share noRatch/iniLensTouch/trsfRenorm");
    }
    c.noRatch = g0.noRatch; c.iniLensTouch = g0.iniLensTouch; c.trsfRenorm = g0.trsfRenorm;
    c.applyForces = contactLoop->applyForces;
    if (contactLoop->updatePhys != 0) throw std::runtime_error("GpuDemLoop: ContactLoop.updatePhys is not supported");
    if (!contactLoop->hook.is_none()) throw std::runtime_error("GpuDemLoop: ContactLoop.hook is not supported");
    c.evalStress = (contactLoop->evalStress && scene->periodic) ? 1 : 0;      // with Scene.trackEnergy: the "gradV" energy (ContactLoop.cpp:243-250)
    auto* lin = dynamic_cast<Law2_L6Geom_FrictPhys_IdealElPl*>(contactLoop->lawDisp[0].get());
    auto* dmt = dynamic_cast<Law2_L6Geom_HertzPhys_DMT*>(contactLoop->lawDisp[0].get());
    auto* cpF = dynamic_cast<Cp2_FrictMat_FrictPhys*>(contactLoop->phyDisp[0].get());
    auto* cpH = dynamic_cast<Cp2_HertzMat_HertzPhys*>(contactLoop->phyDisp[0].get());
    if (lin && cpF) {
      if (!cpF->tanPhi.is_none()) throw std::runtime_error("GpuDemLoop: Cp2_FrictMat_FrictPhys.tanPhi (MatchMaker) is not supported");
      if (lin->iniEqlb) throw std::runtime_error("GpuDemLoop: Law2_L6Geom_FrictPhys_IdealElPl.iniEqlb is not supported");
      c.law = WDEM_LAW_IDEALELPL; c.noSlip = lin->noSlip; c.noBreak = lin->noBreak; c.noFrict = lin->noFrict;
      c.relRollStiff = lin->relRollStiff; c.relTwistStiff = lin->relTwistStiff; c.rollTanPhi = lin->rollTanPhi;
    } else if (dmt && cpH) {
      c.law = WDEM_LAW_HERTZ_DMT; c.noAttraction = dmt->noAttraction; c.poisson = cpH->poisson; c.en = cpH->en;
    } else throw std::runtime_error("GpuDemLoop: supported functor pairs are Cp2_FrictMat_FrictPhys+Law2_L6Geom_FrictPhys_IdealElPl and Cp2_HertzMat_HertzPhys+Law2_L6Geom_HertzPhys_DMT (FUNCTOR2D dispatch of the reference, Hertz.hpp:54,77)");
    if (dynDt && !dynDt->dead) { if (dynDt->dryRun) throw std::runtime_error("GpuDemLoop: DynDt.dryRun is not supported"); c.dynDtPeriod = dynDt->stepPeriod; c.maxRelInc = dynDt->maxRelInc; }
    else c.dynDtPeriod = 0;
    return c;
  }
  static bool sameCfg(const wdem_config& a, const wdem_config& b) {
    if (std::memcmp(&a, &b, sizeof(wdem_config)) == 0) return true;
    wdem_config x = a, y = b; if (std::isnan(x.dt) && std::isnan(y.dt)) x.dt = y.dt = 0; if (std::isnan(x.en) && std::isnan(y.en)) x.en = y.en = 0;
    return std::memcmp(&x, &y, sizeof(wdem_config)) == 0;
  }

  // DemData.impose -> wdem_upload_imposes: one wdem_impose per distinct Impose object, one set per distinct DemData.impose (CombinedImpose = a set of several)
  void uploadImposes() {
    DemField& F = *field; F.imposeOf.clear();
    vector<wdem_impose> imps; vector<vector<double>> keep; keep.reserve(64); std::map<Impose*, int> ixOf, setOf; vector<int32_t> setStart{0}, setOps;
    auto unit = [](const V3& d, double* o) { const double n = std::sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]); for (int k = 0; k < 3; k++) o[k] = d[k] / n; };      // postLoad: dir.normalize()
    std::function<int(const shared_ptr<Impose>&)> one = [&](const shared_ptr<Impose>& I) -> int {
      auto it = ixOf.find(I.get()); if (it != ixOf.end()) return it->second;
      wdem_impose w; std::memset(&w, 0, sizeof(w)); w.ori[3] = 1; w.kind = I->kind();
      if (auto* p = dynamic_cast<HarmonicOscillation*>(I.get())) { w.flags = p->perpFree ? 1 : 0; w.v[0] = p->freq; w.v[1] = p->amp; unit(p->dir, w.v + 2); w.v[5] = p->t0; }
      else if (auto* p = dynamic_cast<AlignedHarmonicOscillations*>(I.get())) { for (int k = 0; k < 3; k++) { w.v[k] = p->freqs[k]; w.v[3 + k] = p->amps[k]; } }
      else if (auto* p = dynamic_cast<ConstantForce*>(I.get())) { for (int k = 0; k < 3; k++) w.v[k] = p->Fc[k]; }
      else if (auto* p = dynamic_cast<RadialForce*>(I.get())) { if (!p->nodeA || !p->nodeB) throw std::runtime_error("RadialForce: nodeA and nodeB must not be None."); for (int k = 0; k < 3; k++) { w.v[k] = p->nodeA->pos[k]; w.v[3 + k] = p->nodeB->pos[k]; } w.v[6] = p->Fmag; }
      else if (auto* p = dynamic_cast<Local6Dofs*>(I.get())) { for (int k = 0; k < 6; k++) { w.v[k] = p->values[k]; w.v[6 + k] = p->whats[k]; } for (int k = 0; k < 4; k++) w.ori[k] = p->ori[k]; }
      else if (auto* p = dynamic_cast<VelocityAndReadForce*>(I.get())) { if (!p->energyName.empty()) throw std::runtime_error("GpuDemLoop: VelocityAndReadForce.energyName is not supported"); w.flags = (p->latBlock ? 1 : 0) | (p->invF ? 2 : 0); unit(p->dir, w.v); w.v[3] = p->vel; w.v[4] = p->dist; }
      else if (auto* p = dynamic_cast<ReadForce*>(I.get())) { if (p->node) { w.flags = 1; for (int k = 0; k < 3; k++) w.v[k] = p->node->pos[k]; for (int k = 0; k < 4; k++) w.ori[k] = p->node->ori[k]; } }
      else if (auto* p = dynamic_cast<VariableVelocity3d*>(I.get())) {
        for (auto& hk : p->hooks) if (!hk.empty()) throw std::runtime_error("GpuDemLoop: VariableVelocity3d.hooks are not supported");
        if (p->times.size() != p->vels.size()) throw std::runtime_error("VariableVelocity3d: times and vels must have the same length (not " + std::to_string(p->times.size()) + ", " + std::to_string(p->vels.size()) + ").");
        if (!p->angVels.empty() && p->angVels.size() != p->times.size()) throw std::runtime_error("VariableVelocity3d.angVels: if not empty, it length (" + std::to_string(p->angVels.size()) + " must be equal to the length of times (" + std::to_string(p->times.size()) + ").");
        w.flags = (p->diff ? 1 : 0) | (p->wrap ? 2 : 0); w.v[0] = p->t0; for (int k = 0; k < 4; k++) w.ori[k] = p->ori[k]; w.nTimes = (int)p->times.size();
        keep.push_back(p->times); w.times = keep.back().data();
        keep.emplace_back(); for (auto& x : p->vels) for (int k = 0; k < 3; k++) keep.back().push_back(x[k]); w.vals = keep.back().data();
        if (!p->angVels.empty()) { keep.emplace_back(); for (auto& x : p->angVels) for (int k = 0; k < 3; k++) keep.back().push_back(x[k]); w.vals2 = keep.back().data(); }
      }
      else if (auto* p = dynamic_cast<VariableAlignedRotation*>(I.get())) {
        w.flags = p->wrap ? 1 : 0; w.v[0] = p->axis; w.nTimes = (int)p->timeAngVel.size();
        keep.emplace_back(); for (auto& x : p->timeAngVel) keep.back().push_back(x[0]); w.times = keep.back().data();
        keep.emplace_back(); for (auto& x : p->timeAngVel) keep.back().push_back(x[1]); w.vals = keep.back().data();
      }
      else if (auto* p = dynamic_cast<InterpolatedMotion*>(I.get())) {
        if (p->poss.size() != p->oris.size() || p->oris.size() != p->times.size()) throw std::runtime_error("InterpolatedMotion: poss, oris, times must have the same length (not " + std::to_string(p->poss.size()) + ", " + std::to_string(p->oris.size()) + ", " + std::to_string(p->times.size()) + ")");      // postLoad (Impose.cpp:115-118)
        w.v[0] = p->t0; w.nTimes = (int)p->times.size();
        keep.push_back(p->times); w.times = keep.back().data();
        keep.emplace_back(); for (auto& x : p->poss) for (int k = 0; k < 3; k++) keep.back().push_back(x[k]); w.vals = keep.back().data();
        keep.emplace_back(); for (auto& x : p->oris) for (int k = 0; k < 4; k++) keep.back().push_back(x[k]); w.vals2 = keep.back().data();
      }
      else if (auto* p = dynamic_cast<CircularOrbit*>(I.get())) {      // also StableCircularOrbit
        if (!p->node) throw std::runtime_error(string(w.kind == 11 ? "StableCircularOrbit" : "CircularOrbit") + ": node must not be None.");
        w.flags = p->rotate ? 1 : 0; for (int k = 0; k < 3; k++) w.v[k] = p->node->pos[k]; for (int k = 0; k < 4; k++) w.ori[k] = p->node->ori[k];
        w.v[3] = p->omega; w.v[5] = p->angle; if (auto* sp = dynamic_cast<StableCircularOrbit*>(p)) w.v[4] = sp->radius;
      }
      else throw std::runtime_error("GpuDemLoop: this Impose class is not supported on the device path");
      I->field = field.get(); imps.push_back(w); F.imposeOf.push_back(I); return ixOf[I.get()] = (int)imps.size() - 1;
    };
    auto setFor = [&](const shared_ptr<Impose>& I) -> int {
      auto it = setOf.find(I.get()); if (it != setOf.end()) return it->second;
      if (auto* c = dynamic_cast<CombinedImpose*>(I.get())) { for (auto& i : c->imps) setOps.push_back(one(i)); I->field = field.get(); } else setOps.push_back(one(I));
      setStart.push_back((int32_t)setOps.size()); return setOf[I.get()] = (int)setStart.size() - 2;
    };
    vector<int32_t> sid, sset, wset(wallOf.size(), -1), gset(F.clumps.size(), -1); bool anyWall = false, anyGroup = false;
    for (auto& p : sphereOf) { auto& imp = p->shape->nodes[0]->dem->impose; if (imp) { sid.push_back(p->id); sset.push_back(setFor(imp)); } }
    for (size_t i = 0; i < wallOf.size(); i++) { auto& imp = wallOf[i]->shape->nodes[0]->dem->impose; if (imp) { wset[i] = setFor(imp); anyWall = true; } }
    for (size_t g = 0; g < F.clumps.size(); g++) { auto& imp = F.clumps[g].master->dem->impose; if (imp) { gset[g] = setFor(imp); anyGroup = true; } }
    vector<int32_t> cset(F.sphereClumps.size(), -1); bool anyClump = false;
    for (size_t c = 0; c < F.sphereClumps.size(); c++) { auto& imp = F.sphereClumps[c]->dem->impose; if (imp) { cset[c] = setFor(imp); anyClump = true; } }
    // facet nodes (py/triangulated.py:59-64: every node of a cylinder carries the same StableCircularOrbit); facets of a kinematic group follow their master node
    vector<int32_t> fset(facetOf.size(), -1); bool anyFacet = false;
    for (size_t i = 0; i < facetOf.size(); i++) {
      auto& nn = facetOf[i]->shape->nodes; auto& imp = nn[0]->dem->impose;
      if (nn[1]->dem->impose != imp || nn[2]->dem->impose != imp) throw std::runtime_error("GpuDemLoop: the three nodes of facet #" + std::to_string(facetOf[i]->id) + " must carry the same DemData.impose");
      if (imp) { fset[i] = setFor(imp); anyFacet = true; }
    }
    if (imps.empty()) return;
    chk(h, wdem_upload_imposes(h, (int32_t)imps.size(), imps.data(), (int32_t)setStart.size() - 1, setStart.data(), setOps.data(), (int64_t)sid.size(), sid.data(), sset.data(), anyWall ? wset.data() : nullptr, anyGroup ? gset.data() : nullptr));
    if (anyClump) chk(h, wdem_upload_clump_imposes(h, cset.data()));
    if (anyFacet) chk(h, wdem_upload_facet_imposes(h, fset.data()));
    facetNodesMove = anyFacet;
  }
  // DemField -> device (SoA host staging; ids are Particle.id)
  void upload() {
    DemField& F = *field;
    if (F.owner && F.owner != this) throw std::runtime_error("GpuDemLoop: the DemField is already attached to another GpuDemLoop");
    if (h && F.owner == this && F.deviceNewer) F.pull();   // a re-upload (particles added, masses changed, ...) starts from the latest device state
    // ... and keeps the contact history: the reference's containers survive ParticleContainer::add/remove (ParticleContainer.cpp:23-73,121-157),
    // so the real contacts are read back and imported again behind the new particle set (contacts of removed particles are dropped)
    if (h && F.owner == this) F.pullContacts();
    wdem_config c = makeConfig();
    if (h && c.law != cfg.law) { wdem_destroy(h); h = nullptr; F.con.v.clear(); F.con.ix.clear(); }   // another law: the stored phys constants mean nothing to it
    if (!h) { int rc = wdem_create(&c, &h); if (rc != WDEM_OK) { h = nullptr; throw std::runtime_error(string("GpuDemLoop: ") + wdem_last_error(nullptr)); } }
    else chk(h, wdem_set_config(h, &c));
    cfg = c; haveCfg = true;
    // materials, de-duplicated by object identity
    vector<Material*> mats; std::map<Material*, int> matIx;
    auto matOf = [&](const shared_ptr<Particle>& p) { Material* m = p->material.get(); if (!m) throw std::runtime_error("particle #" + std::to_string(p->id) + " has no material"); auto it = matIx.find(m); if (it != matIx.end()) return it->second; int k = (int)mats.size(); mats.push_back(m); matIx[m] = k; return k; };
    vector<int32_t> sid, smat, wid, waxis, wsense, wmat, fid, fmat; vector<uint32_t> sflags, smask, wmask, fmask;
    vector<double> spos, sori, svel, sang, smass, sin, srad, sfrc, strq, wpos, wvel, fv, fvv, ffk, fht, fnl;
    sphereIds.clear(); sphereOf.clear(); wallOf.clear(); facetOf.clear();
    for (auto& p : F.par.v) {
      if (!p) continue;
      if (!p->shape) throw std::runtime_error("particle #" + std::to_string(p->id) + " has no shape");
      const int m = matOf(p);
      if (auto* s = dynamic_cast<Sphere*>(p->shape.get())) {
        const Node& n = *s->nodes.at(0); const DemData& d = *n.dem;
        if (!(d.inertia[0] == d.inertia[1] && d.inertia[1] == d.inertia[2])) throw std::runtime_error("GpuDemLoop: aspherical inertia is not supported (particle #" + std::to_string(p->id) + ")");
        sid.push_back(p->id); smat.push_back(m); sflags.push_back(d.flags); smask.push_back(p->mask);
        for (int k = 0; k < 3; k++) { spos.push_back(n.pos[k]); svel.push_back(d.vel[k]); sang.push_back(d.angVel[k]); sfrc.push_back(d.force[k]); strq.push_back(d.torque[k]); }
        for (int k = 0; k < 4; k++) sori.push_back(n.ori[k]);
        smass.push_back(d.mass); sin.push_back(d.inertia[0]); srad.push_back(s->radius);
        sphereIds.push_back(p->id); sphereOf.push_back(p);
      } else if (auto* w = dynamic_cast<Wall*>(p->shape.get())) {
        const Node& n = *w->nodes.at(0);
        if (n.dem->angVel[0] != 0 || n.dem->angVel[1] != 0 || n.dem->angVel[2] != 0) throw std::runtime_error("GpuDemLoop: rotating walls are not supported");
        wid.push_back(p->id); waxis.push_back(w->axis); wsense.push_back(w->sense); wmat.push_back(m); wmask.push_back(p->mask);
        for (int k = 0; k < 3; k++) { wpos.push_back(n.pos[k]); wvel.push_back(n.dem->vel[k]); }
        wallOf.push_back(p);
      } else if (auto* f = dynamic_cast<Facet*>(p->shape.get())) {
        fid.push_back(p->id); fmat.push_back(m); fmask.push_back(p->mask); fht.push_back(f->halfThick);
        for (int v = 0; v < 3; v++) { const Node& n = *f->nodes.at(v); for (int k = 0; k < 3; k++) { fv.push_back(n.pos[k]); fvv.push_back(n.dem->vel[k]); } }
        for (int k = 0; k < 3; k++) { ffk.push_back(f->fakeVel[k]); fnl.push_back(f->n21lim[k]); }
        facetOf.push_back(p);
      } else throw std::runtime_error("GpuDemLoop: particle #" + std::to_string(p->id) + ": only Sphere, Wall and Facet shapes are supported");
    }
    if (mats.empty()) throw std::runtime_error("GpuDemLoop: no particles");
    vector<int32_t> mk; vector<double> md, my, mkt, mtp, mse, mal;
    for (Material* m : mats) {
      auto* fm = dynamic_cast<FrictMat*>(m);
      if (!fm) throw std::runtime_error("GpuDemLoop: materials must be FrictMat or HertzMat");
      auto* hm = dynamic_cast<HertzMat*>(m);
      mk.push_back(m->kind()); md.push_back(m->density); my.push_back(fm->young); mkt.push_back(fm->ktDivKn); mtp.push_back(fm->tanPhi); mse.push_back(hm ? hm->surfEnergy : 0.); mal.push_back(hm ? hm->alpha : 0.);
    }
    chk(h, wdem_upload_materials(h, (int)mats.size(), mk.data(), md.data(), my.data(), mkt.data(), mtp.data(), mse.data(), mal.data()));
    chk(h, wdem_upload_spheres(h, (int64_t)sid.size(), sid.data(), spos.data(), sori.data(), svel.data(), sang.data(), smass.data(), sin.data(), srad.data(), sflags.data(), smask.data(), smat.data()));
    chk(h, wdem_upload_forces(h, sfrc.data(), strq.data()));
    chk(h, wdem_upload_walls(h, (int)wid.size(), wid.data(), wpos.data(), waxis.data(), wsense.data(), wvel.data(), wmask.data(), wmat.data()));
    chk(h, wdem_upload_facets(h, (int)fid.size(), fid.data(), fv.data(), fvv.data(), ffk.data(), fht.data(), fnl.data(), fmask.data(), fmat.data()));
    if (!F.clumps.empty()) {   // kinematic groups of facets: master node pos/ori/vel/angVel, members follow on the device
      std::map<int, int> facetIx; for (size_t i = 0; i < facetOf.size(); i++) facetIx[facetOf[i]->id] = (int)i;
      vector<int32_t> of(facetOf.size(), -1); vector<double> gp, go, gv, gw;
      for (size_t g = 0; g < F.clumps.size(); g++) {
        const Node& M = *F.clumps[g].master;
        if ((M.dem->flags & DOF_ALL) != DOF_ALL) throw std::runtime_error("GpuDemLoop: the master node of a clump of boundary particles must have all DOFs blocked (its motion is prescribed)");
        for (int id : F.clumps[g].members) { auto it = facetIx.find(id); if (it == facetIx.end()) throw std::runtime_error("GpuDemLoop: only Facets can be members of a kinematic clump (particle #" + std::to_string(id) + ")"); of[it->second] = (int32_t)g; }
        for (int k = 0; k < 3; k++) { gp.push_back(M.pos[k]); gv.push_back(M.dem->vel[k]); gw.push_back(M.dem->angVel[k]); }
        for (int k = 0; k < 4; k++) go.push_back(M.ori[k]);
      }
      chk(h, wdem_upload_facet_groups(h, (int)F.clumps.size(), of.data(), gp.data(), go.data(), gv.data(), gw.data()));
    }
    if (!F.sphereClumps.empty()) {   // clumps of spheres with mass (ClumpData): wdem_upload_clumps
      std::map<const Node*, int> idOfNode; for (auto& p : sphereOf) idOfNode[p->shape->nodes[0].get()] = p->id;
      vector<int32_t> cStart{0}, cMem; vector<uint32_t> cFl; vector<double> cp, co, cv, cw, cl, cm, cI, cf, ct, crp, cro;
      for (auto& n : F.sphereClumps) {
        const ClumpData& cd = static_cast<const ClumpData&>(*n->dem);
        for (size_t i = 0; i < cd.nodes.size(); i++) {
          auto it = idOfNode.find(cd.nodes[i].get()); if (it == idOfNode.end()) throw std::runtime_error("GpuDemLoop: a member of a clump is not (or no longer) a Sphere of this DemField");
          cMem.push_back(it->second);
          for (int k = 0; k < 3; k++) crp.push_back(cd.relPos[i][k]);
          for (int k = 0; k < 4; k++) cro.push_back(cd.relOri[i][k]);
        }
        cStart.push_back((int32_t)cMem.size()); cFl.push_back(cd.flags & ~(unsigned)CLUMP_CLUMP); cm.push_back(cd.mass);
        for (int k = 0; k < 3; k++) { cp.push_back(n->pos[k]); cv.push_back(cd.vel[k]); cw.push_back(cd.angVel[k]); cl.push_back(cd.angMom[k]); cI.push_back(cd.inertia[k]); cf.push_back(cd.force[k]); ct.push_back(cd.torque[k]); }
        for (int k = 0; k < 4; k++) co.push_back(n->ori[k]);
      }
      chk(h, wdem_upload_clumps(h, (int32_t)F.sphereClumps.size(), cStart.data(), cMem.data(), cp.data(), co.data(), cv.data(), cw.data(), cl.data(), cm.data(), cI.data(), cFl.data(), crp.data(), cro.data()));
      chk(h, wdem_upload_clump_forces(h, cf.data(), ct.data()));
    }
    uploadImposes();
    if (scene->periodic) { Cell& C = *scene->cell; chk(h, wdem_set_cell(h, C.hSize.data(), C.gradV.data(), C.nextGradV.data(), C.homoDeform)); C.dirty = false; }
    chk(h, wdem_set_scene(h, (int64_t)scene->step, scene->time));
    if (dynDt && !dynDt->dead) chk(h, wdem_set_periodic_state(h, dynDt->stepLast, dynDt->stepPrev));      // a re-upload in mid-run does not restart DynDt's period
    F.pushContacts(h);
    nSpheres = (long)sid.size(); matIndex = matIx;
    F.owner = this; F.topoDirty = false; F.stateDirty = false; F.deviceNewer = false; F.contactsStale = true; F.pendingAdds.clear(); F.pendingRemoves.clear();
  }
  // spheres removed since the upload (ParticleContainer::remove): drop them on the device; the rest keeps state, forces and contact history
  void removeSpheres() {
    DemField& F = *field;
    if (F.stateDirty) { if (F.deviceNewer) throw std::runtime_error("GpuDemLoop: host state edited before the device state was read back"); pushState(); }
    if (F.deviceNewer) F.pull();     // the host copies must be current before the arrays they map to close up
    vector<int32_t> ids(F.pendingRemoves.begin(), F.pendingRemoves.end());
    chk(h, wdem_remove_spheres(h, (int64_t)ids.size(), ids.data()));
    std::set<int> gone(ids.begin(), ids.end());
    vector<int> sid2; vector<shared_ptr<Particle>> sof2;
    for (size_t i = 0; i < sphereOf.size(); i++) if (!gone.count(sphereIds[i])) { sid2.push_back(sphereIds[i]); sof2.push_back(sphereOf[i]); }
    sphereIds.swap(sid2); sphereOf.swap(sof2); nSpheres = (long)sphereOf.size(); F.pendingRemoves.clear(); F.contactsStale = true;
  }
  // spheres added since the upload: append them on the device (ParticleContainer::add in mid-run; the contact history and everything else stays where it is)
  bool appendSpheres() {
    DemField& F = *field;
    vector<int32_t> sid, smat; vector<uint32_t> sflags, smask; vector<double> spos, sori, svel, sang, smass, sin, srad;
    for (auto& p : F.pendingAdds) {
      auto* s = dynamic_cast<Sphere*>(p->shape.get()); auto it = matIndex.find(p->material.get());
      if (!s || it == matIndex.end()) return false;                                 // a new material (or not a sphere): take the full path
      const Node& n = *s->nodes.at(0); const DemData& d = *n.dem;
      if (!(d.inertia[0] == d.inertia[1] && d.inertia[1] == d.inertia[2])) return false;
      sid.push_back(p->id); smat.push_back(it->second); sflags.push_back(d.flags); smask.push_back(p->mask);
      for (int k = 0; k < 3; k++) { spos.push_back(n.pos[k]); svel.push_back(d.vel[k]); sang.push_back(d.angVel[k]); }
      for (int k = 0; k < 4; k++) sori.push_back(n.ori[k]);
      smass.push_back(d.mass); sin.push_back(d.inertia[0]); srad.push_back(s->radius);
    }
    if (F.stateDirty) { if (F.deviceNewer) throw std::runtime_error("GpuDemLoop: host state edited before the device state was read back"); pushState(); }
    chk(h, wdem_add_spheres(h, (int64_t)sid.size(), sid.data(), spos.data(), sori.data(), svel.data(), sang.data(), smass.data(), sin.data(), srad.data(), sflags.data(), smask.data(), smat.data()));
    for (auto& p : F.pendingAdds) { sphereIds.push_back(p->id); sphereOf.push_back(p); }
    nSpheres = (long)sphereOf.size(); F.pendingAdds.clear(); F.contactsStale = true;
    return true;
  }
  void pushState() {   // host edits of pos/vel/angVel since the last run
    DemField& F = *field; const size_t n = sphereOf.size();
    vector<double> pos(3 * n), vel(3 * n), ang(3 * n);
    for (size_t i = 0; i < n; i++) { const Node& nd = *sphereOf[i]->shape->nodes[0]; for (int k = 0; k < 3; k++) { pos[3 * i + k] = nd.pos[k]; vel[3 * i + k] = nd.dem->vel[k]; ang[3 * i + k] = nd.dem->angVel[k]; } }
    if (n) chk(h, wdem_set_state(h, 0, (int64_t)n, pos.data(), vel.data(), ang.data()));
    F.stateDirty = false;
  }
  void prepare() {
    if (!field) throw std::runtime_error("GpuDemLoop: no DemField");
    DemField& F = *field;
    if (F.stateDirty && (!F.clumps.empty() || !F.sphereClumps.empty())) F.topoDirty = true;   // the edit may concern a clump's master node (its vel/angVel): take the full path
    if (!F.sphereClumps.empty() && (!F.pendingAdds.empty() || !F.pendingRemoves.empty())) F.topoDirty = true;   // no incremental add/remove beside clumps (wdem.h)
    if (h && !F.topoDirty && F.owner == this && !F.pendingRemoves.empty()) removeSpheres();
    if (h && !F.topoDirty && F.owner == this && !F.pendingAdds.empty() && !appendSpheres()) F.topoDirty = true;
    if (!h || F.topoDirty || F.owner != this) upload();
    else {
      wdem_config c = makeConfig();
      if (!sameCfg(c, cfg)) { chk(h, wdem_set_config(h, &c)); cfg = c; }
      if (F.stateDirty) { if (F.deviceNewer) throw std::runtime_error("GpuDemLoop: host state edited before the device state was read back"); pushState(); }
      if (scene->periodic && scene->cell->dirty) { Cell& C = *scene->cell; chk(h, wdem_set_next_gradv(h, C.nextGradV.data())); C.dirty = false; }
    }
  }
  double critDt() override { prepare(); double d; chk(h, wdem_crit_dt(h, &d)); return d; }   // DemField::critDt -> pWaveDt (Funcs.cpp:151-175)
  void step(long n) {
    prepare();
    if (scene->trackEnergy && !trackingOn) { chk(h, wdem_track_energy(h, 1)); trackingOn = true; }
    std::memset(&last, 0, sizeof(last));
    { py::gil_scoped_release nogil; chk(h, wdem_step(h, n, &last)); }
    DemField& F = *field; F.deviceNewer = true; F.contactsStale = true;
    int64_t nf = 0; double libDt = NaN;
    chk(h, wdem_get_scalars(h, &libDt, nullptr, nullptr, nullptr, &nf));
    collider->nFullRuns = (long)nf;
    if (dynDt && !dynDt->dead) { int64_t a = -1, b = -1; chk(h, wdem_get_periodic_state(h, &a, &b)); dynDt->stepLast = (long)a; dynDt->stepPrev = (long)b; }
    if (scene->periodic) chk(h, wdem_get_cell(h, scene->cell->hSize.data()));
    chk(h, wdem_track_energy(h, scene->trackEnergy ? 1 : 0));   // takes effect from the next step on, like flipping Scene.trackEnergy between steps
    if (scene->trackEnergy && trackingOn) {
      double e[8]; chk(h, wdem_get_energy(h, e));
      static const char* names[8] = {"kinetic", "grav", "nonviscDamp", "elast", "plast", "broken", "viscN", "viscT"};
      // EnergyTracker semantics: "grav" appears as soon as gravity acts, the ZeroDontCreate entries only once they are non-zero (EnergyTracker.cpp:26-34)
      for (int k = 0; k < 8; k++) { const bool always = (k == 0 || k == 3); if (always || e[k] != 0. || scene->energy.count(names[k])) scene->energy[names[k]] = e[k]; }
      if (cfg.evalStress) { double w = 0; chk(h, wdem_get_gradv_work(h, &w)); if (w != 0. || scene->energy.count("gradV")) scene->energy["gradV"] = w; }      // IsIncrement | ZeroDontCreate
    }
    trackingOn = scene->trackEnergy;
    if (contactLoop->evalStress && scene->periodic) chk(h, wdem_get_stress(h, contactLoop->stress.data()));   // doStress (ContactLoop.cpp:123)
    else contactLoop->stress.fill(0.);
    cfg.dt = libDt;                        // what the library will use next; Scene.dt follows through nextDt in the epilogue
    if (libDt != scene->dt) scene->nextDt = libDt;
  }
  void run() override { step(1); }
};

void Impose::touch() { if (field) field->topoDirty = true; }
void DemField::pull() {
  if (!deviceNewer || !owner || !owner->h) return;
  const size_t n = owner->sphereOf.size();
  vector<double> pos(3 * n), ori(4 * n), vel(3 * n), ang(3 * n), f(3 * n), t(3 * n);
  if (n) { chk(owner->h, wdem_download_state(owner->h, pos.data(), ori.data(), vel.data(), ang.data())); chk(owner->h, wdem_download_forces(owner->h, f.data(), t.data())); }
  for (size_t i = 0; i < n; i++) {
    Node& nd = *owner->sphereOf[i]->shape->nodes[0]; DemData& d = *nd.dem;
    for (int k = 0; k < 3; k++) { nd.pos[k] = pos[3 * i + k]; d.vel[k] = vel[3 * i + k]; d.angVel[k] = ang[3 * i + k]; d.force[k] = f[3 * i + k]; d.torque[k] = t[3 * i + k]; }
    for (int k = 0; k < 4; k++) nd.ori[k] = ori[4 * i + k];
  }
  // walls: position (a wall added with a velocity is moved by Leapfrog) and the force/torque of the last step's contacts (wdem.h: wdem_download_walls)
  const size_t nw = owner->wallOf.size();
  if (nw) {
    vector<double> wp(3 * nw), wf(3 * nw), wt(3 * nw);
    chk(owner->h, wdem_download_walls(owner->h, wp.data(), wf.data(), wt.data()));
    for (size_t i = 0; i < nw; i++) { Node& nd = *owner->wallOf[i]->shape->nodes[0]; for (int k = 0; k < 3; k++) { nd.pos[k] = wp[3 * i + k]; nd.dem->force[k] = wf[3 * i + k]; nd.dem->torque[k] = wt[3 * i + k]; } }
  }
  if (!imposeOf.empty()) {      // boundary nodes whose velocity an imposition changes
    const size_t nwl = owner->wallOf.size(), ng = clumps.size(); vector<double> wv(3 * nwl + 3), gv(3 * ng + 3), gw(3 * ng + 3);
    chk(owner->h, wdem_download_boundary_velocities(owner->h, wv.data(), owner->facetOf.empty() ? nullptr : gv.data(), owner->facetOf.empty() ? nullptr : gw.data()));
    for (size_t i = 0; i < nwl; i++) for (int k = 0; k < 3; k++) owner->wallOf[i]->shape->nodes[0]->dem->vel[k] = wv[3 * i + k];
    if (!owner->facetOf.empty()) for (size_t g = 0; g < ng; g++) for (int k = 0; k < 3; k++) { clumps[g].master->dem->vel[k] = gv[3 * g + k]; clumps[g].master->dem->angVel[k] = gw[3 * g + k]; }
  }
  for (size_t i = 0; i < imposeOf.size(); i++) {      // what the last step's Leapfrog read / accumulated
    double o[8]; chk(owner->h, wdem_get_impose_state(owner->h, (int32_t)i, o));
    Impose& I = *imposeOf[i]; I.sumF = o[0]; I.dist = o[1]; for (int k = 0; k < 3; k++) { I.F[k] = o[2 + k]; I.T[k] = o[5 + k]; }
    if (auto* co = dynamic_cast<CircularOrbit*>(&I)) co->angle = o[1];
  }
  if (!sphereClumps.empty()) {
    const size_t nc = sphereClumps.size(); vector<double> p(3 * nc), o(4 * nc), v(3 * nc), w(3 * nc), l(3 * nc), cf(3 * nc), ct(3 * nc);
    chk(owner->h, wdem_download_clumps(owner->h, p.data(), o.data(), v.data(), w.data(), l.data(), cf.data(), ct.data()));
    for (size_t i = 0; i < nc; i++) {
      Node& nd = *sphereClumps[i]; ClumpData& d = static_cast<ClumpData&>(*nd.dem);
      for (int k = 0; k < 3; k++) { nd.pos[k] = p[3 * i + k]; d.vel[k] = v[3 * i + k]; d.angVel[k] = w[3 * i + k]; d.angMom[k] = l[3 * i + k]; d.force[k] = cf[3 * i + k]; d.torque[k] = ct[3 * i + k]; }
      for (int k = 0; k < 4; k++) nd.ori[k] = o[4 * i + k];
    }
  }
  if ((!clumps.empty() || owner->facetNodesMove) && !owner->facetOf.empty()) {
    const size_t nf = owner->facetOf.size(); vector<double> fv(9 * nf), gp(3 * clumps.size() + 3), go(4 * clumps.size() + 4);
    chk(owner->h, wdem_download_facets(owner->h, fv.data(), gp.data(), go.data()));
    for (size_t i = 0; i < nf; i++) for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) owner->facetOf[i]->shape->nodes[v]->pos[k] = fv[9 * i + 3 * v + k];
    for (size_t g = 0; g < clumps.size(); g++) { Node& M = *clumps[g].master; for (int k = 0; k < 3; k++) M.pos[k] = gp[3 * g + k]; for (int k = 0; k < 4; k++) M.ori[k] = go[4 * g + k]; }
  }
  deviceNewer = false;
}
void DemField::pullContacts() {
  if (!contactsStale || !owner || !owner->h) return;
  wdem_handle h = owner->h;
  int64_t n = wdem_download_contacts(h, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (n < 0) chk(h, 1);
  vector<int32_t> ids(2 * n), cd(3 * n), real(n), sc(n); vector<double> uN(n), f(3 * n), t(3 * n), trsf(9 * n), cp(3 * n), ph(3 * n), raw(5 * n);
  if (n) { int64_t m = wdem_download_contacts(h, 0, n, ids.data(), cd.data(), real.data(), uN.data(), f.data(), t.data(), trsf.data(), cp.data(), ph.data(), sc.data(), raw.data()); if (m < 0) chk(h, 1); }
  con.v.clear(); con.ix.clear();
  for (int64_t i = 0; i < n; i++) {
    auto C = make_shared<Contact>(); C->id1 = ids[2 * i]; C->id2 = ids[2 * i + 1]; C->stepCreated = sc[i];
    for (int k = 0; k < 3; k++) C->cellDist[k] = cd[3 * i + k];
    if (real[i]) {
      C->geom = make_shared<L6Geom>(); C->phys = make_shared<FrictPhys>(); C->geom->uN = uN[i];
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) C->geom->trsf[3 * r + c] = trsf[9 * i + 3 * c + r];   // column-major on the wire
      for (int k = 0; k < 3; k++) { C->geom->node.pos[k] = cp[3 * i + k]; C->phys->force[k] = f[3 * i + k]; C->phys->torque[k] = t[3 * i + k]; }
      C->phys->kn = ph[3 * i]; C->phys->kt = ph[3 * i + 1]; C->phys->tanPhi = ph[3 * i + 2];
      for (int k = 0; k < 5; k++) C->phys->raw[k] = raw[5 * i + k];
    }
    con.ix[{std::min(C->id1, C->id2), std::max(C->id1, C->id2)}] = con.v.size(); con.v.push_back(C);
  }
  contactsStale = false;
}

// real contacts of the host container -> device (after a re-upload of the particle set); contacts whose particles no longer exist are dropped,
// as ParticleContainer::remove deletes the contacts of a removed particle (ParticleContainer.cpp:121-157)
void DemField::pushContacts(wdem_handle h) {
  vector<int32_t> ids, cd, sc; vector<double> uN, f, t, trsf, cp, raw;
  auto alive = [&](int id) { return id >= 0 && (size_t)id < par.v.size() && par.v[id]; };
  for (auto& C : con.v) {
    if (!C || !C->isReal() || !alive(C->id1) || !alive(C->id2)) continue;
    ids.push_back(C->id1); ids.push_back(C->id2); sc.push_back(C->stepCreated); uN.push_back(C->geom->uN);
    for (int k = 0; k < 3; k++) { cd.push_back(C->cellDist[k]); f.push_back(C->phys->force[k]); t.push_back(C->phys->torque[k]); cp.push_back(C->geom->node.pos[k]); }
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) trsf.push_back(C->geom->trsf[3 * r + c]);   // column-major on the wire
    for (int k = 0; k < 5; k++) raw.push_back(C->phys->raw[k]);
  }
  if (!sc.empty()) chk(h, wdem_upload_contacts(h, (int64_t)sc.size(), ids.data(), cd.data(), sc.data(), uN.data(), f.data(), t.data(), trsf.data(), cp.data(), raw.data()));
  contactsStale = true;
}

// Scene::doOneStep (Scene.cpp:366-432) for an engine list whose DEM group is a GpuDemLoop; other entries may be Python callables f(scene)
void Scene::one() {
  GpuDemLoop* gpu = nullptr;
  for (auto& o : engines) {
    if (py::isinstance<Engine>(o)) { Engine& e = o.cast<Engine&>(); e.scene = this; if (!e.field && !fields.empty()) e.field = fields[0]; if (auto g = dynamic_cast<GpuDemLoop*>(&e)) gpu = g; }
  }
  if (std::isnan(dt) || std::isinf(dt) || dt < 0) {   // Scene.cpp:380-398
    double crit = std::numeric_limits<double>::infinity();
    for (auto& o : engines) if (py::isinstance<Engine>(o)) { Engine& e = o.cast<Engine&>(); if (!e.dead) crit = std::min(crit, e.critDt()); }
    if (std::isinf(crit)) throw std::runtime_error("Failed to obtain meaningful dt from engines and fields automatically");
    dt = dtSafety * crit;
  }
  for (auto& o : engines) {
    if (py::isinstance<Engine>(o)) { Engine& e = o.cast<Engine&>(); if (!e.dead) e.run(); }
    else o(py::cast(this, py::return_value_policy::reference));
  }
  (void)gpu;
  step++; time += dt;
  if (!std::isnan(nextDt)) { dt = nextDt; nextDt = NaN; }
}
void Scene::run(long n) {
  // fast path: the GPU group is the only engine -> one C-ABI call for all n steps (no per-step host round trip)
  if (n > 1 && engines.size() == 1 && py::isinstance<GpuDemLoop>(engines[0]) && !(std::isnan(dt) || std::isinf(dt) || dt < 0)) {
    GpuDemLoop& g = engines[0].cast<GpuDemLoop&>();
    if (!g.dead) {
      g.scene = this; if (!g.field && !fields.empty()) g.field = fields[0];
      g.step(n);
      double libDt, libTime; int64_t libStep; chk(g.h, wdem_get_scalars(g.h, &libDt, &libTime, &libStep, nullptr, nullptr));
      step += n; time = libTime; dt = libDt; nextDt = NaN;
      return;
    }
  }
  for (long i = 0; i < n; i++) one();
}

// ---- factories (Particle::make dem/Particle.cpp:51-69; Sphere::lumpMassInertia dem/Sphere.cpp:26-33)
static shared_ptr<Material> defaultMaterial() { auto m = make_shared<FrictMat>(); m->density = 1e3; m->young = 1e7; m->ktDivKn = .2; m->tanPhi = std::tan(.5); return m; }
static shared_ptr<Particle> makeParticle(const shared_ptr<Shape>& sh, shared_ptr<Material> mat, bool fixed, unsigned mask) {
  auto p = make_shared<Particle>(); p->shape = sh; p->material = mat ? mat : defaultMaterial(); p->mask = mask;
  sh->nodes.resize(sh->numNodes());
  for (auto& n : sh->nodes) { if (!n) n = make_shared<Node>(); if (!n->dem) n->dem = make_shared<DemData>(); if (fixed) n->dem->flags |= DOF_ALL; }
  if (auto* s = dynamic_cast<Sphere*>(sh.get())) {
    DemData& d = *sh->nodes[0]->dem; const double r = s->radius;
    d.mass = (4 / 3.) * M_PI * (r * r * r) * p->material->density; const double I = (2. / 5.) * d.mass * (r * r); d.inertia = {{I, I, I}};
  }
  return p;
}

template <class T, class... Bases> static py::class_<T, Bases..., shared_ptr<T>> kwclass(py::module_& m, const char* name) {
  // keyword constructor like the reference's Object(**kw) (lib/object/Object.hpp:353-367): T(attr=value, ...)
  py::class_<T, Bases..., shared_ptr<T>> c(m, name);
  c.def(py::init([](py::kwargs kw) { auto o = make_shared<T>(); py::object po = py::cast(o); for (auto it : kw) py::setattr(po, it.first, it.second); return o; }));
  return c;
}
#define RW(cls, a) .def_readwrite(#a, &cls::a)

PYBIND11_MODULE(_woo, m) {
  m.doc() = "woo_b200: reference-named host objects + GpuDemLoop over the wdem C-ABI";
  // ---- materials
  kwclass<Material>(m, "Material") RW(Material, density) RW(Material, id);
  kwclass<ElastMat, Material>(m, "ElastMat") RW(ElastMat, young);
  kwclass<FrictMat, ElastMat>(m, "FrictMat") RW(FrictMat, tanPhi) RW(FrictMat, ktDivKn);
  kwclass<HertzMat, FrictMat>(m, "HertzMat") RW(HertzMat, surfEnergy) RW(HertzMat, alpha);
  // ---- nodes
#define V3P(cls, pyname, expr, PULL, TOUCH) .def_property(pyname, [](cls& s) { PULL; return np3(expr); }, [](cls& s, py::object o) { PULL; expr = seqN<3>(o, pyname); TOUCH; })
  kwclass<DemData>(m, "DemData")
      V3P(DemData, "vel", s.vel, if (s.field) s.field->pull(), if (s.field) s.field->touchState())
      V3P(DemData, "angVel", s.angVel, if (s.field) s.field->pull(), if (s.field) s.field->touchState())
      V3P(DemData, "force", s.force, if (s.field) s.field->pull(), if (s.field) s.field->topoDirty = true)
      V3P(DemData, "torque", s.torque, if (s.field) s.field->pull(), if (s.field) s.field->topoDirty = true)
      V3P(DemData, "inertia", s.inertia, (void)0, if (s.field) s.field->topoDirty = true)
      .def_property("mass", [](DemData& s) { return s.mass; }, [](DemData& s, double v) { s.mass = v; if (s.field) s.field->topoDirty = true; })
      .def_property("flags", [](DemData& s) { return s.flags; }, [](DemData& s, unsigned v) { s.flags = v; if (s.field) s.field->topoDirty = true; })
      .def_property("blocked", &DemData::blockedGet, [](DemData& s, const string& v) { s.blockedSet(v); if (s.field) s.field->topoDirty = true; })
      .def_property("gravitySkip", [](DemData& s) { return bool(s.flags & GRAVITY_SKIP); }, [](DemData& s, bool b) { s.flags = b ? (s.flags | GRAVITY_SKIP) : (s.flags & ~GRAVITY_SKIP); if (s.field) s.field->topoDirty = true; })
      .def("guessMoving", &DemData::guessMoving)
      .def_property("impose", [](DemData& s) { return s.impose; }, [](DemData& s, shared_ptr<Impose> i) { s.impose = i; if (s.field) s.field->topoDirty = true; })
      .def_property_readonly("clump", [](DemData& s) { return bool(s.flags & CLUMP_CLUMP); }).def_property_readonly("clumped", [](DemData& s) { return bool(s.flags & CLUMP_CLUMPED); })
      .def_property_readonly("master", [](DemData& s) { return s.master.lock(); })
      .def_property("dampingSkip", [](DemData& s) { return bool(s.flags & DAMPING_SKIP); }, [](DemData& s, bool b) { s.flags = b ? (s.flags | DAMPING_SKIP) : (s.flags & ~DAMPING_SKIP); if (s.field) s.field->topoDirty = true; });
  // ---- Impose classes (dem/Impose.hpp); attribute changes after the upload trigger a re-upload (which keeps particle state and contact history)
#define RWI(cls, a) .def_property(#a, [](cls& s) { return s.a; }, [](cls& s, decltype(cls::a) v) { s.a = v; s.touch(); })
#define V3I(cls, pyname, mem) .def_property(pyname, [](cls& s) { return np3(s.mem); }, [](cls& s, py::object o) { s.mem = seqN<3>(o, pyname); s.touch(); })
  py::class_<Impose, shared_ptr<Impose>>(m, "Impose")
      .def("__add__", [](shared_ptr<Impose> a, shared_ptr<Impose> b) {      // Impose::pyAdd (Impose.cpp:302-312)
        auto c = make_shared<CombinedImpose>();
        if (auto* ca = dynamic_cast<CombinedImpose*>(a.get())) c->imps = ca->imps; else c->imps.push_back(a);
        if (auto* cb = dynamic_cast<CombinedImpose*>(b.get())) c->imps.insert(c->imps.end(), cb->imps.begin(), cb->imps.end()); else c->imps.push_back(b);
        return c; });
  kwclass<HarmonicOscillation, Impose>(m, "HarmonicOscillation") RWI(HarmonicOscillation, freq) RWI(HarmonicOscillation, amp) RWI(HarmonicOscillation, t0) RWI(HarmonicOscillation, perpFree) V3I(HarmonicOscillation, "dir", dir);
  kwclass<AlignedHarmonicOscillations, Impose>(m, "AlignedHarmonicOscillations") V3I(AlignedHarmonicOscillations, "freqs", freqs) V3I(AlignedHarmonicOscillations, "amps", amps);
  kwclass<ConstantForce, Impose>(m, "ConstantForce") V3I(ConstantForce, "F", Fc);
  kwclass<RadialForce, Impose>(m, "RadialForce") RWI(RadialForce, nodeA) RWI(RadialForce, nodeB).def_property("F", [](RadialForce& s) { return s.Fmag; }, [](RadialForce& s, double v) { s.Fmag = v; s.touch(); });
  kwclass<Local6Dofs, Impose>(m, "Local6Dofs")
      .def_property("ori", [](Local6Dofs& s) { return np4(s.ori); }, [](Local6Dofs& s, py::object o) { s.ori = seqN<4>(o, "ori"); s.touch(); })
      .def_property("values", [](Local6Dofs& s) { return s.values; }, [](Local6Dofs& s, py::object o) { s.values = seqN<6>(o, "values"); s.touch(); })
      .def_property("whats", [](Local6Dofs& s) { return s.whats; }, [](Local6Dofs& s, std::array<int, 6> w) {
        for (int i = 0; i < 6; i++) if (w[i] != 0 && w[i] != 1 && w[i] != 2) throw std::runtime_error("Local6Dofs.whats components must be 0, 2 or 1 (whats[" + std::to_string(i) + "] invalid)");      // Impose.hpp:96-103
        s.whats = w; s.touch(); });
  kwclass<VelocityAndReadForce, Impose>(m, "VelocityAndReadForce") V3I(VelocityAndReadForce, "dir", dir) RWI(VelocityAndReadForce, vel) RWI(VelocityAndReadForce, latBlock) RWI(VelocityAndReadForce, invF) RWI(VelocityAndReadForce, energyName)
      .def_property("dist", [](VelocityAndReadForce& s) { if (s.field) s.field->pull(); return s.dist; }, [](VelocityAndReadForce& s, double v) { if (s.field) s.field->pull(); s.dist = v; s.touch(); })
      .def_property_readonly("sumF", [](VelocityAndReadForce& s) { if (s.field) s.field->pull(); return s.sumF; });
  kwclass<ReadForce, Impose>(m, "ReadForce") RWI(ReadForce, node)
      .def_property_readonly("F", [](ReadForce& s) { if (s.field) s.field->pull(); return np3(s.F); }).def_property_readonly("T", [](ReadForce& s) { if (s.field) s.field->pull(); return np3(s.T); });
  kwclass<VariableVelocity3d, Impose>(m, "VariableVelocity3d") RWI(VariableVelocity3d, times) RWI(VariableVelocity3d, hooks) RWI(VariableVelocity3d, diff) RWI(VariableVelocity3d, wrap) RWI(VariableVelocity3d, t0)
      .def_property("ori", [](VariableVelocity3d& s) { return np4(s.ori); }, [](VariableVelocity3d& s, py::object o) { s.ori = seqN<4>(o, "ori"); s.touch(); })
      .def_property("vels", [](VariableVelocity3d& s) { py::list l; for (auto& v : s.vels) l.append(np3(v)); return l; }, [](VariableVelocity3d& s, py::iterable o) { s.vels.clear(); for (auto it : o) s.vels.push_back(seqN<3>(py::reinterpret_borrow<py::object>(it), "vels")); s.touch(); })
      .def_property("angVels", [](VariableVelocity3d& s) { py::list l; for (auto& v : s.angVels) l.append(np3(v)); return l; }, [](VariableVelocity3d& s, py::iterable o) { s.angVels.clear(); for (auto it : o) s.angVels.push_back(seqN<3>(py::reinterpret_borrow<py::object>(it), "angVels")); s.touch(); });
  kwclass<InterpolatedMotion, Impose>(m, "InterpolatedMotion") RWI(InterpolatedMotion, times) RWI(InterpolatedMotion, t0)
      .def_property("poss", [](InterpolatedMotion& s) { py::list l; for (auto& v : s.poss) l.append(np3(v)); return l; }, [](InterpolatedMotion& s, py::iterable o) { s.poss.clear(); for (auto it : o) s.poss.push_back(seqN<3>(py::reinterpret_borrow<py::object>(it), "poss")); s.touch(); })
      .def_property("oris", [](InterpolatedMotion& s) { py::list l; for (auto& v : s.oris) l.append(np4(v)); return l; }, [](InterpolatedMotion& s, py::iterable o) { s.oris.clear(); for (auto it : o) s.oris.push_back(seqN<4>(py::reinterpret_borrow<py::object>(it), "oris")); s.touch(); });
  kwclass<CircularOrbit, Impose>(m, "CircularOrbit") RWI(CircularOrbit, node) RWI(CircularOrbit, rotate) RWI(CircularOrbit, omega)
      .def_property("angle", [](CircularOrbit& s) { if (s.field) s.field->pull(); return s.angle; }, [](CircularOrbit& s, double v) { if (s.field) s.field->pull(); s.angle = v; s.touch(); });
  kwclass<StableCircularOrbit, CircularOrbit>(m, "StableCircularOrbit") RWI(StableCircularOrbit, radius);
  kwclass<VariableAlignedRotation, Impose>(m, "VariableAlignedRotation") RWI(VariableAlignedRotation, axis) RWI(VariableAlignedRotation, wrap)
      .def_property("timeAngVel", [](VariableAlignedRotation& s) { return s.timeAngVel; }, [](VariableAlignedRotation& s, py::iterable o) {
        s.timeAngVel.clear(); for (auto it : o) s.timeAngVel.push_back(seqN<2>(py::reinterpret_borrow<py::object>(it), "timeAngVel"));
        for (size_t i = 0; i + 1 < s.timeAngVel.size(); i++) if (!(s.timeAngVel[i + 1][0] >= s.timeAngVel[i][0])) throw py::value_error("VariableAlignedRotation.timeAngVel: time values must be non-decreasing");      // Impose.cpp:89-93
        s.touch(); });
  py::class_<CombinedImpose, Impose, shared_ptr<CombinedImpose>>(m, "CombinedImpose").def_readonly("imps", &CombinedImpose::imps);
  py::class_<ClumpData, DemData, shared_ptr<ClumpData>>(m, "ClumpData")
      .def_readonly("nodes", &ClumpData::nodes).def_readonly("equivRad", &ClumpData::equivRad)
      .def_property_readonly("relPos", [](ClumpData& s) { py::list l; for (auto& v : s.relPos) l.append(np3(v)); return l; })
      .def_property_readonly("relOri", [](ClumpData& s) { py::list l; for (auto& v : s.relOri) l.append(np4(v)); return l; })
      V3P(ClumpData, "angMom", s.angMom, if (s.field) s.field->pull(), if (s.field) s.field->topoDirty = true);
  kwclass<Node>(m, "Node")
      V3P(Node, "pos", s.pos, if (s.dem && s.dem->field) s.dem->field->pull(), if (s.dem && s.dem->field) s.dem->field->touchState())
      .def_property("ori", [](Node& s) { if (s.dem && s.dem->field) s.dem->field->pull(); return np4(s.ori); }, [](Node& s, py::object o) { s.ori = seqN<4>(o, "ori"); if (s.dem && s.dem->field) s.dem->field->topoDirty = true; })
      .def_readwrite("dem", &Node::dem);
  // ---- shapes and particles
  py::class_<Shape, shared_ptr<Shape>>(m, "Shape").def_readwrite("nodes", &Shape::nodes);
  kwclass<Sphere, Shape>(m, "Sphere") RW(Sphere, radius)
      .def_static("make", [](py::object center, double radius, shared_ptr<Material> mat, unsigned mask, bool fixed) {
        auto s = make_shared<Sphere>(); s->radius = radius; auto p = makeParticle(s, mat, fixed, mask); s->nodes[0]->pos = seqN<3>(center, "center"); return p; },
        py::arg("center"), py::arg("radius"), py::arg("mat") = py::none(), py::arg("mask") = 5u, py::arg("fixed") = false);
  kwclass<Wall, Shape>(m, "Wall") RW(Wall, sense) RW(Wall, axis)
      .def_static("make", [](py::object position, int axis, int sense, shared_ptr<Material> mat, unsigned mask, bool fixed) {
        auto w = make_shared<Wall>(); w->axis = axis; w->sense = sense; auto p = makeParticle(w, mat, fixed, mask);
        if (py::isinstance<py::float_>(position) || py::isinstance<py::int_>(position)) { V3 x{{0, 0, 0}}; x[axis] = position.cast<double>(); w->nodes[0]->pos = x; }
        else w->nodes[0]->pos = seqN<3>(position, "position");
        return p; },
        py::arg("position"), py::arg("axis"), py::arg("sense") = 0, py::arg("mat") = py::none(), py::arg("mask") = 3u, py::arg("fixed") = true);
  kwclass<Facet, Shape>(m, "Facet") RW(Facet, halfThick)
      .def_property("fakeVel", [](Facet& f) { return np3(f.fakeVel); }, [](Facet& f, py::object o) { f.fakeVel = seqN<3>(o, "fakeVel"); })
      .def_property("n21lim", [](Facet& f) { return np3(f.n21lim); }, [](Facet& f, py::object o) { f.n21lim = seqN<3>(o, "n21lim"); })
      .def_static("make", [](py::object vertices, shared_ptr<Material> mat, double halfThick, py::object fakeVel, unsigned mask, bool fixed) {
        auto f = make_shared<Facet>(); f->halfThick = halfThick; if (!fakeVel.is_none()) f->fakeVel = seqN<3>(fakeVel, "fakeVel");
        auto p = makeParticle(f, mat, fixed, mask); auto vv = seqN<9>(vertices, "vertices");
        for (int v = 0; v < 3; v++) f->nodes[v]->pos = {{vv[3 * v], vv[3 * v + 1], vv[3 * v + 2]}};
        return p; },
        py::arg("vertices"), py::arg("mat") = py::none(), py::arg("halfThick") = 0., py::arg("fakeVel") = py::none(), py::arg("mask") = 3u, py::arg("fixed") = true);
#define NODE0(p) (*(p).shape->nodes.at(0))
  kwclass<Particle>(m, "Particle") RW(Particle, id) RW(Particle, mask) RW(Particle, shape) RW(Particle, material)
      .def_property("impose", [](Particle& p) { return NODE0(p).dem->impose; }, [](Particle& p, shared_ptr<Impose> i) { for (auto& n : p.shape->nodes) n->dem->impose = i; if (p.field) p.field->topoDirty = true; })      // Particle::setImpose (Particle.cpp)
      .def_property("mat", [](Particle& p) { return p.material; }, [](Particle& p, shared_ptr<Material> mm) { p.material = mm; if (p.field) p.field->topoDirty = true; })
      V3P(Particle, "pos", NODE0(s).pos, if (s.field) s.field->pull(), if (s.field) s.field->touchState())
      V3P(Particle, "vel", NODE0(s).dem->vel, if (s.field) s.field->pull(), if (s.field) s.field->touchState())
      V3P(Particle, "angVel", NODE0(s).dem->angVel, if (s.field) s.field->pull(), if (s.field) s.field->touchState())
      .def_property_readonly("ori", [](Particle& p) { if (p.field) p.field->pull(); return np4(NODE0(p).ori); })
      .def_property_readonly("f", [](Particle& p) { if (p.field) p.field->pull(); return np3(NODE0(p).dem->force); })
      .def_property_readonly("t", [](Particle& p) { if (p.field) p.field->pull(); return np3(NODE0(p).dem->torque); })
      .def_property_readonly("mass", [](Particle& p) { return NODE0(p).dem->mass; })
      .def_property_readonly("inertia", [](Particle& p) { return np3(NODE0(p).dem->inertia); })
      .def_property("blocked", [](Particle& p) { return NODE0(p).dem->blockedGet(); }, [](Particle& p, const string& b) { NODE0(p).dem->blockedSet(b); if (p.field) p.field->topoDirty = true; })
      .def_property_readonly("Ek", [](Particle& p) {   // DemData::getEk_any, aperiodic (Particle.cpp:247-265)
        if (p.field) p.field->pull(); const DemData& d = *NODE0(p).dem; double e = .5 * d.mass * (d.vel[0] * d.vel[0] + d.vel[1] * d.vel[1] + d.vel[2] * d.vel[2]);
        for (int k = 0; k < 3; k++) e += .5 * d.inertia[k] * d.angVel[k] * d.angVel[k]; return e; })
      .def_static("make", [](shared_ptr<Shape> sh, shared_ptr<Material> mat, bool fixed) { return makeParticle(sh, mat, fixed, 1u); }, py::arg("shape"), py::arg("mat"), py::arg("fixed") = false);
  // ---- contacts
  py::class_<CGeomNode>(m, "CGeomNode").def_property_readonly("pos", [](CGeomNode& n) { return np3(n.pos); });
  py::class_<L6Geom, shared_ptr<L6Geom>>(m, "L6Geom").def_readonly("uN", &L6Geom::uN).def_property_readonly("trsf", [](L6Geom& g) { return np33(g.trsf); }).def_readonly("node", &L6Geom::node);
  py::class_<FrictPhys, shared_ptr<FrictPhys>>(m, "FrictPhys").def_property_readonly("force", [](FrictPhys& p) { return np3(p.force); }).def_property_readonly("torque", [](FrictPhys& p) { return np3(p.torque); })
      .def_readonly("kn", &FrictPhys::kn).def_readonly("kt", &FrictPhys::kt).def_readonly("tanPhi", &FrictPhys::tanPhi);
  py::class_<Contact, shared_ptr<Contact>>(m, "Contact").def_readonly("id1", &Contact::id1).def_readonly("id2", &Contact::id2).def_readonly("stepCreated", &Contact::stepCreated)
      .def_readonly("geom", &Contact::geom).def_readonly("phys", &Contact::phys).def_property_readonly("real", &Contact::isReal)
      .def_property_readonly("cellDist", [](Contact& c) { return np3i(c.cellDist); });
  py::class_<ContactContainer>(m, "ContactContainer")
      .def("__len__", [](ContactContainer& c) { c.dem->pullContacts(); return c.v.size(); })
      .def("__getitem__", [](ContactContainer& c, std::pair<int, int> ij) { c.dem->pullContacts(); auto it = c.ix.find({std::min(ij.first, ij.second), std::max(ij.first, ij.second)}); if (it == c.ix.end()) throw py::index_error("No such contact ##" + std::to_string(ij.first) + "+" + std::to_string(ij.second)); return c.v[it->second]; })
      .def("__getitem__", [](ContactContainer& c, long i) { c.dem->pullContacts(); if (i < 0) i += (long)c.v.size(); if (i < 0 || i >= (long)c.v.size()) throw py::index_error(); return c.v[i]; })
      .def("existsReal", [](ContactContainer& c, int a, int b) { c.dem->pullContacts(); auto it = c.ix.find({std::min(a, b), std::max(a, b)}); return it != c.ix.end() && c.v[it->second]->isReal(); })
      .def("countReal", [](ContactContainer& c) { c.dem->pullContacts(); size_t n = 0; for (auto& x : c.v) n += x->isReal(); return n; })
      .def("__iter__", [](ContactContainer& c) { c.dem->pullContacts(); return py::make_iterator(c.v.begin(), c.v.end()); }, py::keep_alive<0, 1>());
  py::class_<ParticleContainer>(m, "ParticleContainer")
      .def("__len__", [](ParticleContainer& c) { return c.v.size(); })
      .def("__getitem__", [](ParticleContainer& c, long i) { if (i < 0) i += (long)c.v.size(); if (i < 0 || i >= (long)c.v.size() || !c.v[i]) throw py::index_error("No such particle: #" + std::to_string(i)); return c.v[i]; })
      .def("__iter__", [](ParticleContainer& c) { return py::make_iterator(c.v.begin(), c.v.end()); }, py::keep_alive<0, 1>())
      .def("add", [](ParticleContainer& c, py::object o) -> py::object {   // ParticleContainer::add + pyAppend, nodes=-1 (ParticleContainer.cpp:23-73,193-198)
        auto addOne = [&](shared_ptr<Particle> p) {
          if (p->id >= 0) throw std::runtime_error("Particle already has id " + std::to_string(p->id) + " set; appending such particle (for the second time) is not allowed.");
          const int id = c.findFreeId(); if ((size_t)id >= c.v.size()) c.v.resize(id + 1);      // ParticleContainer::insert -> insertAt (ParticleContainer.cpp:55-73)
          p->id = id; p->field = c.dem; c.v[id] = p;
          for (auto& n : p->shape->nodes) {
            n->dem->field = c.dem; const DemData& d = *n->dem;
            const bool moving = d.guessMoving();
            if (moving) c.dem->nodes.push_back(n);
          }
          if (c.dem->owner && !c.dem->topoDirty && dynamic_cast<Sphere*>(p->shape.get())) c.dem->pendingAdds.push_back(p);
          else c.dem->topoDirty = true;
          return p->id; };
        if (py::isinstance<Particle>(o)) return py::int_(addOne(o.cast<shared_ptr<Particle>>()));
        py::list ids; for (auto it : o) ids.append(addOne(it.cast<shared_ptr<Particle>>())); return ids; });
  m.attr("ParticleContainer").attr("remove") = py::cpp_function([](ParticleContainer& c, int id) {
        // ParticleContainer::remove (ParticleContainer.cpp:121-157): the particle and its contacts go, its id joins freeIds; removing the last element shrinks the container
        if (id < 0 || (size_t)id >= c.v.size() || !c.v[id]) throw py::index_error("No such particle: #" + std::to_string(id));
        shared_ptr<Particle> p = c.v[id];
        DemField& F = *c.dem;
        const bool sphere = dynamic_cast<Sphere*>(p->shape.get()) != nullptr;
        auto pa = std::find(F.pendingAdds.begin(), F.pendingAdds.end(), p);
        if (pa != F.pendingAdds.end()) F.pendingAdds.erase(pa);                       // never reached the device
        else if (F.owner && !F.topoDirty && sphere) F.pendingRemoves.push_back(id);
        else F.topoDirty = true;
        if (F.topoDirty && F.owner) { F.pull(); F.pullContacts(); }                  // the full path re-uploads from the host copies: make them current first
        for (auto& n : p->shape->nodes) { auto it = std::find(F.nodes.begin(), F.nodes.end(), n); if (it != F.nodes.end()) F.nodes.erase(it); }
        c.freeIds.push_back(id);
        c.v[id].reset(); p->id = -1; F.contactsStale = true;
        if ((size_t)(id + 1) == c.v.size()) { int k = id; while (k >= 0 && !c.v[k]) k--; c.v.resize(k + 1); }
      }, py::is_method(m.attr("ParticleContainer")), py::arg("id"));
  m.attr("ParticleContainer").attr("addClumped") = py::cpp_function([](ParticleContainer& c, py::object pars, shared_ptr<Node> centralNode) -> shared_ptr<Node> {
        // ParticleContainer::addClumped (ParticleContainer.cpp) for boundary particles with a user-given master node: the members keep their positions, the master's
        // vel/angVel (all DOFs blocked) move them rigidly from now on (examples/mill.py:38-42)
        if (!centralNode) {
          // ClumpData::makeClump for particles with mass (Clump.cpp:187-261): centroid, principal axes and inertia from the members (spheres only on this path)
          vector<shared_ptr<Particle>> pp; for (auto it : pars) pp.push_back(it.cast<shared_ptr<Particle>>());
          if (pp.empty()) throw std::runtime_error("ClumpData::makeClump: 0 nodes.");
          py::object add = py::cast(&c, py::return_value_policy::reference).attr("add");
          auto cd = make_shared<ClumpData>(); cd->flags |= CLUMP_CLUMP; cd->field = c.dem;
          auto cNode = make_shared<Node>(); cNode->dem = cd;
          double M = 0; V3 Sg{{0, 0, 0}}; double Ig[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
          for (auto& p : pp) {
            if (!dynamic_cast<Sphere*>(p->shape.get())) throw std::runtime_error("addClumped: without centralNode only Spheres can be clumped on this path");
            Node& n = *p->shape->nodes[0]; DemData& d = *n.dem;
            if (d.flags & CLUMP_CLUMPED) throw std::runtime_error("Node of particle #" + std::to_string(p->id) + ": already clumped.");
            M += d.mass;
            for (int k = 0; k < 3; k++) Sg[k] += d.mass * n.pos[k];
            // inertiaTensorTranslate(diag(inertia) [a sphere: rotation changes nothing], mass, -pos) (Volumetric.cpp:194-197)
            const double oo = n.pos[0] * n.pos[0] + n.pos[1] * n.pos[1] + n.pos[2] * n.pos[2];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ig[i][j] += (i == j ? d.inertia[i] : 0.) + d.mass * ((i == j ? oo : 0.) - n.pos[i] * n.pos[j]);
          }
          if (!(M > 0)) throw std::runtime_error("ClumpData::makeClump: no nodes with mass, therefore centralNode must be given, of which position will be used instead");
          if (pp.size() == 1) {   // Clump.cpp:204-217
            Node& n0 = *pp[0]->shape->nodes[0]; cNode->pos = n0.pos; cNode->ori = n0.ori; cd->mass = n0.dem->mass; cd->inertia = n0.dem->inertia;
          } else {
            // Volumetric::computePrincipalAxes (Volumetric.cpp:173-191): centroid, inertia about it, eigen-decomposition.  Jacobi rotations stand in for Eigen's
            // SelfAdjointEigenSolver (eigenvalues ascending as there); the eigenvector matrix is made a proper rotation before it becomes the orientation
            for (int k = 0; k < 3; k++) cNode->pos[k] = Sg[k] / M;
            const V3& P = cNode->pos; const double pp2 = P[0] * P[0] + P[1] * P[1] + P[2] * P[2];
            double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[i][j] = Ig[i][j] - M * ((i == j ? pp2 : 0.) - P[i] * P[j]);
            for (int sweep = 0; sweep < 60; sweep++) {
              const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
              if (off <= 1e-60 * (A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2])) break;
              for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
                if (A[p][q] == 0.) continue;
                const double th = (A[q][q] - A[p][p]) / (2 * A[p][q]), t = (th >= 0 ? 1. : -1.) / (std::abs(th) + std::sqrt(th * th + 1)), cs = 1 / std::sqrt(t * t + 1), sn = t * cs;
                for (int k = 0; k < 3; k++) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq; }
                for (int k = 0; k < 3; k++) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk; }
                for (int k = 0; k < 3; k++) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
              }
            }
            int ord[3] = {0, 1, 2}; std::sort(ord, ord + 3, [&](int a, int b) { return A[a][a] < A[b][b]; });
            double R[3][3]; for (int j = 0; j < 3; j++) { cd->inertia[j] = A[ord[j]][ord[j]]; for (int i = 0; i < 3; i++) R[i][j] = V[i][ord[j]]; }
            const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) + R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
            if (det < 0) for (int i = 0; i < 3; i++) R[i][2] = -R[i][2];
            // Eigen Quaternion(Matrix3), then normalize
            double qw, qx, qy, qz; const double tr = R[0][0] + R[1][1] + R[2][2];
            if (tr > 0) { double t = std::sqrt(tr + 1.0); qw = 0.5 * t; t = 0.5 / t; qx = (R[2][1] - R[1][2]) * t; qy = (R[0][2] - R[2][0]) * t; qz = (R[1][0] - R[0][1]) * t; }
            else {
              int i = 0; if (R[1][1] > R[0][0]) i = 1; if (R[2][2] > R[i][i]) i = 2; const int j = (i + 1) % 3, k = (j + 1) % 3;
              double t = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0); double qv[3]; qv[i] = 0.5 * t; t = 0.5 / t; qw = (R[k][j] - R[j][k]) * t; qv[j] = (R[j][i] + R[i][j]) * t; qv[k] = (R[k][i] + R[i][k]) * t;
              qx = qv[0]; qy = qv[1]; qz = qv[2];
            }
            const double qn = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw); cNode->ori = {{qx / qn, qy / qn, qz / qn, qw / qn}};
            cd->mass = M;
          }
          cd->equivRad = (std::sqrt(cd->inertia[0] / cd->mass) + std::sqrt(cd->inertia[1] / cd->mass) + std::sqrt(cd->inertia[2] / cd->mass)) / 3;
          // relPos = ori^-1 * (pos - clumpPos), relOri = ori^-1 * ori_member (Clump.cpp:252-253)
          const Q4& Q = cNode->ori; const double cw = Q[3], cx = -Q[0], cy = -Q[1], cz = -Q[2];
          for (auto& p : pp) {
            if (p->id < 0) add(p);
            Node& n = *p->shape->nodes[0]; DemData& d = *n.dem;
            const double v[3] = {n.pos[0] - cNode->pos[0], n.pos[1] - cNode->pos[1], n.pos[2] - cNode->pos[2]};
            double uv[3] = {cy * v[2] - cz * v[1], cz * v[0] - cx * v[2], cx * v[1] - cy * v[0]}; for (double& u : uv) u += u;      // Eigen _transformVector
            const double c2[3] = {cy * uv[2] - cz * uv[1], cz * uv[0] - cx * uv[2], cx * uv[1] - cy * uv[0]};
            cd->relPos.push_back(V3{{(v[0] + cw * uv[0]) + c2[0], (v[1] + cw * uv[1]) + c2[1], (v[2] + cw * uv[2]) + c2[2]}});
            const double bx = n.ori[0], by = n.ori[1], bz = n.ori[2], bw = n.ori[3];
            cd->relOri.push_back(Q4{{cw * bx + cx * bw + cy * bz - cz * by, cw * by + cy * bw + cz * bx - cx * bz, cw * bz + cz * bw + cx * by - cy * bx, cw * bw - cx * bx - cy * by - cz * bz}});
            cd->nodes.push_back(p->shape->nodes[0]);
            d.flags |= CLUMP_CLUMPED; d.master = cNode;
            // pyAppendClumped adds the particles without their nodes: only the clump node goes to DemField.nodes (ParticleContainer.cpp:213)
            auto itn = std::find(c.dem->nodes.begin(), c.dem->nodes.end(), p->shape->nodes[0]); if (itn != c.dem->nodes.end()) c.dem->nodes.erase(itn);
          }
          c.dem->sphereClumps.push_back(cNode); c.dem->nodes.push_back(cNode); c.dem->topoDirty = true;
          return cNode;
        }
        if (!centralNode->dem) { centralNode->dem = make_shared<DemData>(); centralNode->dem->flags |= DOF_ALL; }
        centralNode->dem->field = c.dem;
        ClumpGroup G; G.master = centralNode;
        py::object add = py::cast(&c, py::return_value_policy::reference).attr("add");
        for (auto it : pars) { auto p = it.cast<shared_ptr<Particle>>(); if (p->id < 0) add(p); G.members.push_back(p->id); }
        c.dem->clumps.push_back(G); c.dem->nodes.push_back(centralNode); c.dem->topoDirty = true;
        return centralNode; }, py::is_method(m.attr("ParticleContainer")), py::arg("pars"), py::arg("centralNode") = shared_ptr<Node>());
  kwclass<DemField>(m, "DemField")
      .def_property_readonly("par", [](DemField& f) -> ParticleContainer& { return f.par; }, py::return_value_policy::reference_internal)
      .def_property_readonly("con", [](DemField& f) -> ContactContainer& { return f.con; }, py::return_value_policy::reference_internal)
      .def_readonly("nodes", &DemField::nodes)
      .def("nodesAppend", [](DemField& f, shared_ptr<Node> n) {   // DemField::pyNodesAppend (Particle.cpp): a clumped node in DemField.nodes is skipped by Leapfrog (Leapfrog.cpp:153) — harmless
        if (!n || !n->dem) throw std::runtime_error("DemField.nodesAppend: Node.dem is None");
        if (std::find(f.nodes.begin(), f.nodes.end(), n) != f.nodes.end()) throw std::runtime_error("DemField.nodesAppend: the node is already in DemField.nodes");
        f.nodes.push_back(n); })
      .def_property("gravity", [](DemField& f) { return np3(f.gravity); }, [](DemField& f, py::object o) { f.gravity = seqN<3>(o, "gravity"); })
      RW(DemField, loneMask) RW(DemField, distFactor)
      .def_property_readonly_static("defaultMovableMask", [](py::object) { return 5u; }).def_property_readonly_static("defaultBoundaryMask", [](py::object) { return 3u; })
      .def_property_readonly_static("defaultLoneMask", [](py::object) { return 2u; });
  // ---- cell
  py::class_<Cell, shared_ptr<Cell>>(m, "Cell")
      .def_property("hSize", [](Cell& c) { return np33(c.hSize); }, [](Cell& c, py::object o) {
        M3 h = seqN<9>(py::module_::import("numpy").attr("asarray")(o, "float64").attr("reshape")(9), "hSize");
        const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) + h[2] * (h[3] * h[7] - h[4] * h[6]);
        if (det == 0) throw std::runtime_error("Cell is degenerate (zero volume).");   // Cell.cpp:99
        c.hSize = h; c.dirty = true; })
      .def("setBox", [](Cell& c, double x, double y, double z) { c.hSize = {{x, 0, 0, 0, y, 0, 0, 0, z}}; c.dirty = true; })
      .def_property("gradV", [](Cell& c) { return np33(c.gradV); }, [](Cell& c, py::object o) { c.gradV = seqN<9>(py::module_::import("numpy").attr("asarray")(o, "float64").attr("reshape")(9), "gradV"); c.dirty = true; })
      .def_property("nextGradV", [](Cell& c) { return np33(c.nextGradV); }, [](Cell& c, py::object o) { c.nextGradV = seqN<9>(py::module_::import("numpy").attr("asarray")(o, "float64").attr("reshape")(9), "nextGradV"); c.dirty = true; })
      .def_property("homoDeform", [](Cell& c) { return c.homoDeform; }, [](Cell& c, py::object o) {
        if (py::isinstance<py::str>(o)) { static const std::map<string, int> nm{{"none", 0}, {"pos", 1}, {"vel", 2}, {"vel2", 3}, {"gradV2", 4}}; auto it = nm.find(o.cast<string>()); if (it == nm.end()) throw std::invalid_argument("Cell.homoDeform: unknown name"); c.homoDeform = it->second; }
        else c.homoDeform = o.cast<int>();
        c.dirty = true; });
  // ---- engines and functors (attribute carriers; names and defaults from the reference headers cited in include/wdem.h)
  kwclass<Engine>(m, "Engine") RW(Engine, dead) RW(Engine, label) RW(Engine, field).def("__call__", [](Engine& e, Scene& s) { e.scene = &s; if (!e.field && !s.fields.empty()) e.field = s.fields[0]; e.run(); });
  kwclass<Leapfrog, Engine>(m, "Leapfrog") RW(Leapfrog, damping) RW(Leapfrog, reset) RW(Leapfrog, kinSplit) RW(Leapfrog, dontCollect);
  py::class_<BoundFunctor, shared_ptr<BoundFunctor>>(m, "BoundFunctor");
  kwclass<Bo1_Sphere_Aabb, BoundFunctor>(m, "Bo1_Sphere_Aabb"); kwclass<Bo1_Facet_Aabb, BoundFunctor>(m, "Bo1_Facet_Aabb"); kwclass<Bo1_Wall_Aabb, BoundFunctor>(m, "Bo1_Wall_Aabb") RW(Bo1_Wall_Aabb, halfThick);
  py::class_<InsertionSortCollider, Engine, shared_ptr<InsertionSortCollider>>(m, "InsertionSortCollider")
      .def(py::init([](vector<shared_ptr<BoundFunctor>> bo, py::kwargs kw) { auto o = make_shared<InsertionSortCollider>(); o->boundDispatcher = bo; py::object po = py::cast(o); for (auto it : kw) py::setattr(po, it.first, it.second); return o; }), py::arg("boundDispatcher") = vector<shared_ptr<BoundFunctor>>())
      RW(InsertionSortCollider, boundDispatcher) RW(InsertionSortCollider, verletDist).def_readonly("nFullRuns", &InsertionSortCollider::nFullRuns);
  py::class_<CGeomFunctor, shared_ptr<CGeomFunctor>>(m, "CGeomFunctor") RW(CGeomFunctor, noRatch) RW(CGeomFunctor, iniLensTouch) RW(CGeomFunctor, trsfRenorm) RW(CGeomFunctor, approxMask);
  kwclass<Cg2_Sphere_Sphere_L6Geom, CGeomFunctor>(m, "Cg2_Sphere_Sphere_L6Geom"); kwclass<Cg2_Wall_Sphere_L6Geom, CGeomFunctor>(m, "Cg2_Wall_Sphere_L6Geom"); kwclass<Cg2_Facet_Sphere_L6Geom, CGeomFunctor>(m, "Cg2_Facet_Sphere_L6Geom");
  py::class_<CPhysFunctor, shared_ptr<CPhysFunctor>>(m, "CPhysFunctor") RW(CPhysFunctor, label);
  kwclass<Cp2_FrictMat_FrictPhys, CPhysFunctor>(m, "Cp2_FrictMat_FrictPhys") RW(Cp2_FrictMat_FrictPhys, tanPhi);
  kwclass<Cp2_HertzMat_HertzPhys, CPhysFunctor>(m, "Cp2_HertzMat_HertzPhys") RW(Cp2_HertzMat_HertzPhys, poisson) RW(Cp2_HertzMat_HertzPhys, en);
  py::class_<LawFunctor, shared_ptr<LawFunctor>>(m, "LawFunctor") RW(LawFunctor, label);
  kwclass<Law2_L6Geom_FrictPhys_IdealElPl, LawFunctor>(m, "Law2_L6Geom_FrictPhys_IdealElPl") RW(Law2_L6Geom_FrictPhys_IdealElPl, noSlip) RW(Law2_L6Geom_FrictPhys_IdealElPl, noBreak) RW(Law2_L6Geom_FrictPhys_IdealElPl, noFrict)
      RW(Law2_L6Geom_FrictPhys_IdealElPl, iniEqlb) RW(Law2_L6Geom_FrictPhys_IdealElPl, relRollStiff) RW(Law2_L6Geom_FrictPhys_IdealElPl, relTwistStiff) RW(Law2_L6Geom_FrictPhys_IdealElPl, rollTanPhi);
  kwclass<Law2_L6Geom_HertzPhys_DMT, LawFunctor>(m, "Law2_L6Geom_HertzPhys_DMT") RW(Law2_L6Geom_HertzPhys_DMT, noAttraction);
  py::class_<ContactLoop, Engine, shared_ptr<ContactLoop>>(m, "ContactLoop")
      .def(py::init([](vector<shared_ptr<CGeomFunctor>> g, vector<shared_ptr<CPhysFunctor>> p, vector<shared_ptr<LawFunctor>> l, py::kwargs kw) {
        auto o = make_shared<ContactLoop>(); o->geoDisp = g; o->phyDisp = p; o->lawDisp = l; py::object po = py::cast(o); for (auto it : kw) py::setattr(po, it.first, it.second); return o; }),
        py::arg("geoDisp") = vector<shared_ptr<CGeomFunctor>>(), py::arg("phyDisp") = vector<shared_ptr<CPhysFunctor>>(), py::arg("lawDisp") = vector<shared_ptr<LawFunctor>>())
      RW(ContactLoop, geoDisp) RW(ContactLoop, phyDisp) RW(ContactLoop, lawDisp) RW(ContactLoop, applyForces) RW(ContactLoop, evalStress).def_property_readonly("stress", [](ContactLoop& c) { return np33(c.stress); }) RW(ContactLoop, dist00) RW(ContactLoop, updatePhys) RW(ContactLoop, reorderEvery) RW(ContactLoop, hook);
  kwclass<DynDt, Engine>(m, "DynDt") RW(DynDt, stepPeriod) RW(DynDt, maxRelInc) RW(DynDt, dryRun) RW(DynDt, stepLast) RW(DynDt, stepPrev);
  py::class_<GpuDemLoop, Engine, shared_ptr<GpuDemLoop>>(m, "GpuDemLoop")
      .def(py::init([](shared_ptr<Leapfrog> lf, shared_ptr<InsertionSortCollider> co, shared_ptr<ContactLoop> cl, shared_ptr<DynDt> dd, int device, py::kwargs kw) {
        auto o = make_shared<GpuDemLoop>(); o->leapfrog = lf; o->collider = co; o->contactLoop = cl; o->dynDt = dd; o->device = device; py::object po = py::cast(o); for (auto it : kw) py::setattr(po, it.first, it.second); return o; }),
        py::arg("leapfrog"), py::arg("collider"), py::arg("contactLoop"), py::arg("dynDt") = py::none(), py::arg("device") = 0)
      RW(GpuDemLoop, leapfrog) RW(GpuDemLoop, collider) RW(GpuDemLoop, contactLoop) RW(GpuDemLoop, dynDt) RW(GpuDemLoop, device)
      .def("sync", [](GpuDemLoop& g) { if (g.field) { g.field->pull(); g.field->pullContacts(); } }, "Download particle state and contacts into the host objects now (otherwise done lazily on first access).")
      .def("release", &GpuDemLoop::release)
      .def_property_readonly("stats", [](GpuDemLoop& g) { py::dict d; d["n_spheres"] = g.last.n_spheres; d["n_contacts"] = g.last.n_contacts; d["n_real"] = g.last.n_real; d["n_rebuilds"] = g.last.n_rebuilds; d["n_launches"] = g.last.n_launches; d["device_ms"] = g.last.device_ms; return d; });
  // ---- scene
  py::class_<Scene, shared_ptr<Scene>>(m, "Scene")
      .def(py::init([](vector<shared_ptr<DemField>> fields, py::kwargs kw) { auto s = make_shared<Scene>(); s->fields = fields; py::object po = py::cast(s); for (auto it : kw) py::setattr(po, it.first, it.second); return s; }), py::arg("fields") = vector<shared_ptr<DemField>>())
      RW(Scene, dt) RW(Scene, nextDt) RW(Scene, dtSafety) RW(Scene, time) RW(Scene, step) RW(Scene, periodic) RW(Scene, trackEnergy).def_property_readonly("energy", [](Scene& S) { py::dict d; for (auto& kv : S.energy) d[py::str(kv.first)] = kv.second; return d; }) RW(Scene, deterministic) RW(Scene, engines) RW(Scene, fields)
      .def_readonly("cell", &Scene::cell).def_property_readonly("dem", &Scene::dem)
      .def("one", &Scene::one).def("run", [](Scene& s, long n, bool) { s.run(n); }, py::arg("steps") = 1, py::arg("wait") = true).def("wait", [](Scene&) {});
}
