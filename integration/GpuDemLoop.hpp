// integration/GpuDemLoop.hpp — the in-Woo build of the drop-in engine (SURVEY.md §8b "mode 1").
//
// Drop this file and GpuDemLoop.cpp into woo/src/dem/, add the .cpp to WOO_CXX_SOURCES (CMakeLists.txt:311-428) and link libwdem.so
// (target_link_libraries(woo PRIVATE wdem); the library is built by `python -m woo_b200.build`, its C-ABI is include/wdem.h).
// The class registers like every other engine of the reference (WOO_DECL__CLASS_BASE_DOC_ATTRS + WOO_REGISTER_OBJECT here,
// WOO_PLUGIN(dem,(GpuDemLoop)) + WOO_IMPL__… in the .cpp: Object.hpp:316-413,47-51) and shows up as woo.dem.GpuDemLoop.
//
//   S.engines=[GpuDemLoop(leapfrog=Leapfrog(damping=.2,reset=True),
//                         collider=InsertionSortCollider([Bo1_Sphere_Aabb(),Bo1_Facet_Aabb(),Bo1_Wall_Aabb()]),
//                         contactLoop=ContactLoop([Cg2_Sphere_Sphere_L6Geom(),Cg2_Facet_Sphere_L6Geom(),Cg2_Wall_Sphere_L6Geom()],[Cp2_FrictMat_FrictPhys()],[Law2_L6Geom_FrictPhys_IdealElPl()]),
//                         dynDt=DynDt(stepPeriod=100))]
//
// It occupies ONE slot of Scene.engines in place of [Leapfrog, InsertionSortCollider, ContactLoop(, DynDt)] (py/utils.py:122-129); run() is one full
// step of that group in the reference's order.  The wrapped engine objects are only attribute carriers here: their own run() is never called.
// This file cannot be compiled in the repository's container (Woo's headers need Eigen and Boost); the standalone twin with the same member
// functions, woo_b200/csrc/woo_plugin.cpp, is what the tests run.  Keep the two in step.
#pragma once
#include"Particle.hpp"
#include"Leapfrog.hpp"
#include"InsertionSortCollider.hpp"
#include"ContactLoop.hpp"
#include"DynDt.hpp"
#include<wdem.h>

struct GpuDemLoop: public Engine{
	bool acceptsField(Field* f) override { return dynamic_cast<DemField*>(f); }      // Engine.hpp:42
	void run() override;                                                           // Scene::doOneStep calls this once per step (Scene.cpp:420-426)
	Real critDt() override;                                                        // polled while Scene.dt is NaN (Scene.cpp:380-397): DemFuncs::pWaveDt
	void selfTest() override;                                                      // step 0 (Scene.cpp:344-363): unsupported settings fail here, loudly
	void notifyDead() override { if(dead) release(); }
	~GpuDemLoop(){ release(); }
	// host <-> device
	void release();
	wdem_config makeConfig();            // attributes of the wrapped engines / functors -> wdem_config (re-read every step, as the reference engines do)
	void upload();                       // DemField -> device: materials, spheres, clumps of spheres, walls, facets, kinematic clumps of facets, cell, contact history
	void prepare();                      // before a step: (re-)upload if particles changed, else push attribute / state edits
	void pull();                         // device -> Node.pos/ori, DemData.vel/angVel/force/torque (lazily: only when `sync` or somebody asks)
	void pullContacts();                 // device -> DemField.contacts as real Contact/L6Geom/FrictPhys|HertzPhys objects (S.save(), S.dem.con[i,j], VtkExport keep working)
	void uploadImposes(const vector<shared_ptr<Impose>>& sphImpose, const vector<int32_t>& sphImposeId);   // DemData.impose of spheres, walls and facet-clump masters -> wdem_upload_imposes
	void pushContacts();                 // real contacts of DemField.contacts -> wdem_upload_contacts (resume from a loaded Scene)
	WOO_DECL_LOGGER;
	// not attributes: device handle and the host-side bookkeeping of what lives there
	wdem_handle h=nullptr; wdem_config cfg; bool haveCfg=false;
	vector<shared_ptr<Particle>> sphereOf, wallOf, facetOf;      // particles in upload order
	vector<shared_ptr<Impose>> imposeOf;                         // Impose objects in upload order (their sumF / dist / F / T are read back in pull())
	vector<shared_ptr<Node>> clumpOf;                            // master nodes of the clumps of spheres, in upload order
	size_t nParticlesUploaded=0; long stepUploaded=-1; bool trackingOn=false;
	#define woo_dem_GpuDemLoop__CLASS_BASE_DOC_ATTRS \
		GpuDemLoop,Engine,"Leapfrog + InsertionSortCollider + ContactLoop (+ DynDt) for spheres, walls and facets as one engine running on a CUDA device (B200, sm_100a; no CPU fallback). Attributes of the wrapped engines keep their meaning and are re-read every step; settings the device path does not implement raise in :obj:`selfTest`.", \
		((shared_ptr<Leapfrog>,leapfrog,,,":obj:`Leapfrog` whose attributes (:obj:`~Leapfrog.damping`, :obj:`~Leapfrog.reset`) are used.")) \
		((shared_ptr<InsertionSortCollider>,collider,,,":obj:`InsertionSortCollider` whose :obj:`~AabbCollider.verletDist` and :obj:`Bo1_Wall_Aabb.halfThick` are used; :obj:`~InsertionSortCollider.nFullRuns` is updated.")) \
		((shared_ptr<ContactLoop>,contactLoop,,,":obj:`ContactLoop` with the Cg2 functors for Sphere/Wall/Facet+Sphere, exactly one Cp2 (:obj:`Cp2_FrictMat_FrictPhys` or :obj:`Cp2_HertzMat_HertzPhys`) and the matching Law2 functor.")) \
		((shared_ptr<DynDt>,dynDt,,,"Optional :obj:`DynDt`; its :obj:`~PeriodicEngine.stepPeriod` and :obj:`~DynDt.maxRelInc` are used.")) \
		((int,device,0,,"CUDA device ordinal.")) \
		((bool,sync,false,,"Read the particle state back after every step (for scripts that look at particles every step); otherwise it is read back when the engine is asked to (:obj:`pull`), before saving, and when particles are added or removed.")) \
		((long,nRebuilds,0,AttrTrait<Attr::readonly>(),"Collider full runs in the last call of :obj:`run`.")) \
		((Real,deviceMs,0,AttrTrait<Attr::readonly>(),"Device time of the last call of :obj:`run` [ms]."))
	WOO_DECL__CLASS_BASE_DOC_ATTRS(woo_dem_GpuDemLoop__CLASS_BASE_DOC_ATTRS);
};
WOO_REGISTER_OBJECT(GpuDemLoop);
