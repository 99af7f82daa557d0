// integration/GpuDemLoop.cpp — see GpuDemLoop.hpp.  Written against the reference's own types (file:line refer to the woodem/woo tree);
// same member functions, same error texts and the same order of C-ABI calls as the standalone twin woo_b200/csrc/woo_plugin.cpp, which is
// the one the tests of this repository execute.  Not compiled here: Woo's headers need Eigen and Boost.
#include"GpuDemLoop.hpp"
#include"Sphere.hpp"
#include"Wall.hpp"
#include"Facet.hpp"
#include"Clump.hpp"
#include"Impose.hpp"
#include"FrictMat.hpp"
#include"Hertz.hpp"
#include"IdealElPl.hpp"
#include"L6Geom.hpp"
#include"Funcs.hpp"

WOO_PLUGIN(dem,(GpuDemLoop));
WOO_IMPL__CLASS_BASE_DOC_ATTRS(woo_dem_GpuDemLoop__CLASS_BASE_DOC_ATTRS);
WOO_IMPL_LOGGER(GpuDemLoop);

// every C-ABI call returns 0 or sets wdem_last_error; exceptions propagate to S.one() / are stored by the background loop (Scene.cpp:110-117)
static inline void chk(wdem_handle h, int rc){ if(rc!=WDEM_OK) throw std::runtime_error(string("GpuDemLoop: ")+wdem_last_error(h)); }

void GpuDemLoop::release(){ if(h){ wdem_destroy(h); h=nullptr; } haveCfg=false; }

void GpuDemLoop::selfTest(){ makeConfig(); }

wdem_config GpuDemLoop::makeConfig(){
	if(!leapfrog || !collider || !contactLoop) throw std::runtime_error("GpuDemLoop needs leapfrog, collider and contactLoop");
	DemField& dem=field->cast<DemField>();
	wdem_config c; wdem_default_config(&c);
	c.device=device; c.dt=scene->dt; c.dtSafety=scene->dtSafety; c.periodic=scene->isPeriodic?1:0;
	for(int k:{0,1,2}) c.gravity[k]=dem.gravity[k];
	c.loneMask=dem.loneMask; c.distFactor=dem.distFactor;
	c.damping=leapfrog->damping; c.reset=leapfrog->reset?1:0;                              // Leapfrog.hpp:50-53
	if(leapfrog->kinSplit) throw std::runtime_error("GpuDemLoop: Leapfrog.kinSplit is not supported");
	c.verletDist=collider->verletDist;                                                  // Collision.hpp:94
	for(const auto& b: collider->boundDispatcher->functors){ if(auto w=dynamic_pointer_cast<Bo1_Wall_Aabb>(b)) c.wallHalfThick=w->halfThick; }
	const auto& gg=contactLoop->geoDisp->functors; const auto& pp=contactLoop->phyDisp->functors; const auto& ll=contactLoop->lawDisp->functors;
	if(gg.empty() || pp.size()!=1 || ll.size()!=1) throw std::runtime_error("GpuDemLoop: ContactLoop needs Cg2 functors, exactly one Cp2 and one Law2 functor");
	shared_ptr<Cg2_Any_Any_L6Geom__Base> g0;
	for(const auto& g: gg){
		auto gb=dynamic_pointer_cast<Cg2_Any_Any_L6Geom__Base>(g);
		if(!gb) continue;          // functors for shapes the device path does not know are harmless as long as no such particle exists (checked in upload)
		if(gb->approxMask!=0) throw std::runtime_error("GpuDemLoop: Cg2_*.approxMask!=0 is not supported");
		if(!g0) g0=gb;
		else if(gb->noRatch!=g0->noRatch || gb->iniLensTouch!=g0->iniLensTouch || gb->trsfRenorm!=g0->trsfRenorm) throw std::runtime_error("GpuDemLoop: all Cg2 functors must share noRatch/iniLensTouch/trsfRenorm");
	}
	if(!g0) throw std::runtime_error("GpuDemLoop: no Cg2_*_L6Geom functor");
	c.noRatch=g0->noRatch; c.iniLensTouch=g0->iniLensTouch; c.trsfRenorm=g0->trsfRenorm;   // L6Geom.hpp:35-45
	c.applyForces=contactLoop->applyForces;
	if(contactLoop->updatePhys!=ContactLoop::UPDATE_PHYS_NEVER) throw std::runtime_error("GpuDemLoop: ContactLoop.updatePhys is not supported");
	if(contactLoop->hook) throw std::runtime_error("GpuDemLoop: ContactLoop.hook is not supported");
	c.evalStress=(contactLoop->evalStress && scene->isPeriodic)?1:0;          // with Scene.trackEnergy: the "gradV" energy (ContactLoop.cpp:243-250)
	auto lin=dynamic_pointer_cast<Law2_L6Geom_FrictPhys_IdealElPl>(ll[0]); auto dmt=dynamic_pointer_cast<Law2_L6Geom_HertzPhys_DMT>(ll[0]);
	auto cpF=dynamic_pointer_cast<Cp2_FrictMat_FrictPhys>(pp[0]); auto cpH=dynamic_pointer_cast<Cp2_HertzMat_HertzPhys>(pp[0]);
	if(lin && cpF){
		if(cpF->tanPhi) throw std::runtime_error("GpuDemLoop: Cp2_FrictMat_FrictPhys.tanPhi (MatchMaker) is not supported");
		if(lin->iniEqlb) throw std::runtime_error("GpuDemLoop: Law2_L6Geom_FrictPhys_IdealElPl.iniEqlb is not supported");
		c.law=WDEM_LAW_IDEALELPL; c.noSlip=lin->noSlip; c.noBreak=lin->noBreak; c.noFrict=lin->noFrict;          // IdealElPl.hpp:17-24
		c.relRollStiff=lin->relRollStiff; c.relTwistStiff=lin->relTwistStiff; c.rollTanPhi=lin->rollTanPhi;
	} else if(dmt && cpH){
		c.law=WDEM_LAW_HERTZ_DMT; c.noAttraction=dmt->noAttraction; c.poisson=cpH->poisson; c.en=cpH->en;         // Hertz.hpp:58-59,96
	} else throw std::runtime_error("GpuDemLoop: supported functor pairs are Cp2_FrictMat_FrictPhys+Law2_L6Geom_FrictPhys_IdealElPl and Cp2_HertzMat_HertzPhys+Law2_L6Geom_HertzPhys_DMT");
	if(dynDt && !dynDt->dead){ if(dynDt->dryRun) throw std::runtime_error("GpuDemLoop: DynDt.dryRun is not supported"); c.dynDtPeriod=dynDt->stepPeriod; c.maxRelInc=dynDt->maxRelInc; }
	else c.dynDtPeriod=0;
	return c;
}

void GpuDemLoop::upload(){
	DemField& dem=field->cast<DemField>();
	if(h){ pull(); pullContacts(); }            // a re-upload starts from the latest device state and keeps the contact history
	wdem_config c=makeConfig();
	if(h && c.law!=cfg.law){ release(); dem.contacts->clear(); }
	if(!h){ int rc=wdem_create(&c,&h); if(rc!=WDEM_OK){ h=nullptr; throw std::runtime_error(string("GpuDemLoop: ")+wdem_last_error(nullptr)); } }
	else chk(h,wdem_set_config(h,&c));
	cfg=c; haveCfg=true;
	// materials, de-duplicated by object identity
	vector<Material*> mats; std::map<Material*,int> matIx;
	auto matOf=[&](const shared_ptr<Particle>& p)->int{
		Material* m=p->material.get(); if(!m) throw std::runtime_error("particle #"+to_string(p->id)+" has no material");
		auto it=matIx.find(m); if(it!=matIx.end()) return it->second;
		int k=(int)mats.size(); mats.push_back(m); matIx[m]=k; return k;
	};
	vector<int32_t> sid,smat,wid,waxis,wsense,wmat,fid,fmat; vector<uint32_t> sflags,smask,wmask,fmask;
	vector<double> spos,sori,svel,sang,smass,sin,srad,sfrc,strq,wpos,wvel,fv,fvv,ffk,fht,fnl;
	vector<shared_ptr<Impose>> sphImpose; vector<int32_t> sphImposeId;
	sphereOf.clear(); wallOf.clear(); facetOf.clear();
	for(const auto& p: *dem.particles){
		if(!p) continue;
		if(!p->shape) throw std::runtime_error("particle #"+to_string(p->id)+" has no shape");
		const int m=matOf(p);
		if(auto* s=dynamic_cast<Sphere*>(p->shape.get())){
			const Node& n=*s->nodes[0]; const DemData& d=n.getData<DemData>();
			if(d.isClump()) throw std::runtime_error("GpuDemLoop: particle #"+to_string(p->id)+": a clump's master node cannot be the node of a particle");      // members (isClumped) are fine: wdem_upload_clumps below
			if(d.impose){ sphImpose.push_back(d.impose); sphImposeId.push_back(p->id); }
			if(!(d.inertia[0]==d.inertia[1] && d.inertia[1]==d.inertia[2])) throw std::runtime_error("GpuDemLoop: aspherical inertia is not supported (particle #"+to_string(p->id)+")");
			sid.push_back(p->id); smat.push_back(m); sflags.push_back(d.flags); smask.push_back(p->mask);
			for(int k:{0,1,2}){ spos.push_back(n.pos[k]); svel.push_back(d.vel[k]); sang.push_back(d.angVel[k]); sfrc.push_back(d.force[k]); strq.push_back(d.torque[k]); }
			sori.push_back(n.ori.x()); sori.push_back(n.ori.y()); sori.push_back(n.ori.z()); sori.push_back(n.ori.w());
			smass.push_back(d.mass); sin.push_back(d.inertia[0]); srad.push_back(s->radius);
			sphereOf.push_back(p);
		} else if(auto* w=dynamic_cast<Wall*>(p->shape.get())){
			const Node& n=*w->nodes[0]; const DemData& d=n.getData<DemData>();
			if(d.angVel!=Vector3r::Zero()) throw std::runtime_error("GpuDemLoop: rotating walls are not supported");
			wid.push_back(p->id); waxis.push_back(w->axis); wsense.push_back(w->sense); wmat.push_back(m); wmask.push_back(p->mask);
			for(int k:{0,1,2}){ wpos.push_back(n.pos[k]); wvel.push_back(d.vel[k]); }
			wallOf.push_back(p);
		} else if(auto* f=dynamic_cast<Facet*>(p->shape.get())){
			fid.push_back(p->id); fmat.push_back(m); fmask.push_back(p->mask); fht.push_back(f->halfThick);
			for(int v:{0,1,2}){ const Node& n=*f->nodes[v]; const DemData& d=n.getData<DemData>(); for(int k:{0,1,2}){ fv.push_back(n.pos[k]); fvv.push_back(d.vel[k]); } }
			for(int k:{0,1,2}){ ffk.push_back(f->fakeVel[k]); fnl.push_back(f->n21lim[k]); }
			facetOf.push_back(p);
		} else throw std::runtime_error("GpuDemLoop: particle #"+to_string(p->id)+": only Sphere, Wall and Facet shapes are supported");
	}
	if(mats.empty()) throw std::runtime_error("GpuDemLoop: no particles");
	vector<int32_t> mk; vector<double> md,my,mkt,mtp,mse,mal;
	for(Material* m: mats){
		auto* fm=dynamic_cast<FrictMat*>(m); if(!fm) throw std::runtime_error("GpuDemLoop: materials must be FrictMat or HertzMat");
		auto* hm=dynamic_cast<HertzMat*>(m);
		mk.push_back(hm?1:0); md.push_back(m->density); my.push_back(fm->young); mkt.push_back(fm->ktDivKn); mtp.push_back(fm->tanPhi); mse.push_back(hm?hm->surfEnergy:0.); mal.push_back(hm?hm->alpha:0.);
	}
	chk(h,wdem_upload_materials(h,(int)mats.size(),mk.data(),md.data(),my.data(),mkt.data(),mtp.data(),mse.data(),mal.data()));
	chk(h,wdem_upload_spheres(h,(int64_t)sid.size(),sid.data(),spos.data(),sori.data(),svel.data(),sang.data(),smass.data(),sin.data(),srad.data(),sflags.data(),smask.data(),smat.data()));
	chk(h,wdem_upload_forces(h,sfrc.data(),strq.data()));
	chk(h,wdem_upload_walls(h,(int)wid.size(),wid.data(),wpos.data(),waxis.data(),wsense.data(),wvel.data(),wmask.data(),wmat.data()));
	chk(h,wdem_upload_facets(h,(int)fid.size(),fid.data(),fv.data(),fvv.data(),ffk.data(),fht.data(),fnl.data(),fmask.data(),fmat.data()));
	// clumps: nodes in DemField.nodes that carry ClumpData (Clump.cpp:187-261).  Members that are spheres: a clump with mass, integrated on the device
	// (wdem_upload_clumps); members that are facets: a kinematic group whose master has all DOFs blocked (examples/mill.py:38-42)
	clumpOf.clear();
	{
		std::map<const Node*,int> facetOfNode; for(size_t i=0;i<facetOf.size();i++) for(const auto& n: facetOf[i]->shape->nodes) facetOfNode[n.get()]=(int)i;
		std::map<const Node*,int> idOfSphereNode; for(const auto& p: sphereOf) idOfSphereNode[p->shape->nodes[0].get()]=p->id;
		vector<int32_t> of(facetOf.size(),-1); vector<double> gp,go,gv,gw; int ng=0;
		vector<int32_t> cStart{0},cMem; vector<uint32_t> cFl; vector<double> cp,co,cv,cw,cl,cm,cI,cf,ct,crp,cro;
		for(const auto& n: dem.nodes){
			if(!n->hasData<DemData>() || !n->getData<DemData>().isClump()) continue;
			const ClumpData& cd=n->getData<DemData>().cast<ClumpData>();
			if(!cd.nodes.empty() && idOfSphereNode.count(cd.nodes[0].get())){
				if(cd.impose) throw std::runtime_error("GpuDemLoop: DemData.impose on a clump is not supported");
				for(size_t i=0;i<cd.nodes.size();i++){
					auto it=idOfSphereNode.find(cd.nodes[i].get()); if(it==idOfSphereNode.end()) throw std::runtime_error("GpuDemLoop: a clump must consist of Spheres only, or of Facets only");
					cMem.push_back(it->second);
					for(int k:{0,1,2}) crp.push_back(cd.relPos[i][k]);
					cro.push_back(cd.relOri[i].x()); cro.push_back(cd.relOri[i].y()); cro.push_back(cd.relOri[i].z()); cro.push_back(cd.relOri[i].w());
				}
				cStart.push_back((int32_t)cMem.size()); cFl.push_back(cd.flags); cm.push_back(cd.mass);
				for(int k:{0,1,2}){ cp.push_back(n->pos[k]); cv.push_back(cd.vel[k]); cw.push_back(cd.angVel[k]); cl.push_back(cd.angMom[k]); cI.push_back(cd.inertia[k]); cf.push_back(cd.force[k]); ct.push_back(cd.torque[k]); }
				co.push_back(n->ori.x()); co.push_back(n->ori.y()); co.push_back(n->ori.z()); co.push_back(n->ori.w());
				clumpOf.push_back(n);
				continue;
			}
			if(!cd.isBlockedAll()) throw std::runtime_error("GpuDemLoop: the master node of a clump of boundary particles must have all DOFs blocked (its motion is prescribed)");
			for(const auto& mn: cd.nodes){ auto it=facetOfNode.find(mn.get()); if(it==facetOfNode.end()) throw std::runtime_error("GpuDemLoop: only Facets can be members of a kinematic clump"); of[it->second]=ng; }
			for(int k:{0,1,2}){ gp.push_back(n->pos[k]); gv.push_back(cd.vel[k]); gw.push_back(cd.angVel[k]); }
			go.push_back(n->ori.x()); go.push_back(n->ori.y()); go.push_back(n->ori.z()); go.push_back(n->ori.w());
			ng++;
		}
		if(ng>0) chk(h,wdem_upload_facet_groups(h,ng,of.data(),gp.data(),go.data(),gv.data(),gw.data()));
		if(!clumpOf.empty()){
			chk(h,wdem_upload_clumps(h,(int32_t)clumpOf.size(),cStart.data(),cMem.data(),cp.data(),co.data(),cv.data(),cw.data(),cl.data(),cm.data(),cI.data(),cFl.data(),crp.data(),cro.data()));
			chk(h,wdem_upload_clump_forces(h,cf.data(),ct.data()));
		}
	}
	uploadImposes(sphImpose,sphImposeId);
	if(scene->isPeriodic){
		const Cell& C=*scene->cell; double H[9],G[9],N[9];
		for(int i:{0,1,2}) for(int j:{0,1,2}){ H[3*i+j]=C.hSize(i,j); G[3*i+j]=C.gradV(i,j); N[3*i+j]=C.nextGradV(i,j); }
		chk(h,wdem_set_cell(h,H,G,N,C.homoDeform));
	}
	chk(h,wdem_set_scene(h,(int64_t)scene->step,scene->time));
	if(dynDt && !dynDt->dead) chk(h,wdem_set_periodic_state(h,dynDt->stepLast,dynDt->stepPrev));   // a re-upload in mid-run does not restart DynDt's period
	pushContacts();
	nParticlesUploaded=dem.particles->size(); stepUploaded=scene->step;
}

// DemData.impose -> wdem_upload_imposes: one wdem_impose per distinct Impose object, one set per distinct DemData.impose (a CombinedImpose is a set of several)
void GpuDemLoop::uploadImposes(const vector<shared_ptr<Impose>>& sphImpose, const vector<int32_t>& sphImposeId){
	DemField& dem=field->cast<DemField>();
	imposeOf.clear();
	vector<wdem_impose> imps; vector<vector<double>> keep; std::map<Impose*,int> ixOf; std::map<Impose*,int> setOf; vector<int32_t> setStart{0},setOps;
	auto q4=[](double* o, const Quaternionr& q){ o[0]=q.x(); o[1]=q.y(); o[2]=q.z(); o[3]=q.w(); };
	std::function<int(const shared_ptr<Impose>&)> one=[&](const shared_ptr<Impose>& I)->int{
		auto it=ixOf.find(I.get()); if(it!=ixOf.end()) return it->second;
		wdem_impose w; memset(&w,0,sizeof(w)); w.ori[3]=1;
		if(auto* h=dynamic_cast<HarmonicOscillation*>(I.get())){ w.kind=1; w.flags=h->perpFree?1:0; w.v[0]=h->freq; w.v[1]=h->amp; for(int k:{0,1,2}) w.v[2+k]=h->dir[k]; w.v[5]=h->t0; }
		else if(auto* a=dynamic_cast<AlignedHarmonicOscillations*>(I.get())){ w.kind=2; for(int k:{0,1,2}){ w.v[k]=a->freqs[k]; w.v[3+k]=a->amps[k]; } }
		else if(auto* c=dynamic_cast<ConstantForce*>(I.get())){ w.kind=3; for(int k:{0,1,2}) w.v[k]=c->F[k]; }
		else if(auto* r=dynamic_cast<RadialForce*>(I.get())){ if(!r->nodeA||!r->nodeB) throw std::runtime_error("RadialForce: nodeA and nodeB must not be None."); w.kind=4; for(int k:{0,1,2}){ w.v[k]=r->nodeA->pos[k]; w.v[3+k]=r->nodeB->pos[k]; } w.v[6]=r->F; }
		else if(auto* l=dynamic_cast<Local6Dofs*>(I.get())){ w.kind=5; for(int k=0;k<6;k++){ w.v[k]=l->values[k]; w.v[6+k]=l->whats[k]; } q4(w.ori,l->ori); }
		else if(auto* v=dynamic_cast<VelocityAndReadForce*>(I.get())){ if(!v->energyName.empty()) throw std::runtime_error("GpuDemLoop: VelocityAndReadForce.energyName is not supported"); w.kind=6; w.flags=(v->latBlock?1:0)|(v->invF?2:0); for(int k:{0,1,2}) w.v[k]=v->dir[k]; w.v[3]=v->vel; }
		else if(auto* f=dynamic_cast<ReadForce*>(I.get())){ w.kind=7; if(f->node){ w.flags=1; for(int k:{0,1,2}) w.v[k]=f->node->pos[k]; q4(w.ori,f->node->ori); } }
		else if(auto* vv=dynamic_cast<VariableVelocity3d*>(I.get())){
			for(const auto& hk: vv->hooks) if(!hk.empty()) throw std::runtime_error("GpuDemLoop: VariableVelocity3d.hooks are not supported");
			w.kind=8; w.flags=(vv->diff?1:0)|(vv->wrap?2:0); w.v[0]=vv->t0; q4(w.ori,vv->ori); w.nTimes=(int)vv->times.size();
			keep.push_back(vv->times); w.times=keep.back().data();
			keep.emplace_back(); for(const auto& x: vv->vels) for(int k:{0,1,2}) keep.back().push_back(x[k]); w.vals=keep.back().data();
			if(!vv->angVels.empty()){ keep.emplace_back(); for(const auto& x: vv->angVels) for(int k:{0,1,2}) keep.back().push_back(x[k]); w.vals2=keep.back().data(); }
		}
		else if(auto* va=dynamic_cast<VariableAlignedRotation*>(I.get())){
			w.kind=9; w.flags=va->wrap?1:0; w.v[0]=va->axis; w.nTimes=(int)va->timeAngVel.size();
			keep.emplace_back(); for(const auto& x: va->timeAngVel) keep.back().push_back(x[0]); w.times=keep.back().data();
			keep.emplace_back(); for(const auto& x: va->timeAngVel) keep.back().push_back(x[1]); w.vals=keep.back().data();
		}
		else if(auto* im=dynamic_cast<InterpolatedMotion*>(I.get())){      // on master nodes of facet clumps only (libwdem rejects it elsewhere)
			w.kind=12; w.v[0]=im->t0; w.nTimes=(int)im->times.size();
			keep.push_back(im->times); w.times=keep.back().data();
			keep.emplace_back(); for(const auto& x: im->poss) for(int k:{0,1,2}) keep.back().push_back(x[k]); w.vals=keep.back().data();
			keep.emplace_back(); for(const auto& q: im->oris){ keep.back().push_back(q.x()); keep.back().push_back(q.y()); keep.back().push_back(q.z()); keep.back().push_back(q.w()); } w.vals2=keep.back().data();
		}
		else if(auto* so=dynamic_cast<StableCircularOrbit*>(I.get())){      // derives from CircularOrbit (Impose.hpp:40): test it first
			if(!so->node) throw std::runtime_error("StableCircularOrbit: node must not be None.");
			w.kind=11; w.flags=so->rotate?1:0; for(int k:{0,1,2}) w.v[k]=so->node->pos[k]; w.v[3]=so->omega; w.v[4]=so->radius; w.v[5]=so->angle;
			w.ori[0]=so->node->ori.x(); w.ori[1]=so->node->ori.y(); w.ori[2]=so->node->ori.z(); w.ori[3]=so->node->ori.w();
		}
		else if(auto* co=dynamic_cast<CircularOrbit*>(I.get())){
			if(!co->node) throw std::runtime_error("CircularOrbit: node must not be None.");
			w.kind=10; w.flags=co->rotate?1:0; for(int k:{0,1,2}) w.v[k]=co->node->pos[k]; w.v[3]=co->omega; w.v[5]=co->angle;
			w.ori[0]=co->node->ori.x(); w.ori[1]=co->node->ori.y(); w.ori[2]=co->node->ori.z(); w.ori[3]=co->node->ori.w();
		}
		else throw std::runtime_error("GpuDemLoop: "+I->pyStr()+" is not supported on the device path");
		imps.push_back(w); imposeOf.push_back(I); return ixOf[I.get()]=(int)imps.size()-1;
	};
	auto setFor=[&](const shared_ptr<Impose>& I)->int{
		auto it=setOf.find(I.get()); if(it!=setOf.end()) return it->second;
		if(auto* c=dynamic_cast<CombinedImpose*>(I.get())){ for(const auto& i: c->imps) setOps.push_back(one(i)); } else setOps.push_back(one(I));
		setStart.push_back((int32_t)setOps.size()); return setOf[I.get()]=(int)setStart.size()-2;
	};
	// NB: vector<double> in `keep` must not reallocate after its data() was taken: reserve generously
	keep.reserve(64);
	vector<int32_t> sphSet; for(const auto& I: sphImpose) sphSet.push_back(setFor(I));
	vector<int32_t> wallSet(wallOf.size(),-1); bool anyWall=false;
	for(size_t i=0;i<wallOf.size();i++){ const auto& imp=wallOf[i]->shape->nodes[0]->getData<DemData>().impose; if(imp){ wallSet[i]=setFor(imp); anyWall=true; } }
	vector<int32_t> groupSet; bool anyGroup=false;
	for(const auto& n: dem.nodes){
		if(!n->hasData<DemData>() || !n->getData<DemData>().isClump()) continue;
		const ClumpData& cd=n->getData<DemData>().cast<ClumpData>();
		const bool sphereClump=std::find(clumpOf.begin(),clumpOf.end(),n)!=clumpOf.end();
		if(sphereClump) continue;      // clumps of spheres: below, in the order of clumpOf (= the order of wdem_upload_clumps)
		groupSet.push_back(cd.impose?setFor(cd.impose):-1); if(cd.impose) anyGroup=true;
	}
	// facet nodes (py/triangulated.py:59-64: every node of a cylinder carries the same StableCircularOrbit); facets of a kinematic clump follow their master node
	vector<int32_t> facetSet(facetOf.size(),-1); bool anyFacet=false;
	for(size_t i=0;i<facetOf.size();i++){
		const auto& nn=facetOf[i]->shape->nodes; const auto& imp=nn[0]->getData<DemData>().impose;
		if(nn[1]->getData<DemData>().impose!=imp || nn[2]->getData<DemData>().impose!=imp) throw std::runtime_error("GpuDemLoop: the three nodes of facet #"+to_string(facetOf[i]->id)+" must carry the same DemData.impose");
		if(imp){ facetSet[i]=setFor(imp); anyFacet=true; }
	}
	vector<int32_t> clumpSet(clumpOf.size(),-1); bool anyClump=false;      // master nodes of clumps of spheres
	for(size_t c=0;c<clumpOf.size();c++){ const auto& imp=clumpOf[c]->getData<DemData>().impose; if(imp){ clumpSet[c]=setFor(imp); anyClump=true; } }
	if(imps.empty()) return;
	chk(h,wdem_upload_imposes(h,(int32_t)imps.size(),imps.data(),(int32_t)setStart.size()-1,setStart.data(),setOps.data(),(int64_t)sphImposeId.size(),sphImposeId.data(),sphSet.data(),anyWall?wallSet.data():nullptr,anyGroup?groupSet.data():nullptr));
	if(anyClump) chk(h,wdem_upload_clump_imposes(h,clumpSet.data()));
	if(anyFacet) chk(h,wdem_upload_facet_imposes(h,facetSet.data()));
}

void GpuDemLoop::pushContacts(){
	DemField& dem=field->cast<DemField>();
	vector<int32_t> ids,cd,sc; vector<double> uN,fl,tl,tr,cp,ph;
	const bool hertz=cfg.law==WDEM_LAW_HERTZ_DMT;
	for(const auto& C: *dem.contacts){
		if(!C->isReal()) continue;
		const L6Geom& g=C->geom->cast<L6Geom>(); const FrictPhys& p=C->phys->cast<FrictPhys>();
		ids.push_back(C->leakPA()->id); ids.push_back(C->leakPB()->id);
		for(int k:{0,1,2}){ cd.push_back(C->cellDist[k]); fl.push_back(p.force[k]); tl.push_back(p.torque[k]); cp.push_back(C->geom->node->pos[k]); }
		sc.push_back(C->stepCreated); uN.push_back(g.uN);
		for(int col:{0,1,2}) for(int row:{0,1,2}) tr.push_back(g.trsf(row,col));          // columns one after another (wdem.h)
		if(!hertz){ ph.push_back(p.kn); ph.push_back(p.kt); ph.push_back(p.tanPhi); ph.push_back(0); ph.push_back(0); }
		else { const HertzPhys& hp=C->phys->cast<HertzPhys>(); ph.push_back(hp.K*sqrt(hp.R)); ph.push_back(hp.kt0); ph.push_back(hp.tanPhi); ph.push_back(hp.alpha_sqrtMK); ph.push_back(hp.alpha==0.?4*M_PI*hp.R*hp.gamma:-1.); }
	}
	if(!uN.empty()) chk(h,wdem_upload_contacts(h,(int64_t)uN.size(),ids.data(),cd.data(),sc.data(),uN.data(),fl.data(),tl.data(),tr.data(),cp.data(),ph.data()));
}

void GpuDemLoop::pull(){
	if(!h) return;
	const size_t n=sphereOf.size(); if(n==0) return;
	vector<double> pos(3*n),ori(4*n),vel(3*n),ang(3*n),frc(3*n),trq(3*n);
	chk(h,wdem_download_state(h,pos.data(),ori.data(),vel.data(),ang.data()));
	chk(h,wdem_download_forces(h,frc.data(),trq.data()));
	for(size_t i=0;i<n;i++){
		Node& nd=*sphereOf[i]->shape->nodes[0]; DemData& d=nd.getData<DemData>();
		nd.pos=Vector3r(pos[3*i],pos[3*i+1],pos[3*i+2]); nd.ori=Quaternionr(ori[4*i+3],ori[4*i],ori[4*i+1],ori[4*i+2]);
		d.vel=Vector3r(vel[3*i],vel[3*i+1],vel[3*i+2]); d.angVel=Vector3r(ang[3*i],ang[3*i+1],ang[3*i+2]);
		d.force=Vector3r(frc[3*i],frc[3*i+1],frc[3*i+2]); d.torque=Vector3r(trq[3*i],trq[3*i+1],trq[3*i+2]);
	}
	if(!wallOf.empty()){
		const size_t nw=wallOf.size(); vector<double> wp(3*nw),wf(3*nw),wt(3*nw);
		chk(h,wdem_download_walls(h,wp.data(),wf.data(),wt.data()));
		for(size_t i=0;i<nw;i++){ Node& nd=*wallOf[i]->shape->nodes[0]; DemData& d=nd.getData<DemData>(); nd.pos=Vector3r(wp[3*i],wp[3*i+1],wp[3*i+2]); d.force=Vector3r(wf[3*i],wf[3*i+1],wf[3*i+2]); d.torque=Vector3r(wt[3*i],wt[3*i+1],wt[3*i+2]); }
	}
	if(!imposeOf.empty() && !wallOf.empty()){      // walls whose velocity an imposition changes
		vector<double> wv(3*wallOf.size()); chk(h,wdem_download_boundary_velocities(h,wv.data(),nullptr,nullptr));
		for(size_t i=0;i<wallOf.size();i++) wallOf[i]->shape->nodes[0]->getData<DemData>().vel=Vector3r(wv[3*i],wv[3*i+1],wv[3*i+2]);
	}
	for(size_t i=0;i<imposeOf.size();i++){      // what the last step's Leapfrog read / accumulated (wdem_get_impose_state)
		double o[8]; chk(h,wdem_get_impose_state(h,(int32_t)i,o));
		if(auto* v=dynamic_cast<VelocityAndReadForce*>(imposeOf[i].get())){ v->sumF.set(o[0]); v->dist=o[1]; }
		else if(auto* f=dynamic_cast<ReadForce*>(imposeOf[i].get())){ f->F.set(Vector3r(o[2],o[3],o[4])); f->T.set(Vector3r(o[5],o[6],o[7])); }
		else if(auto* so=dynamic_cast<StableCircularOrbit*>(imposeOf[i].get())) so->angle=o[1];
		else if(auto* co=dynamic_cast<CircularOrbit*>(imposeOf[i].get())) co->angle=o[1];
	}
	if(!clumpOf.empty()){
		const size_t nc=clumpOf.size(); vector<double> p(3*nc),o(4*nc),v(3*nc),w(3*nc),l(3*nc),f(3*nc),t(3*nc);
		chk(h,wdem_download_clumps(h,p.data(),o.data(),v.data(),w.data(),l.data(),f.data(),t.data()));
		for(size_t i=0;i<nc;i++){
			Node& nd=*clumpOf[i]; ClumpData& d=nd.getData<DemData>().cast<ClumpData>();
			nd.pos=Vector3r(p[3*i],p[3*i+1],p[3*i+2]); nd.ori=Quaternionr(o[4*i+3],o[4*i],o[4*i+1],o[4*i+2]);
			d.vel=Vector3r(v[3*i],v[3*i+1],v[3*i+2]); d.angVel=Vector3r(w[3*i],w[3*i+1],w[3*i+2]); d.angMom=Vector3r(l[3*i],l[3*i+1],l[3*i+2]);
			d.force=Vector3r(f[3*i],f[3*i+1],f[3*i+2]); d.torque=Vector3r(t[3*i],t[3*i+1],t[3*i+2]);
		}
	}
	if(!facetOf.empty()){
		const size_t nf=facetOf.size(); vector<double> v(9*nf);
		chk(h,wdem_download_facets(h,v.data(),nullptr,nullptr));
		for(size_t i=0;i<nf;i++) for(int k:{0,1,2}) facetOf[i]->shape->nodes[k]->pos=Vector3r(v[9*i+3*k],v[9*i+3*k+1],v[9*i+3*k+2]);
	}
}

void GpuDemLoop::pullContacts(){
	if(!h) return;
	DemField& dem=field->cast<DemField>();
	const int64_t n=wdem_download_contacts(h,/*realOnly*/1,0,nullptr,nullptr,nullptr,nullptr,nullptr,nullptr,nullptr,nullptr,nullptr,nullptr,nullptr);
	if(n<0) chk(h,WDEM_ERR_RUNTIME);
	dem.contacts->clear();
	if(n==0) return;
	vector<int32_t> ids(2*n),cd(3*n),real(n),sc(n); vector<double> uN(n),fl(3*n),tl(3*n),tr(9*n),cp(3*n),ph(3*n),raw(5*n);
	if(wdem_download_contacts(h,1,n,ids.data(),cd.data(),real.data(),uN.data(),fl.data(),tl.data(),tr.data(),cp.data(),ph.data(),sc.data(),raw.data())<0) chk(h,WDEM_ERR_RUNTIME);
	const bool hertz=cfg.law==WDEM_LAW_HERTZ_DMT;
	for(int64_t i=0;i<n;i++){
		auto C=make_shared<Contact>();
		C->pA=(*dem.particles)[ids[2*i]]; C->pB=(*dem.particles)[ids[2*i+1]];
		C->cellDist=Vector3i(cd[3*i],cd[3*i+1],cd[3*i+2]); C->stepCreated=sc[i];
		auto g=make_shared<L6Geom>();
		for(int col:{0,1,2}) for(int row:{0,1,2}) g->trsf(row,col)=tr[9*i+3*col+row];
		g->uN=uN[i]; g->node->pos=Vector3r(cp[3*i],cp[3*i+1],cp[3*i+2]); g->node->ori=Quaternionr(g->trsf);
		C->geom=g;
		shared_ptr<FrictPhys> p;
		if(!hertz) p=make_shared<FrictPhys>();
		else { auto hp=make_shared<HertzPhys>(); hp->kt0=raw[5*i+1]; hp->alpha_sqrtMK=raw[5*i+3]; p=hp; }      // K, R: recomputed by Cp2 on request; kn0=K*sqrt(R) is what the law uses
		p->kn=ph[3*i]; p->kt=ph[3*i+1]; p->tanPhi=ph[3*i+2];
		p->force=Vector3r(fl[3*i],fl[3*i+1],fl[3*i+2]); p->torque=Vector3r(tl[3*i],tl[3*i+1],tl[3*i+2]);
		C->phys=p;
		dem.contacts->add(C);
	}
}

void GpuDemLoop::prepare(){
	DemField& dem=field->cast<DemField>();
	// particles added or removed since the upload (ParticleContainer::add/remove mark the contacts dirty, ParticleContainer.cpp:23-73,121-157): re-upload, keeping
	// state and contact history.  (The standalone twin appends / removes spheres in place with wdem_add_spheres / wdem_remove_spheres; here the dirty flag is the signal.)
	if(!h || dem.contacts->dirty || dem.particles->size()!=nParticlesUploaded){ upload(); dem.contacts->dirty=false; return; }
	wdem_config c=makeConfig();
	if(std::memcmp(&c,&cfg,sizeof(c))!=0 && !(std::isnan(c.dt) && std::isnan(cfg.dt))){ chk(h,wdem_set_config(h,&c)); cfg=c; }
	if(scene->isPeriodic){ const Cell& C=*scene->cell; double N[9]; for(int i:{0,1,2}) for(int j:{0,1,2}) N[3*i+j]=C.nextGradV(i,j); chk(h,wdem_set_next_gradv(h,N)); }
}

Real GpuDemLoop::critDt(){ prepare(); double d; chk(h,wdem_crit_dt(h,&d)); return d; }

void GpuDemLoop::run(){
	prepare();
	if(scene->trackEnergy!=trackingOn){ chk(h,wdem_track_energy(h,scene->trackEnergy?1:0)); trackingOn=scene->trackEnergy; }
	wdem_stats st; std::memset(&st,0,sizeof(st));
	chk(h,wdem_step(h,1,&st));          // never touches Python: safe from the background thread without the GIL (Scene.cpp:45-48)
	nRebuilds=st.n_rebuilds; deviceMs=st.device_ms;
	int64_t nf=0; double libDt=NaN;
	chk(h,wdem_get_scalars(h,&libDt,nullptr,nullptr,nullptr,&nf));
	collider->nFullRuns=(int)nf;
	if(dynDt && !dynDt->dead){ int64_t a=-1,b=-1; chk(h,wdem_get_periodic_state(h,&a,&b)); dynDt->stepLast=(long)a; dynDt->stepPrev=(long)b; }
	if(scene->isPeriodic){ double H[9]; chk(h,wdem_get_cell(h,H)); Matrix3r m; for(int i:{0,1,2}) for(int j:{0,1,2}) m(i,j)=H[3*i+j]; scene->cell->setHSize(m); }
	if(scene->trackEnergy){
		double e[8]; chk(h,wdem_get_energy(h,e));
		static const char* names[8]={"kinetic","grav","nonviscDamp","elast","plast","broken","viscN","viscT"};
		for(int k=0;k<8;k++){ int ix=-1; const bool reset=(k==0||k==3); scene->energy->add(e[k]-(reset?0.:scene->energy->getValue(names[k],0.)),names[k],ix,reset?EnergyTracker::IsResettable:EnergyTracker::ZeroDontCreate); }
		if(cfg.evalStress){ double w=0; chk(h,wdem_get_gradv_work(h,&w)); int ix=-1; scene->energy->add(w-scene->energy->getValue("gradV",0.),"gradV",ix,EnergyTracker::IsIncrement|EnergyTracker::ZeroDontCreate); }
	}
	if(contactLoop->evalStress && scene->isPeriodic){ double s9[9]; chk(h,wdem_get_stress(h,s9)); for(int i:{0,1,2}) for(int j:{0,1,2}) contactLoop->stress(i,j)=s9[3*i+j]; }
	// the library has already applied DynDt's result to its own dt; Scene.dt follows through nextDt in Scene::doOneStep's epilogue (Scene.cpp:432)
	cfg.dt=libDt; if(libDt!=scene->dt) scene->nextDt=libDt;
	if(sync){ pull(); }
}
