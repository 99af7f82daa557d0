"""The reference-facing plugin layer (woo_b200.woo: GpuDemLoop + reference-named objects, C++/pybind11 over the C-ABI).
GPU tests are ports of the reference's own tests for this path (py/tests/test_hertz.py, test_pbc.py, test_core.py) with
``woo.utils.defaultEngines`` replaced by ``woo.utils.gpuEngines``; CPU tests cover the object model and the loud failure without a GPU."""
import math
import numpy as np
import pytest
from woo_b200 import woo
from woo_b200.woo.core import Scene
from woo_b200.woo.dem import (DemField, Sphere, Wall, Facet, FrictMat, HertzMat, Leapfrog, InsertionSortCollider, ContactLoop, DynDt, GpuDemLoop, Bo1_Wall_Aabb,
                              Cg2_Sphere_Sphere_L6Geom, Cp2_FrictMat_FrictPhys, Cp2_HertzMat_HertzPhys, Law2_L6Geom_FrictPhys_IdealElPl, Law2_L6Geom_HertzPhys_DMT)


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ------------------------------------------------------------------------------------------------ CPU: object model
def test_reference_defaults():
    # attribute names and defaults of the reference headers (SURVEY §8b "Attributes to mirror")
    assert Leapfrog().damping == .2 and Leapfrog().reset is False
    assert InsertionSortCollider().verletDist == -.05
    cl = ContactLoop()
    assert cl.applyForces is True and cl.updatePhys == 0 and cl.dist00 is True and cl.reorderEvery == 1000 and cl.evalStress is False
    g = Cg2_Sphere_Sphere_L6Geom()
    assert (g.noRatch, g.iniLensTouch, g.trsfRenorm, g.approxMask) == (False, True, 100, 0)
    c = Cp2_HertzMat_HertzPhys()
    assert c.poisson == .2 and math.isnan(c.en)
    l = Law2_L6Geom_FrictPhys_IdealElPl()
    assert (l.noSlip, l.noBreak, l.noFrict, l.iniEqlb, l.relRollStiff, l.relTwistStiff, l.rollTanPhi) == (False, False, False, False, 0., 0., 0.)
    assert Law2_L6Geom_HertzPhys_DMT().noAttraction is True
    d = DynDt()
    assert d.stepPeriod == 100 and d.maxRelInc == 1e-4 and d.dryRun is False
    f = DemField()
    assert f.loneMask == 2 and f.distFactor == -1 and tuple(f.gravity) == (0, 0, 0)
    assert DemField.defaultMovableMask == 0b0101 and DemField.defaultBoundaryMask == 0b0011
    m = woo.utils.defaultMaterial()
    assert (m.density, m.young, m.ktDivKn) == (1e3, 1e7, .2) and m.tanPhi == math.tan(.5)
    S = Scene()
    assert math.isnan(S.dt) and S.dtSafety == .9 and S.step == 0 and S.periodic is False
    assert woo.utils.spherePWaveDt(1e-3, 2400, 30e9) == 2.8284271247461903e-07      # py/utils.py:80-81


def test_particle_container_and_nodes():
    S = Scene(fields=[DemField()])
    ids = S.dem.par.add([woo.utils.sphere((0, 0, 0), .05), woo.utils.sphere((0, .2, 0), .05, fixed=True), woo.utils.wall(0, axis=2, sense=1)])
    assert ids == [0, 1, 2] and len(S.dem.par) == 3
    p = S.dem.par[0]
    assert p.mass == pytest.approx(4 / 3 * math.pi * .05 ** 3 * 1e3) and p.inertia[0] == pytest.approx(.4 * p.mass * .05 ** 2)     # Sphere.cpp:26-33
    assert S.dem.par[1].blocked == "xyzXYZ" and S.dem.par[2].shape.axis == 2 and S.dem.par[2].mask == 3 and p.mask == 5
    assert len(S.dem.nodes) == 1        # only moving nodes are collected (ParticleContainer.cpp:193-198)
    with pytest.raises(RuntimeError):
        S.dem.par.add(p)                # already has an id
    S.dem.nodes[0].pos += (13, 14, 15)
    assert tuple(S.dem.par[0].pos) == (13, 14, 15)
    with pytest.raises(IndexError):
        S.dem.par[7]


def test_particle_container_reuses_freed_ids_like_the_reference():
    """ParticleContainer::findFreeId / remove (ParticleContainer.cpp:23-36,121-157): freed ids are handed out again oldest first; removing the last particle shrinks the container."""
    S = Scene(fields=[DemField()])
    par = S.dem.par
    ids = par.add([woo.utils.sphere((i, 0, 0), .1) for i in range(6)])
    assert list(ids) == [0, 1, 2, 3, 4, 5] and len(par) == 6
    par.remove(1); par.remove(3)
    assert len(par) == 6
    assert par.add(woo.utils.sphere((9, 0, 0), .1)) == 1 and par.add(woo.utils.sphere((9, 1, 0), .1)) == 3 and par.add(woo.utils.sphere((9, 2, 0), .1)) == 6
    par.remove(6); par.remove(5)
    assert len(par) == 5                                   # shrunk past both
    with pytest.raises(IndexError):
        par[5]
    assert par.add(woo.utils.sphere((9, 3, 0), .1)) == 5   # the stale free id 6 is beyond the size: skipped... 5 <= size: taken
    assert [p.id for p in par] == [0, 1, 2, 3, 4, 5]


def test_cpu_engines_do_not_run_on_their_own():
    S = Scene(fields=[DemField()])
    S.dem.par.add(woo.utils.sphere((0, 0, 0), .05))
    S.dt = 1e-5
    S.engines = woo.utils.defaultEngines()
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        S.one()


def test_unsupported_settings_raise():
    S = Scene(fields=[DemField()]); S.dem.par.add(woo.utils.sphere((0, 0, 0), .05)); S.dt = 1e-5
    e = woo.utils.defaultEngines()
    e[2].updatePhys = 1
    S.engines = [GpuDemLoop(*e)]
    with pytest.raises(RuntimeError, match="updatePhys"):
        S.one()
    S.engines = [GpuDemLoop(*woo.utils.defaultEngines(cp2=Cp2_FrictMat_FrictPhys(), law=Law2_L6Geom_HertzPhys_DMT()))]
    with pytest.raises(RuntimeError, match="supported functor pairs"):
        S.one()
    S.engines = woo.utils.gpuEngines(); S.periodic = False
    with pytest.raises(RuntimeError, match="degenerate"):
        S.cell.hSize = [[1, 0, 0], [0, 0, 0], [0, 0, 1]]      # py/tests/test_pbc.py:60-62


@pytest.mark.skipif(_has_cuda(), reason="this box has a GPU")
def test_fails_loudly_without_gpu():
    S = Scene(fields=[DemField()]); S.dem.par.add(woo.utils.sphere((0, 0, 0), .05)); S.dt = 1e-5
    S.engines = woo.utils.gpuEngines()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.one()


# ------------------------------------------------------------------------------------------------ GPU: ports of reference tests
@pytest.mark.gpu
@pytest.mark.parametrize("v0", [1, 1e-2, 1e-4])
def test_hertz_elastic_collision_time(v0):
    """py/tests/test_hertz.py:13-47 (the LawTester is replaced by initial velocities and a Python observer engine)."""
    S = Scene(fields=[DemField()])
    m = HertzMat(density=1e3, young=1e8, ktDivKn=.2, tanPhi=.5, surfEnergy=0, alpha=0)
    S.dem.par.add([woo.utils.sphere((0, 0, 0), .05, mat=m, vel=(0, v0 / 2, 0)), woo.utils.sphere((0, .1 + 1e-9, 0), .05, mat=m, vel=(0, -v0 / 2, 0))])
    S.dtSafety = .4
    seen = {}

    def observer(S):
        real = S.dem.con.existsReal(0, 1)
        if real and "t0" not in seen:
            seen["t0"] = S.time
        if not real and "t0" in seen and "t1" not in seen:
            seen["t1"] = S.time
    S.engines = woo.utils.gpuEngines(damping=0., cp2=Cp2_HertzMat_HertzPhys(en=0, label='cp2', poisson=.2), law=Law2_L6Geom_HertzPhys_DMT(), dynDtPeriod=10) + [observer]
    for _ in range(200000):
        S.one()
        if "t1" in seen:
            break
    E, nu, rho, d = 1e8, .2, 1e3, .1
    Eeff = (E * E) / (E * (1 - nu ** 2) + E * (1 - nu ** 2))
    tau = 2.214 * (rho / Eeff) ** (2 / 5.) * (d / v0 ** (1 / 5.))
    assert abs((seen["t1"] - seen["t0"]) - tau) < .01 * tau + 2 * S.dt


def _pbc_scene(spheres):
    S = Scene(fields=[DemField()])
    S.periodic = True
    S.cell.setBox(1., 1., 1.)
    S.engines = woo.utils.gpuEngines(damping=0.)
    S.dem.par.add([Sphere.make(c, radius=r) for c, r in spheres])
    S.dt = 1e-9
    return S


@pytest.mark.gpu
def test_pbc_collisions():
    # py/tests/test_pbc.py:130-171
    S = _pbc_scene([((.3, .5, .5), .4), ((.7, .9, .5), .1)]); S.one()
    assert tuple(S.dem.con[0, 1].cellDist) == (0, 0, 0)
    S = _pbc_scene([((.3, .5, .5), .4), ((.7, .9, .5), .1)]); S.dem.nodes[0].pos += (13, 14, 15); S.one()
    assert tuple(S.dem.con[0, 1].cellDist) == (13, 14, 15)
    S = _pbc_scene([((.3, .5, .5), .4), ((.8, .9, .5), .2)])
    with pytest.raises(RuntimeError):
        S.one()
    for shift2 in [(0, 0, 0), (1, 2, 3), (6, 8, -4)]:
        S = _pbc_scene([((.5, .5, .5), .2), ((.8, .6, .5), .2)]); S.dem.nodes[1].pos += shift2; S.one()
        assert tuple(S.dem.con[0, 1].cellDist) == tuple(-np.array(shift2))


@pytest.mark.gpu
def test_ten_spheres_contact_force_sign():
    # py/tests/test_core.py:436-453
    S = Scene(fields=[DemField()])
    S.dem.par.add([woo.utils.sphere((x * .95, 0, 0), .5) for x in range(10)])
    S.engines = woo.utils.gpuEngines(damping=0.); S.dt = 1e-5
    S.one()
    c = S.dem.con[4, 5]
    assert c.real and c.phys.force[0] < 0 and c.geom.uN < 0
    assert S.dem.par[0].f[0] < 0 and S.dem.par[9].f[0] > 0 and abs(S.dem.par[4].f[0]) < 1e-9 * abs(S.dem.par[0].f[0])
    assert c.phys.kn == pytest.approx(1 / (2 * .5 / (1e7 * math.pi * .5 ** 2)))       # py/models.py:16-19


@pytest.mark.gpu
def test_plugin_matches_array_level_engine_and_run_fast_path():
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    Sp = sc.random_gas(300, box=.4, r=.02, psd_ratio=2.0, seed=41)
    E = Engine(Sp.params).load(Sp); E.step(120)
    ref = E.spheres()

    def build():
        S = Scene(fields=[DemField(gravity=Sp.params.gravity)])
        m = FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)
        S.dem.par.add([woo.utils.sphere(Sp.pos[i], Sp.radius[i], mat=m, vel=Sp.vel[i]) for i in range(Sp.n)])
        for i in range(Sp.n):
            S.dem.par[i].angVel = Sp.angVel[i]
        S.dem.par.add([Wall.make(w["pos"], axis=w["axis"], sense=w["sense"], mat=m) for w in Sp.walls])
        S.engines = woo.utils.gpuEngines(damping=.2, dynDtPeriod=10)
        return S
    S = build()
    for _ in range(120):
        S.one()
    got = np.array([S.dem.par[i].pos for i in range(Sp.n)])
    assert np.array_equal(got, ref["pos"]) and S.step == 120 and S.dt == E.scalars()["dt"]
    assert np.array_equal(np.array([S.dem.par[i].f for i in range(Sp.n)]), ref["force"])
    assert S.engines[0].collider.nFullRuns == E.scalars()["nFullRuns"]
    S2 = build(); S2.one(); S2.run(119, True)        # first step fixes dt, then one C-ABI call for the rest
    assert np.array_equal(np.array([S2.dem.par[i].pos for i in range(Sp.n)]), ref["pos"]) and S2.step == 120 and S2.time == pytest.approx(S.time, rel=1e-12)


@pytest.mark.gpu
def test_host_edits_reach_the_device():
    S = Scene(fields=[DemField(gravity=(0, 0, -10))])
    S.dem.par.add([woo.utils.sphere((0, 0, 1), .1), woo.utils.wall(0, axis=2, sense=1)])
    S.engines = woo.utils.gpuEngines(damping=0.); S.dt = 1e-4
    S.run(10, True)
    z = S.dem.par[0].pos[2]
    assert z < 1
    S.dem.par[0].vel = (0, 0, 5.)
    S.one()
    assert S.dem.par[0].pos[2] > z and S.dem.par[0].vel[2] == pytest.approx(5. - 10 * 1e-4)
    n = len(S.dem.par); z, vz, step = S.dem.par[0].pos[2], S.dem.par[0].vel[2], S.step
    S.dem.par.add(woo.utils.sphere((1, 0, 1), .1))          # topology change -> re-upload from the latest device state, forces included
    S.one(); S.one()
    assert len(S.dem.par) == n + 1 and S.step == step + 2
    assert S.dem.par[0].vel[2] == pytest.approx(vz - 2 * 10 * 1e-4) and S.dem.par[0].pos[2] > z
    # the new sphere saw zero force in its first Leapfrog, gravity from the second on
    assert S.dem.par[n].vel[2] == pytest.approx(-10 * 1e-4) and S.dem.par[n].pos[2] < 1


@pytest.mark.gpu
def test_contact_history_survives_adding_particles():
    """ParticleContainer.add in mid-run re-uploads the particle set; the real contacts (L6Geom frame, shear force, stepCreated) must survive it like
    the reference's ContactContainer does: the run with a bystander sphere added at step 150 equals the run that had it from the start."""
    from woo_b200 import scenes as sc
    Sp = sc.random_gas(250, box=.4, r=.02, psd_ratio=2.0, seed=43)

    def build(with_bystander):
        S = Scene(fields=[DemField(gravity=Sp.params.gravity)])
        m = FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)
        S.dem.par.add([woo.utils.sphere(Sp.pos[i], Sp.radius[i], mat=m, vel=Sp.vel[i]) for i in range(Sp.n)])
        S.dem.par.add([Wall.make(w["pos"], axis=w["axis"], sense=w["sense"], mat=m) for w in Sp.walls])
        if with_bystander:
            S.dem.par.add(woo.utils.sphere((5., 5., 5.), .02, mat=m, fixed=True))
        S.engines = woo.utils.gpuEngines(damping=.2, dynDtPeriod=0); S.dt = 2e-5
        return S, m
    A, _ = build(True); A.run(300, True)
    B, m = build(False); B.run(150, True)
    nreal = sum(1 for c in B.dem.con if c.real)
    created = {(c.id1, c.id2): c.stepCreated for c in B.dem.con if c.real}
    B.dem.par.add(woo.utils.sphere((5., 5., 5.), .02, mat=m, fixed=True))
    B.run(150, True)
    assert nreal > 20 and B.step == A.step == 300
    for i in range(Sp.n):
        assert tuple(A.dem.par[i].pos) == tuple(B.dem.par[i].pos) and tuple(A.dem.par[i].f) == tuple(B.dem.par[i].f), i
    kept = [(c.id1, c.id2) for c in B.dem.con if c.real and (c.id1, c.id2) in created and c.stepCreated == created[(c.id1, c.id2)]]
    assert len(kept) > 10
    ca = {(c.id1, c.id2): tuple(c.phys.force) for c in A.dem.con if c.real}
    cb = {(c.id1, c.id2): tuple(c.phys.force) for c in B.dem.con if c.real}
    assert ca == cb


@pytest.mark.gpu
def test_energy_tracking_like_the_reference_test():
    """py/tests/test_energy.py:69-90 (TestLeapfrog) with the GPU engine group: Scene(trackEnergy=True), a wall and a sphere under gravity with damping .4."""
    def scene():
        S = Scene(fields=[DemField(gravity=(0, 0, -10))]); S.dtSafety = .9; S.trackEnergy = True
        m = woo.utils.defaultMaterial()
        S.dem.par.add([Wall.make(0, axis=2, sense=1, mat=m), Sphere.make((0, 0, .9), radius=.8, mat=m)])
        S.engines = woo.utils.gpuEngines(damping=.4)
        return S
    S = scene()
    for n in S.dem.nodes:
        n.dem.gravitySkip = True
    S.dem.par[1].vel = (0, 0, -1)
    S.run(400, True)
    assert "grav" not in S.energy and S.energy["nonviscDamp"] != 0.
    S = scene()
    for n in S.dem.nodes:
        n.dem.dampingSkip = True
    S.run(400, True)
    assert S.energy["grav"] != 0. and "nonviscDamp" not in S.energy
    assert "kinetic" in S.energy and "elast" in S.energy


@pytest.mark.gpu
def test_rotating_drum_with_add_clumped():
    """examples/mill.py:38-42 in small: facets clumped to a central node that carries the angular velocity; the drum turns, carries the grains along, and the
    result equals the array-level engine driven through wdem_upload_facet_groups."""
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    from woo_b200.woo.core import Node
    Sp = sc.drum(300, div=12, axial=3, vz0=.5, vfluct=.3, omega=6., rotate=True)
    E = Engine(Sp.params).load(Sp); E.step(400)
    S = Scene(fields=[DemField(gravity=Sp.params.gravity)])
    m = FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)
    S.dem.par.add([woo.utils.sphere(Sp.pos[i], Sp.radius[i], mat=m, vel=Sp.vel[i]) for i in range(Sp.n)])
    S.dem.par.add([Wall.make(w["pos"], axis=w["axis"], sense=w["sense"], mat=m) for w in Sp.walls])
    facets = [Facet.make([tuple(v) for v in f["v"]], mat=m) for f in Sp.facets]
    g = Sp.facet_groups[0]
    central = Node(pos=g["pos"])
    S.dem.par.addClumped(facets, centralNode=central)
    central.dem.angVel = g["angVel"]
    S.engines = woo.utils.gpuEngines(damping=.2, dynDtPeriod=100)
    v0 = np.array(facets[0].shape.nodes[0].pos)
    S.run(400, True)
    assert np.array_equal(np.array([S.dem.par[i].pos for i in range(Sp.n)]), E.spheres()["pos"])
    v1 = np.array(facets[0].shape.nodes[0].pos)
    assert np.array_equal(v1, E.facets()[0, 0]) and np.linalg.norm(v1 - v0) > 1e-3
    assert abs(np.linalg.norm(np.array(central.ori)) - 1) < 1e-12 and abs(central.ori[3]) < 1.      # the master node has turned


@pytest.mark.gpu
def test_rotating_cylinder_with_stable_circular_orbit_on_every_node():
    """py/triangulated.py:59-64 (woo.triangulated.cylinder(angVel=..., realRot=True)): one StableCircularOrbit shared by all nodes of the mesh; the plugin sends it to
    the device as the imposition of the facets' nodes (wdem_upload_facet_imposes) and the result equals the array-level engine (same device code: bit for bit)."""
    import math
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    from woo_b200.woo.core import Node
    from woo_b200.woo.dem import StableCircularOrbit
    om = 6.
    Sp = sc.drum(300, div=12, axial=3, vz0=.5, vfluct=.3, omega=om, rotate=False)
    R = float(np.hypot(Sp.facets[0]["v"][0][1], Sp.facets[0]["v"][0][2]))
    for f in Sp.facets: f["fakeVel"] = (0., 0., 0.)
    q = (0., math.sin(math.pi / 4), 0., math.cos(math.pi / 4))
    sc.attach_impose(Sp, [sc.imp_stable_circular_orbit((0., 0., 0.), om, R, node_ori=q)], facets=range(len(Sp.facets)))
    E = Engine(Sp.params).load(Sp); E.step(400)
    S = Scene(fields=[DemField(gravity=Sp.params.gravity)])
    m = FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)
    S.dem.par.add([woo.utils.sphere(Sp.pos[i], Sp.radius[i], mat=m, vel=Sp.vel[i]) for i in range(Sp.n)])
    S.dem.par.add([Wall.make(w["pos"], axis=w["axis"], sense=w["sense"], mat=m) for w in Sp.walls])
    orbit = StableCircularOrbit(node=Node(pos=(0, 0, 0), ori=q), omega=om, radius=R)
    facets = [Facet.make([tuple(v) for v in f["v"]], mat=m) for f in Sp.facets]
    for f in facets:
        for n in f.shape.nodes: n.dem.impose = orbit
    S.dem.par.add(facets)
    S.engines = woo.utils.gpuEngines(damping=.2, dynDtPeriod=100)
    v0 = np.array(facets[0].shape.nodes[0].pos)
    S.run(400, True)
    assert np.array_equal(np.array([S.dem.par[i].pos for i in range(Sp.n)]), E.spheres()["pos"])
    v1 = np.array(facets[0].shape.nodes[0].pos)
    assert np.array_equal(v1, E.facets()[0, 0]) and np.linalg.norm(v1 - v0) > 1e-3
    assert math.hypot(v1[1], v1[2]) == pytest.approx(R, rel=1e-12)
    assert orbit.angle == pytest.approx(om * S.time, rel=1e-9)


@pytest.mark.gpu
def test_inlet_like_additions_use_the_device_append_path():
    """Spheres added batch by batch while the scene runs (what an Inlet does): the plugin appends them on the device (wdem_add_spheres); the result equals the
    array-level engine fed the same batches, and a batch with a new material falls back to the full re-upload with the contact history kept."""
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    Sp = sc.random_gas(200, box=.4, r=.02, psd_ratio=2.0, seed=91)
    E = Engine(Sp.params).load(Sp)
    S = Scene(fields=[DemField(gravity=Sp.params.gravity)])
    m = FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)
    S.dem.par.add([Wall.make(w["pos"], axis=w["axis"], sense=w["sense"], mat=m) for w in Sp.walls])          # boundaries first: ids 0..5
    S.dem.par.add([woo.utils.sphere(Sp.pos[i], Sp.radius[i], mat=m, vel=Sp.vel[i]) for i in range(Sp.n)])
    for i in range(Sp.n):
        S.dem.par[len(Sp.walls) + i].angVel = Sp.angVel[i]
    S.engines = woo.utils.gpuEngines(damping=.2, dynDtPeriod=0); S.dt = 2e-5
    Sp.boundaries_first = True; Sp.params.dynDtPeriod = 0; Sp.params.dt = 2e-5
    E = Engine(Sp.params).load(Sp)
    rng = np.random.default_rng(92)
    nxt = len(Sp.walls) + Sp.n
    for batch in range(3):
        S.run(60, True); E.step(60)
        k = 25
        pos = rng.uniform(.05, .35, (k, 3)); rad = rng.uniform(.012, .03, k); vel = rng.uniform(-1, 1, (k, 3))
        mat = m if batch < 2 else FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)                    # last batch: another material object
        ids = S.dem.par.add([woo.utils.sphere(pos[i], rad[i], mat=mat, vel=vel[i]) for i in range(k)])
        assert ids == list(range(nxt, nxt + k)); nxt += k
        mass = (4 / 3.) * math.pi * (rad * rad * rad) * 2500.          # Sphere::lumpMassInertia's expression (Sphere.cpp:26-33), as scenes.mass_inertia
        E.add_spheres(np.array(ids, np.int32), pos, rad, mass, (2. / 5.) * mass * (rad * rad), vel=vel)
    S.run(60, True); E.step(60)
    got = np.array([S.dem.par[i].pos for i in range(len(Sp.walls), nxt)])
    assert np.array_equal(got, E.spheres()["pos"])
    assert np.array_equal(np.array([S.dem.par[i].f for i in range(len(Sp.walls), nxt)]), E.spheres()["force"])


@pytest.mark.gpu
def test_outlet_like_removals_use_the_device_path():
    """ParticleContainer.remove in mid-run (what an Outlet does) drops the spheres on the device (wdem_remove_spheres); equals the array-level engine."""
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    Sp = sc.random_gas(250, box=.4, r=.02, psd_ratio=2.0, seed=93); Sp.boundaries_first = True; Sp.params.dynDtPeriod = 0; Sp.params.dt = 2e-5
    E = Engine(Sp.params).load(Sp)
    S = Scene(fields=[DemField(gravity=Sp.params.gravity)])
    m = FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)
    nb = len(Sp.walls)
    S.dem.par.add([Wall.make(w["pos"], axis=w["axis"], sense=w["sense"], mat=m) for w in Sp.walls])
    S.dem.par.add([woo.utils.sphere(Sp.pos[i], Sp.radius[i], mat=m, vel=Sp.vel[i]) for i in range(Sp.n)])
    for i in range(Sp.n):
        S.dem.par[nb + i].angVel = Sp.angVel[i]
    S.engines = woo.utils.gpuEngines(damping=.2, dynDtPeriod=0); S.dt = 2e-5
    S.run(100, True); E.step(100)
    rng = np.random.default_rng(94)
    gone = sorted(int(x) for x in rng.choice(np.arange(nb, nb + Sp.n), 40, replace=False))
    for i in gone:
        S.dem.par.remove(i)
    E.remove_spheres(np.array(gone, np.int32))
    with pytest.raises(IndexError):
        S.dem.par[gone[0]]
    S.run(120, True); E.step(120)
    left = [i for i in range(nb, nb + Sp.n) if i not in set(gone)]
    assert np.array_equal(np.array([S.dem.par[i].pos for i in left]), E.spheres()["pos"])
    assert np.array_equal(np.array([S.dem.par[i].f for i in left]), E.spheres()["force"])
    assert all(c.id1 not in gone and c.id2 not in gone for c in S.dem.con)


# ------------------------------------------------------------------------------------------------ CPU: object model of the later additions
def test_object_model_clumps_removal_energy_stress_cpu():
    from woo_b200.woo.core import Node
    S = Scene(fields=[DemField()])
    m = woo.utils.defaultMaterial()
    ids = S.dem.par.add([woo.utils.sphere((0, 0, i * .2), .05, mat=m) for i in range(4)])
    assert ids == [0, 1, 2, 3] and len(S.dem.nodes) == 4
    S.dem.par.remove(2)
    with pytest.raises(IndexError):
        S.dem.par[2]
    with pytest.raises(IndexError):
        S.dem.par.remove(2)
    assert len(S.dem.nodes) == 3 and S.dem.par.add(woo.utils.sphere((1, 1, 1), .05, mat=m)) == 2      # the freed id is handed out again (findFreeId)
    assert S.dem.par.add(woo.utils.sphere((1, 1, 2), .05, mat=m)) == 4
    facets = [Facet.make([(0, 0, 0), (1, 0, 0), (0, 1, 0)], mat=m), Facet.make([(1, 0, 0), (1, 1, 0), (0, 1, 0)], mat=m)]
    central = Node(pos=(.5, .5, 0))
    assert S.dem.par.addClumped(facets, centralNode=central) is central
    assert [f.id for f in facets] == [5, 6] and central.dem.blocked == "xyzXYZ" and central in S.dem.nodes
    central.dem.angVel = (0, 0, 3.)
    assert tuple(central.dem.angVel) == (0, 0, 3.)
    with pytest.raises(RuntimeError):
        S.dem.par.addClumped(facets, centralNode=None)
    assert S.energy == {} and S.trackEnergy is False
    cl = ContactLoop()
    assert cl.evalStress is False and np.array_equal(cl.stress, np.zeros((3, 3)))


def test_guess_moving():
    """py/tests/test_demfield.py:11-22"""
    from woo_b200.woo.dem import DemData
    assert DemData(blocked='', mass=0).guessMoving() is False           # zero mass, don't move
    assert DemData(blocked='xyz', mass=0).guessMoving() is False
    assert DemData(blocked='xyzXY', mass=1).guessMoving() is True       # something not blocked, move
    assert DemData(blocked='xyzXYZ', mass=1).guessMoving() is False     # everything blocked, not move
    assert DemData(blocked='xyzXYZ', vel=(1, 1, 1), mass=1).guessMoving() is True      # velocity assigned, move


# ---- clumps of spheres with mass: py/tests/test_clump.py TestSimpleClump, against the plugin -----------------------------------------------------------------
def _simple_clump():
    S = Scene(fields=[DemField()])
    S.dem.par.addClumped([woo.utils.sphere((0, 0, 0), 1), woo.utils.sphere((0, 0, 3), .5)])
    for n in (S.dem.par[0].shape.nodes[0], S.dem.par[1].shape.nodes[0]):
        S.dem.nodesAppend(n)
    return (S,) + tuple(S.dem.nodes)


def test_clump_consistency_and_static_properties():
    """test_clump.py:30-64 (testConsistency, testStaticProperties)"""
    S, bC, b1, b2 = _simple_clump()
    assert b1 in bC.dem.nodes and b2 in bC.dem.nodes
    assert bC.dem.clump and b1.dem.clumped and b2.dem.clumped
    assert b1.dem.master == bC and b2.dem.master == bC
    assert bC.dem.mass == b1.dem.mass + b2.dem.mass
    c = (b1.dem.mass * np.array(b1.pos) + b2.dem.mass * np.array(b2.pos)) / bC.dem.mass
    assert np.array_equal(np.array(bC.pos), c)
    i1, i2 = [(8. / 15) * math.pi * S.dem.par[i].mat.density * S.dem.par[i].shape.radius ** 5 for i in (0, 1)]
    iMax = i1 + i2 + b1.dem.mass * np.linalg.norm(np.array(b1.pos) - c) ** 2 + b2.dem.mass * np.linalg.norm(np.array(b2.pos) - c) ** 2
    iC = sorted(bC.dem.inertia)
    assert iC[0] == pytest.approx(i1 + i2, rel=1e-7) and iC[1] == pytest.approx(iMax, rel=1e-7) and iC[2] == pytest.approx(iMax, rel=1e-7)
    # the orientation matches the order of the inertia: the axis of the smallest moment is the line of centres
    x, y, z, w = bC.ori
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    assert abs(abs(R[:, int(np.argmin(bC.dem.inertia))] @ np.array([0, 0, 1.])) - 1) < 1e-12
    assert np.allclose(np.array(bC.pos) + R @ np.array(bC.dem.relPos[1]), b2.pos, atol=1e-14)


@pytest.mark.gpu
def test_clump_member_velocities_assigned_by_leapfrog():
    """test_clump.py:65-81 (testVelocity): dt=0, one step updates the members' velocities from the master node"""
    S, bC, b1, b2 = _simple_clump()
    S.dt = 0
    bC.dem.vel = (1., .2, .4); bC.dem.angVel = (0, .4, .1)
    S.engines = woo.utils.gpuEngines(damping=0., dynDtPeriod=0)
    S.one()
    for b in (b1, b2):
        assert np.array_equal(np.array(b.dem.vel), np.array(bC.dem.vel) + np.cross(np.array(bC.dem.angVel), np.array(b.pos) - np.array(bC.pos)))
        assert np.array_equal(np.array(b.dem.angVel), np.array(bC.dem.angVel))


@pytest.mark.gpu
def test_clump_members_do_not_collide():
    """test_clump.py:82-91 (testNoCollide)"""
    S = Scene(fields=[DemField()])
    S.dem.par.addClumped([woo.utils.sphere(c, r) for c, r in [((1, 0, 0), 1), ((0, 1, 0), 1), ((0, 0, 1), 1)]])
    S.engines = woo.utils.gpuEngines(damping=0., dynDtPeriod=0); S.dt = 1e-5
    S.one()
    assert len(S.dem.con) == 0


@pytest.mark.gpu
def test_clumps_through_the_plugin_equal_the_array_level_engine():
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    Sp = sc.clump_deposition(20, 15)
    S = Scene(fields=[DemField(gravity=Sp.params.gravity)])
    m = FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)
    pars = [woo.utils.sphere(Sp.pos[i], Sp.radius[i], mat=m) for i in range(Sp.n)]
    member = set(i for c in Sp.clumps for i in c["members"])
    masters = []
    # ids in the order of the array-level scene: spheres first (clump members get theirs inside addClumped)
    k = 0
    for c in Sp.clumps:
        assert c["members"][0] == k; k += len(c["members"])
        masters.append(S.dem.par.addClumped([pars[i] for i in c["members"]]))
    S.dem.par.add([pars[i] for i in range(Sp.n) if i not in member])
    S.dem.par.add([Wall.make(w["pos"], axis=w["axis"], sense=w["sense"], mat=m) for w in Sp.walls])
    for i in range(Sp.n):
        if i not in member:
            pars[i].vel = Sp.vel[i]
    # the plugin derives centroid / principal axes itself (Jacobi instead of the scene helper's eigh): feed ITS numbers to the array-level engine
    for c, M in zip(Sp.clumps, masters):
        M.dem.vel = c["vel"]; M.dem.angVel = c["angVel"]
        assert np.allclose(sorted(M.dem.inertia), sorted(c["inertia"]), rtol=1e-10) and np.allclose(M.pos, c["pos"], rtol=1e-13)
        c.update(pos=np.array(M.pos), ori=np.array(M.ori), inertia=np.array(M.dem.inertia), mass=M.dem.mass, relPos=np.array(M.dem.relPos), relOri=np.array(M.dem.relOri))
    E = Engine(Sp.params).load(Sp)
    S.engines = woo.utils.gpuEngines(damping=.2, dynDtPeriod=100); S.dtSafety = Sp.params.dtSafety
    E.step(3000); S.run(3000, True)
    assert np.array_equal(np.array([p.pos for p in pars]), E.spheres()["pos"])
    kc = E.clumps()
    assert np.array_equal(np.array([M.pos for M in masters]), kc["pos"]) and np.array_equal(np.array([M.dem.angVel for M in masters]), kc["angVel"])
    assert np.array_equal(np.array([M.dem.angMom for M in masters]), kc["angMom"])
    # reading the state back (lazy pull of spheres and master nodes) does not disturb the run
    E.step(500); S.run(500, True)
    assert np.array_equal(np.array([p.pos for p in pars]), E.spheres()["pos"])


# ---- Impose through the plugin -----------------------------------------------------------------------------------------------------------------------------
def test_impose_object_model_cpu():
    from woo_b200.woo.dem import (HarmonicOscillation, AlignedHarmonicOscillations, ConstantForce, RadialForce, Local6Dofs, VelocityAndReadForce, ReadForce, VariableVelocity3d,
                                  VariableAlignedRotation, CombinedImpose, DemData)
    from woo_b200.woo.core import Node
    a, b = HarmonicOscillation(freq=10., amp=.1, dir=(0, 0, 2)), ConstantForce(F=(0, 0, 1))
    c = a + b
    assert isinstance(c, CombinedImpose) and len(c.imps) == 2 and len((c + ReadForce()).imps) == 3      # Impose.__add__ (Impose.cpp:302-312)
    assert DemData(blocked='xyzXYZ', mass=1, impose=a).guessMoving() is True                             # DemData::guessMoving: ... || impose (Particle.cpp:638-640)
    with pytest.raises(RuntimeError):
        Local6Dofs(whats=(0, 3, 0, 0, 0, 0))
    with pytest.raises(ValueError):
        VariableAlignedRotation(timeAngVel=[(0, 0), (2, 1), (1, 2)])
    p = woo.utils.sphere((0, 0, 0), .1); p.impose = a
    assert p.shape.nodes[0].dem.impose is a
    v = VelocityAndReadForce(dir=(1, 0, 0), vel=.5)
    assert v.latBlock is True and v.invF is False and v.dist == 0 and v.sumF == 0                        # Impose.hpp:160-168 defaults
    assert VariableVelocity3d(times=[0, 1], vels=[(0, 0, 0), (1, 0, 0)]).diff is False and RadialForce(nodeA=Node(pos=(0, 0, 0)), nodeB=Node(pos=(0, 0, 1)), F=2.).F == 2.


@pytest.mark.gpu
def test_impose_through_the_plugin_equals_the_array_level_engine():
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    from woo_b200.woo.dem import HarmonicOscillation, ConstantForce, VelocityAndReadForce, ReadForce, VariableVelocity3d
    Sp = sc.random_gas(150, box=.4, r=.02, psd_ratio=1.5, seed=12)
    Sp.boundaries_first = True                                                                             # ids: walls 0..5, then the spheres (the order of the adds below)
    nan = float("nan")
    k1 = sc.attach_impose(Sp, [sc.imp_harmonic(20., .01, dir=(1, 1, 0), perpFree=True), sc.imp_constant_force((0, 0, .03))], spheres=range(0, 10))
    k2 = sc.attach_impose(Sp, [sc.imp_velocity_and_read_force((0, 0, 3), -.1, latBlock=False), sc.imp_read_force()], spheres=range(10, 30))
    k3 = sc.attach_impose(Sp, [sc.imp_variable_velocity3d([0, .005, .01], [(nan, nan, 0), (nan, nan, .2), (nan, nan, -.2)], wrap=True)], walls=[4])
    E = Engine(Sp.params).load(Sp)
    S = Scene(fields=[DemField(gravity=Sp.params.gravity)])
    m = FrictMat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)
    walls = [Wall.make(w["pos"], axis=w["axis"], sense=w["sense"], mat=m) for w in Sp.walls]
    wimp = VariableVelocity3d(times=[0, .005, .01], vels=[(nan, nan, 0), (nan, nan, .2), (nan, nan, -.2)], wrap=True)
    walls[4].impose = wimp
    S.dem.par.add(walls)                                                                                   # boundaries first in this scene: ids 0..5
    pars = [woo.utils.sphere(Sp.pos[i], Sp.radius[i], mat=m, vel=Sp.vel[i]) for i in range(Sp.n)]
    i1 = HarmonicOscillation(freq=20., amp=.01, dir=(1, 1, 0), perpFree=True) + ConstantForce(F=(0, 0, .03))
    vr, rf = VelocityAndReadForce(dir=(0, 0, 3), vel=-.1, latBlock=False), ReadForce()
    i2 = vr + rf
    for i in range(0, 10): pars[i].impose = i1
    for i in range(10, 30): pars[i].impose = i2
    S.dem.par.add(pars)
    for i in range(Sp.n):
        pars[i].angVel = Sp.angVel[i]
    assert walls[4].shape.nodes[0] in S.dem.nodes                                                          # an imposition makes the wall's node move (guessMoving)
    S.engines = woo.utils.gpuEngines(damping=Sp.params.damping, dynDtPeriod=Sp.params.dynDtPeriod); S.dt = Sp.params.dt; S.dtSafety = Sp.params.dtSafety
    E.step(600); S.run(600, True)
    assert np.array_equal(np.array([p.pos for p in pars]), E.spheres()["pos"]) and np.array_equal(np.array([p.vel for p in pars]), E.spheres()["vel"])
    assert np.array_equal(np.array(walls[4].pos), E.walls()["pos"][4])
    st = E.impose_state(Sp.impose_sets[k2][0]); st2 = E.impose_state(Sp.impose_sets[k2][1])
    assert vr.dist == st["dist"] and vr.sumF == pytest.approx(st["sumF"], rel=1e-9, abs=1e-12) and np.allclose(rf.F, st2["F"], rtol=1e-9, atol=1e-12)
    # changing an attribute of an Impose in mid-run takes the full re-upload path (state and contact history kept)
    vr.vel = .05
    S.run(50, True)
    assert vr.dist == pytest.approx(st["dist"] + 50 * S.dt * .05, rel=2e-3)      # DynDt changes dt a little every 10 steps
