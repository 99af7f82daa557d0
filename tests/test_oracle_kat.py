"""Known-answer tests pinning the CPU oracle against the reference's own tests and documented formulas
(SURVEY.md §8c "Golden vectors / KATs").  CPU only."""
import math
import numpy as np
import pytest
from woo_b200 import scenes as sc
from _harness import Oracle


def two_spheres(law, v0=0.0, gap=1e-5, r=.05, mat=None, en=0.0):
    S = sc.SceneSpec()
    S.mats = [mat or sc.hertz_mat(density=1e3, young=1e8, ktDivKn=.2, tanPhi=.5)]
    S.pos = np.array([[0, 0, 0], [0, 2 * r + gap, 0]], dtype=float)
    S.radius = np.array([r, r])
    S.vel = np.array([[0, v0 / 2, 0], [0, -v0 / 2, 0]])
    S.params = sc.Params(law=law, damping=0., dtSafety=.4, dynDtPeriod=10, en=en, poisson=.2)
    return S.finalize()


def test_pwave_dt():
    # woo.utils.spherePWaveDt(1e-3,2400,30e9)=2.8284271247461903e-07 (py/utils.py:80-81)
    S = sc.SceneSpec(); S.mats = [sc.frict_mat(density=2400, young=30e9)]
    S.pos = np.zeros((1, 3)); S.radius = np.array([1e-3]); S.finalize()
    assert Oracle(S).pWaveDt() == 2.8284271247461903e-07


@pytest.mark.parametrize("v0", [1, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5])
def test_hertz_collision_time(v0):
    """py/tests/test_hertz.py:30-47 — collision time of two Hertzian spheres within 1 % of
    tau = 2.214 (rho/Eeff)^(2/5) d / v0^(1/5), dtSafety .4, DynDt every 10 steps."""
    S = two_spheres(sc.LAW_HERTZ_DMT, v0=v0, gap=1e-5 * 0 + 1e-9)
    O = Oracle(S)
    E, nu, rho, d = 1e8, .2, 1e3, .1
    Eeff = (E * E) / (E * (1 - nu ** 2) + E * (1 - nu ** 2))
    tau = 2.214 * (rho / Eeff) ** (2 / 5.) * (d / v0 ** (1 / 5.))
    t0 = t1 = None
    for i in range(2000000):
        O.step(1)
        real = O.num_contacts(True)
        if t0 is None and real:
            t0 = O.time
        if t0 is not None and not real:
            t1 = O.time; break
    assert t0 is not None and t1 is not None
    assert abs((t1 - t0) - tau) < 0.01 * tau + 2 * O.dt


def test_cundall_kn_and_hertz_force():
    # linear: kn=1/(r1/(E1 A)+r2/(E2 A)), A=pi min(r)^2 (py/models.py:16-19); kt=ktDivKn*kn
    r, E = .05, 1e7
    S = two_spheres(sc.LAW_IDEALELPL, gap=-1e-4, mat=sc.frict_mat(young=E, ktDivKn=.2, tanPhi=.5))
    S.params.dt = 1e-6
    O = Oracle(S); O.step(1)
    c = O.contacts(True)
    assert len(c["uN"]) == 1
    A = math.pi * r ** 2
    kn = 1 / (r / (E * A) + r / (E * A))
    assert c["phys"][0, 0] == pytest.approx(kn, rel=1e-14)
    assert c["phys"][0, 1] == pytest.approx(.2 * kn, rel=1e-14)
    assert c["force"][0, 0] == pytest.approx(kn * c["uN"][0], rel=1e-14)
    assert c["uN"][0] == pytest.approx(-1e-4, rel=1e-9)
    # Hertz: F = K sqrt(R) delta^(3/2), K=(4/3)E*/..., R=r/2 (py/models.py:43-54)
    S = two_spheres(sc.LAW_HERTZ_DMT, gap=-1e-4); S.params.dt = 1e-7
    O = Oracle(S); O.step(1)
    c = O.contacts(True)
    E, nu = 1e8, .2
    K = (4 / 3.) / (2 * (1 - nu ** 2) / E); R = r / 2
    assert c["force"][0, 0] == pytest.approx(-K * math.sqrt(R) * (1e-4) ** 1.5, rel=1e-8)


def test_update_phys_kn_proportional_to_young():
    # py/tests/test_demfield.py:25-50: kn ∝ young
    k = []
    for E in (1e7, 2e7):
        S = two_spheres(sc.LAW_IDEALELPL, gap=-1e-4, mat=sc.frict_mat(young=E)); S.params.dt = 1e-6
        O = Oracle(S); O.step(1); k.append(O.contacts(True)["phys"][0, 0])
    assert k[1] == pytest.approx(2 * k[0], rel=1e-14)


def test_force_sign_ten_spheres():
    # py/tests/test_core.py:436-453: 10 overlapping spheres in a row -> contact (4,5) exists after one step
    S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
    S.pos = np.array([[x * .95, 0, 0] for x in range(10)], dtype=float); S.radius = np.full(10, .5)
    S.params = sc.Params(dt=1e-5, damping=0.); S.finalize()
    O = Oracle(S); O.step(1)
    c = O.contacts(True)
    ids = [tuple(i) for i in c["ids"]]
    assert (4, 5) in ids
    k = ids.index((4, 5))
    assert c["force"][k, 0] < 0        # compressive normal force on pA in local frame
    s = O.spheres()
    # particle 4 is pushed to -x by 5 and to +x by 3: inner particles in equilibrium, ends are not
    assert s["force"][0, 0] < 0 and s["force"][9, 0] > 0
    assert abs(s["force"][4, 0]) < 1e-9 * abs(s["force"][0, 0])


def pbc_scene(p, r):
    S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
    S.pos = np.array(p, dtype=float); S.radius = np.array(r, dtype=float)
    S.params = sc.Params(dt=1e-9, periodic=1, damping=0.)
    S.cell = dict(hSize=np.eye(3), homoDeform=sc.HOMO_GRADV2)
    return S.finalize()


@pytest.mark.parametrize("shift", [(0, 0, 0), (1, 2, 3), (6, 8, -4)])
def test_pbc_normal_contact(shift):
    # py/tests/test_pbc.py:157-171: C.cellDist == -shift
    S = pbc_scene([(.5, .5, .5), (.8 + shift[0], .6 + shift[1], .5 + shift[2])], [.2, .2])
    O = Oracle(S); O.collider_run()
    c = O.contacts()
    assert len(c["ids"]) == 1 and tuple(c["ids"][0]) == (0, 1)
    assert tuple(c["cellDist"][0]) == tuple(-np.array(shift))


def test_pbc_over_half_and_double_contact():
    # py/tests/test_pbc.py:130-154
    O = Oracle(pbc_scene([(.3, .5, .5), (.7, .9, .5)], [.4, .1])); O.collider_run()
    assert tuple(O.contacts()["cellDist"][0]) == (0, 0, 0)
    O = Oracle(pbc_scene([(.3 + 13, .5 + 14, .5 + 15), (.7, .9, .5)], [.4, .1])); O.collider_run()
    assert tuple(O.contacts()["cellDist"][0]) == (13, 14, 15)
    O = Oracle(pbc_scene([(.3, .5, .5), (.8, .9, .5)], [.4, .2]))
    with pytest.raises(RuntimeError):
        O.collider_run()


def test_pbc_incident_velocity():
    # py/tests/test_pbc.py:86-102: |L6Geom.vel| == |initVel| for a contact across 10 periodic cells (gradV=0)
    S = pbc_scene([(.5, .5, .5), (.5, .5 + .39999, .5 + 10)], [.2, .2])
    S.vel = np.array([[0, 0, 0], [0, 0, 5.]]); S.params.dt = 1e-6
    O = Oracle(S); O.step(2)
    c = O.contacts(True)
    assert len(c["uN"]) == 1
    assert np.linalg.norm(c["vel"][0]) == pytest.approx(5., rel=1e-9)


def test_wall_sphere_leapfrog_400_steps():
    """py/tests/test_energy.py:69-90 scene: wall z=0 sense +1, sphere r=.8 at z=.1, g=(0,0,-10), damping .4, 400 steps.
    The reference asserts on energies; here: the sphere is pushed up, settles towards z≈r, wall force accumulates
     and gravitySkip/dampingSkip change the trajectory."""
    def run(flags):
        S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
        S.pos = np.array([[0, 0, .1]]); S.radius = np.array([.8]); S.flags = np.array([flags], dtype=np.uint32)
        S.walls = [dict(pos=(0, 0, 0), axis=2, sense=1, mat=0, mask=3)]
        S.params = sc.Params(gravity=(0, 0, -10), damping=.4, dtSafety=.9)
        S.finalize()
        O = Oracle(S); O.step(400)
        return O
    O = run(0)
    z = O.spheres()["pos"][0, 2]
    assert .1 < z and np.isfinite(z)
    assert O.wall_forces()[0, 2] < 0          # wall pushed down; never reset
    z_ng = run(sc.GRAVITY_SKIP).spheres()["pos"][0, 2]
    z_nd = run(sc.DAMPING_SKIP).spheres()["pos"][0, 2]
    assert z_ng != z and z_nd != z


def test_dyndt_rate_limit():
    # DynDt.cpp:99-102: dt may grow at most by (1+maxRelInc)^nSteps between runs
    S = two_spheres(sc.LAW_IDEALELPL, gap=-1e-4, mat=sc.frict_mat(young=1e7)); S.params.dt = 1e-8; S.params.dynDtPeriod = 10
    O = Oracle(S); O.step(11)
    assert O.dt == pytest.approx(1e-8 * (1 + 1e-4) ** 10, rel=1e-12)


def test_stress_two_spheres_across_the_cell_boundary():
    """ContactLoop.evalStress (ContactLoop.cpp:218-231,243-244): one contact across the periodic boundary, normal along x: stress_xx = Fn*|dPos_x|/V with
    the branch taken through the boundary (Contact::dPos adds hSize*cellDist, Contact.cpp:54-59); everything else is zero."""
    S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
    S.pos = np.array([(.05, .5, .5), (.96, .5, .5)]); S.radius = np.array([.05, .05])      # overlap 0.01 through x=0/1
    S.params = sc.Params(dt=1e-7, periodic=1, damping=0., dynDtPeriod=0)
    S.cell = dict(hSize=np.diag([1., 2., .5]), homoDeform=sc.HOMO_GRADV2)
    O = Oracle(S.finalize()); O.step(2)
    c = O.contacts(True)
    assert len(c["ids"]) == 1 and tuple(c["cellDist"][0]) == (-1, 0, 0)
    st = O.stress()
    fn = c["force"][0][0]; assert fn < 0                                  # compressive normal force in the local frame
    # F on pA is along the normal (pA -> pB = -x through the boundary): F_x = fn * n_x = -fn; branch_x = -0.09
    assert st[0, 0] == pytest.approx((-fn) * (-.09) / 1.0, rel=1e-6)
    st[0, 0] = 0
    assert np.abs(st).max() < 1e-12 * abs(fn)


# ---------------------------------------------------------------------------------------------------------------
# golden vectors produced by the reference's own Python contact models (py/models.py CundallModel/HertzModel, scripts/make_golden_models.py)
# ---------------------------------------------------------------------------------------------------------------
def _ref_model_rows():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_models.json")))["rows"]


def ref_model_scene(row, law):
    """Two spheres of the golden row at rest, overlapping by delta."""
    S = sc.SceneSpec()
    mk = sc.frict_mat if law == sc.LAW_IDEALELPL else sc.hertz_mat
    S.mats = [mk(density=1e3, young=row["E1"], ktDivKn=.2, tanPhi=.5), mk(density=1e3, young=row["E2"], ktDivKn=.2, tanPhi=.5)]
    S.pos = np.array([[0, 0, 0], [0, 0, row["r1"] + row["r2"] - row["delta"]]], dtype=float)
    S.radius = np.array([row["r1"], row["r2"]]); S.mat = np.array([0, 1], dtype=np.int32)
    S.params = sc.Params(law=law, damping=0., dt=1e-9, dynDtPeriod=0, poisson=row["nu"], en=float("nan"))
    return S.finalize()


def check_against_ref_models(make_engine, contacts_of):
    for row in _ref_model_rows():
        E = make_engine(ref_model_scene(row, sc.LAW_IDEALELPL)); E.step(1)
        c = contacts_of(E)
        assert len(c["uN"]) == 1
        # the reference's CundallModel.kn = 1/(r1/(E1 A)+r2/(E2 A)) vs Cp2_FrictMat_FrictPhys' 1/(1/(E1 A/l0)+1/(E2 A/l1)): same value up to rounding
        assert c["phys"][0, 0] == pytest.approx(row["cundall_kn"], rel=1e-13)
        assert c["force"][0, 0] == pytest.approx(-row["cundall_kn"] * row["delta"], rel=1e-9)
        E = make_engine(ref_model_scene(row, sc.LAW_HERTZ_DMT)); E.step(1)
        c = contacts_of(E)
        assert len(c["uN"]) == 1
        assert c["force"][0, 0] == pytest.approx(-row["hertz_F"], rel=1e-9)                         # F = K sqrt(R) delta^(3/2)
        assert c["phys"][0, 0] == pytest.approx(1.5 * row["hertz_K"] * math.sqrt(row["hertz_R"]) * math.sqrt(row["delta"]), rel=1e-9)   # tangent stiffness dF/d(delta)


def test_oracle_matches_reference_python_models():
    check_against_ref_models(lambda S: Oracle(S), lambda O: O.contacts(True))


# ---------------------------------------------------------------------------------------------------------------
# Scene.trackEnergy: the EnergyTracker entries of this path
# ---------------------------------------------------------------------------------------------------------------
def test_energy_conservation_elastic_collision():
    """Two spheres collide head-on, linear law, no friction/damping/gravity: kinetic + elast stays at the initial kinetic energy (leapfrog: to O(dt^2))."""
    S = two_spheres(sc.LAW_IDEALELPL, v0=1., gap=1e-4, mat=sc.frict_mat(density=1e3, young=1e7, ktDivKn=.2, tanPhi=0.)); S.params.dynDtPeriod = 0; S.params.dt = 2e-6
    O = Oracle(S); O.track_energy(True)
    m = S.mass_inertia()[0][0]; e0 = 2 * .5 * m * .5 ** 2
    seen_elast = 0.
    for _ in range(60):
        O.step(50)
        e = O.energy(); seen_elast = max(seen_elast, e["elast"])
        assert e["kinetic"] + e["elast"] == pytest.approx(e0, rel=5e-3)       # kinetic is the on-step estimate, elast belongs to the end of the step
        assert e["plast"] == 0 and e["nonviscDamp"] == 0 and e["grav"] == 0
    assert seen_elast > .5 * e0 and O.num_contacts(True) == 0       # they did touch and separate again


def test_energy_free_fall_and_damping_like_the_reference_test():
    """py/tests/test_energy.py:69-90 (TestLeapfrog): a wall, a sphere, gravity, damping .4.  grav + kinetic + elast + nonviscDamp (+plast) add up to ~0 (start at
    rest with zero potential); gravitySkip leaves no 'grav', dampingSkip no 'nonviscDamp'."""
    def scene(flags=0):
        S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
        S.pos = np.array([[0, 0, .9]]); S.radius = np.array([.8]); S.flags = np.array([flags], dtype=np.uint32)
        S.walls = [dict(pos=(0, 0, 0), axis=2, sense=1, mat=0, mask=sc.MASK_BOUNDARY)]
        S.params = sc.Params(gravity=(0, 0, -10), damping=.4, dtSafety=.9, dynDtPeriod=100)
        return S.finalize()
    O = Oracle(scene()); O.track_energy(True); O.step(400)
    e = O.energy()
    assert e["grav"] != 0 and e["nonviscDamp"] > 0 and e["kinetic"] >= 0
    tot = sum(e.values()); sabs = sum(abs(v) for v in e.values())
    assert abs(tot) < 2e-2 * sabs, e                                   # EnergyTracker.relErr (EnergyTracker.cpp:41-43) stays small
    S = scene(flags=sc.GRAVITY_SKIP); S.vel[0] = (0, 0, -1)
    O = Oracle(S); O.track_energy(True); O.step(400)
    assert O.energy()["grav"] == 0 and O.energy()["nonviscDamp"] != 0
    O = Oracle(scene(flags=sc.DAMPING_SKIP)); O.track_energy(True); O.step(400)
    assert O.energy()["grav"] != 0 and O.energy()["nonviscDamp"] == 0


# ---------------------------------------------------------------------------------------------------------------
# py/tests/test_pbc.py TestPBC (:16-128): cell integration and the homoDeform modes of Leapfrog.  The reference runs [Leapfrog] alone; the oracle runs the whole
# loop, which makes no difference within the single step these tests look at (contact forces act from the next step on).
# ---------------------------------------------------------------------------------------------------------------
def _pbc_two_spheres(homo, dt):
    S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
    init_pos = np.array([1., 1. + .999999999999999999, 1. + 10 * 3.]); init_vel = np.array([0., 0., 5.])
    S.pos = np.array([[1., 1., 1.], init_pos]); S.radius = np.array([.5, .5]); S.vel = np.array([[0., 0., 0.], init_vel])
    S.params = sc.Params(dt=dt, periodic=1, damping=0., dynDtPeriod=0)
    S.cell = dict(hSize=np.diag([2.5, 2.5, 3.]), gradV=np.zeros((3, 3)), nextGradV=np.diag([0., 0., -1.]), homoDeform=homo)
    return S.finalize(), init_pos, init_vel


def test_pbc_velgrad_changes_hsize():
    # testVelGrad (:33-49): hSize += dt*gradV*hSize per step (gradV is still zero in the first step: nextGradV takes effect in the epilogue), homoDeform='vel2'
    S, _, _ = _pbc_two_spheres(sc.HOMO_VEL_2ND, 1e-3)
    O = Oracle(S)
    h = np.diag([2.5, 2.5, 3.]); g = np.zeros((3, 3))
    for i in range(10):
        h = h + 1e-3 * g @ h
        O.step(1)
        g = np.diag([0., 0., -1.])
    assert np.allclose(O.cell_hsize(), h, rtol=0, atol=5e-8)         # assertAlmostEqual: 7 places


def test_pbc_homothetic_resize_vel_and_pos():
    # testHomotheticResizeVel (:67-77): homoDeform='vel': v_z = v0_z + z0*gradV_zz after one step
    S, p0, v0 = _pbc_two_spheres(sc.HOMO_VEL, 1e-5)
    O = Oracle(S); O.step(1)
    assert O.spheres()["vel"][1, 2] == pytest.approx(v0[2] + p0[2] * -1., abs=5e-8)
    # testHomotheticResizePos (:78-85): homoDeform='pos' with dt=0: velocity untouched, position moved by z0*gradV_zz*dt (= 0)
    S, p0, v0 = _pbc_two_spheres(sc.HOMO_POS, 0.)
    O = Oracle(S); O.step(1)
    s = O.spheres()
    assert s["vel"][1, 2] == pytest.approx(v0[2], abs=5e-8) and s["pos"][1, 2] == pytest.approx(p0[2] + p0[2] * -1. * 0., abs=5e-8)


@pytest.mark.parametrize("homo", [sc.HOMO_VEL_2ND, sc.HOMO_POS])
def test_pbc_kinetic_energy_is_of_the_fluctuation_velocity(homo):
    # testKineticEnergy / testKineticEnergy_homoPos (:104-121): Ek counts the fluctuation velocity only, not the velocity gradient
    S, p0, v0 = _pbc_two_spheres(homo, 0.)
    O = Oracle(S); O.step(1)
    m = S.mass_inertia()[0][1]
    assert O.kineticEnergy() == pytest.approx(.5 * m * float(v0 @ v0), rel=1e-7)


def test_dyndt_two_sphere_analytic():
    """DynDt::nodalCritDtSq (DynDt.cpp:16-58) for two equal spheres in contact along y with kt = .2 kn: translational stiffness (kt, kn, kt), rotational
    r^2 kt (1,0,1) => the normal DOF is critical: critDt = sqrt(2 m / kn); DynDt::run (:90-113) then sets nextDt = dtSafety*critDt when the rate limit does not bind."""
    S = two_spheres(sc.LAW_IDEALELPL, gap=-1e-4, mat=sc.frict_mat(density=1e3, young=1e7, ktDivKn=.2, tanPhi=.5))
    S.params.dt = 1e-7; S.params.dynDtPeriod = 1; S.params.maxRelInc = 1e6; S.params.dtSafety = .9
    O = Oracle(S); O.step(2)            # step 0 starts the period count, step 1 computes nextDt, which is Scene.dt from step 2 on
    kn = O.contacts(True)["phys"][0, 0]; m = S.mass_inertia()[0][0]
    assert O.dt == pytest.approx(.9 * math.sqrt(2 * m / kn), rel=1e-12)
    # with stiffer tangential springs the rotational DOF takes over: I/(r^2 kt) = .4 m/kt < m/kn for kt > .4 kn
    S = two_spheres(sc.LAW_IDEALELPL, gap=-1e-4, mat=sc.frict_mat(density=1e3, young=1e7, ktDivKn=.8, tanPhi=.5))
    S.params.dt = 1e-7; S.params.dynDtPeriod = 1; S.params.maxRelInc = 1e6; S.params.dtSafety = .9
    O = Oracle(S); O.step(2)            # step 0 starts the period count, step 1 computes nextDt, which is Scene.dt from step 2 on
    c = O.contacts(True); kn, kt = c["phys"][0, 0], c["phys"][0, 1]
    assert kt == pytest.approx(.8 * kn, rel=1e-14)
    assert O.dt == pytest.approx(.9 * math.sqrt(2 * .4 * m / kt), rel=1e-12)


def test_leapfrog_damping_gravity_analytic():
    """Leapfrog (Leapfrog.cpp:33-48,104-255) on one free sphere: the first step sees zero force (forces are those of the previous step); from then on
    f = m g and Cundall's non-viscous damping scales the acceleration by (1 - damping*sign(f*(v + dt/2 a))) = 1 - damping for a sphere accelerating with the force."""
    def run(flags):
        S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
        S.pos = np.array([[0., 0., 1.]]); S.radius = np.array([.1]); S.flags = np.array([flags], dtype=np.uint32)
        S.params = sc.Params(gravity=(0, 0, -10.), damping=.4, dt=1e-4, dynDtPeriod=0)
        O = Oracle(S.finalize()); O.step(11)
        return O.spheres()
    s = run(0)
    assert s["vel"][0, 2] == pytest.approx(10 * -10. * (1 - .4) * 1e-4, rel=1e-12)
    assert s["pos"][0, 2] == pytest.approx(1. + sum(k * -10. * .6 * 1e-4 for k in range(1, 11)) * 1e-4, rel=1e-14)
    assert run(sc.DAMPING_SKIP)["vel"][0, 2] == pytest.approx(10 * -10. * 1e-4, rel=1e-12)
    assert run(sc.GRAVITY_SKIP)["vel"][0, 2] == 0.
    assert run(4)["vel"][0, 2] == 0. and run(4)["pos"][0, 2] == 1.          # z translation blocked (DemData flag DOF_Z)


def test_coulomb_limit_and_noslip():
    """Law2_L6Geom_FrictPhys_IdealElPl (IdealElPl.cpp:49-77): the shear force grows by kt*velT*dt per step until |Ft| reaches |Fn|*tanPhi and stays on the cone;
    with noSlip it keeps growing.  Two spheres with all DOFs blocked, one sliding past the other at constant velocity."""
    def run(noSlip, steps):
        S = two_spheres(sc.LAW_IDEALELPL, gap=-1e-4, mat=sc.frict_mat(density=1e3, young=1e7, ktDivKn=.2, tanPhi=.5))
        S.flags[:] = sc.DOF_ALL; S.vel[1] = (1., 0., 0.)
        S.params.dt = 1e-6; S.params.dynDtPeriod = 0; S.params.noSlip = noSlip
        O = Oracle(S); O.step(steps)
        return O.contacts(True)
    c = run(0, 600)
    fn, ft = c["force"][0, 0], math.hypot(c["force"][0, 1], c["force"][0, 2])
    assert fn < 0 and ft == pytest.approx(abs(fn) * .5, rel=1e-12)
    c2 = run(1, 600)
    ft2 = math.hypot(c2["force"][0, 1], c2["force"][0, 2])
    kt = c2["phys"][0, 1]
    assert ft2 > ft and ft2 == pytest.approx(kt * 1. * 599 * 1e-6, rel=2e-2)      # kt * |velT| * t (the frame turns a little as the sphere slides past)
    early = run(0, 50)
    assert math.hypot(early["force"][0, 1], early["force"][0, 2]) == pytest.approx(early["phys"][0, 1] * 1. * 49 * 1e-6, rel=1e-3)   # below the cone: elastic


def test_sphere_resting_on_wall_equilibrium():
    """A sphere settling on a Wall under gravity with non-viscous damping comes to rest where kn*|uN| = m*g, with the Wall-Sphere stiffness kn = E*pi*r/2
    (Cg2_Wall_Sphere_L6Geom passes radii (-r, r), L6Geom.lens = (r, r), contA = pi r^2: Wall.cpp:82, L6Geom.cpp:61, FrictMat.cpp:20-44)."""
    r, E, rho, g = .1, 1e7, 1e3, 10.
    S = sc.SceneSpec(); S.mats = [sc.frict_mat(density=rho, young=E)]
    S.pos = np.array([[0., 0., r + 1e-4]]); S.radius = np.array([r])
    S.walls = [dict(pos=(0, 0, 0), axis=2, sense=1, mat=0, mask=sc.MASK_BOUNDARY)]
    S.params = sc.Params(gravity=(0, 0, -g), damping=.4, dtSafety=.5, dynDtPeriod=100)
    O = Oracle(S.finalize()); O.step(60000)
    c = O.contacts(True)
    assert len(c["uN"]) == 1
    kn = c["phys"][0, 0]; m = S.mass_inertia()[0][0]
    assert kn == pytest.approx(E * math.pi * r / 2, rel=1e-13)
    assert abs(O.spheres()["vel"][0, 2]) < 1e-9
    assert -c["uN"][0] == pytest.approx(m * g / kn, rel=1e-6)
    assert O.spheres()["force"][0, 2] == pytest.approx(0., abs=1e-6 * m * g)


@pytest.mark.parametrize("centre,cp,region", [((.25, .25, .19), (.25, .25, 0.), "face"), ((.5, -.1, .1), (.5, 0., 0.), "edge"), ((1.1, -.1, .05), (1., 0., 0.), "vertex")])
def test_facet_sphere_closest_point(centre, cp, region):
    """Cg2_Facet_Sphere_L6Geom::go (Facet.cpp:333-391): the contact point is the point of the triangle closest to the sphere centre (face / edge / vertex regions),
    the normal points from it to the centre and uN = distance - radius (halfThick = 0)."""
    r = .2
    S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
    S.pos = np.array([centre], dtype=float); S.radius = np.array([r])
    S.facets = [dict(v=np.array([(0., 0., 0.), (1., 0., 0.), (0., 1., 0.)]), halfThick=0., mat=0, mask=sc.MASK_BOUNDARY)]
    S.params = sc.Params(dt=1e-9, damping=0., dynDtPeriod=0)
    O = Oracle(S.finalize()); O.step(1)
    c = O.contacts(True)
    assert len(c["uN"]) == 1, region
    d = np.array(centre) - np.array(cp); dist = np.linalg.norm(d)
    assert np.allclose(c["contPt"][0], cp, rtol=0, atol=1e-12)
    assert c["uN"][0] == pytest.approx(dist - r, rel=1e-9)
    assert np.allclose(c["trsf"][0][:3], d / dist, rtol=0, atol=1e-9)         # first column of trsf = contact normal
    assert c["force"][0, 0] == pytest.approx(c["phys"][0, 0] * c["uN"][0], rel=1e-12)


@pytest.mark.parametrize("en", [.3, .5, .8])
def test_hertz_restitution_coefficient(en):
    """Cp2_HertzMat_HertzPhys.en -> alpha_sqrtMK (Hertz.cpp:63-75, Antypov & Elliott 2012): the viscous term is built so that a head-on collision of two spheres
    loses velocity by the factor en, independently of the impact speed — with Law2_L6Geom_HertzPhys_DMT.noAttraction=False.  The default noAttraction=True clips the
    viscous force where it would pull the spheres together at the end of the collision (Hertz.cpp:177), so less is dissipated and the rebound is faster."""
    def rebound(v0, noAttraction):
        S = two_spheres(sc.LAW_HERTZ_DMT, v0=v0, gap=1e-9, en=en); S.params.noAttraction = noAttraction
        O = Oracle(S)
        touched = False
        for _ in range(400000):
            O.step(20)
            real = O.num_contacts(True)
            touched = touched or real
            if touched and not real:
                break
        s = O.spheres()
        assert touched
        return (s["vel"][1, 1] - s["vel"][0, 1]) / v0
    for v0 in (1., 1e-2):
        free = rebound(v0, 0)
        assert free == pytest.approx(en, rel=3e-2), (en, v0)
        assert rebound(v0, 1) >= free * (1 - 1e-9)


@pytest.mark.parametrize("ma,mb,collide", [(0b01, 0b10, False), (0b10, 0b10, False), (0b0101, 0b0011, True), (0b0110, 0b0010, False), (0b0111, 0b0011, False), (0b0101, 0b0111, True)])
def test_collision_masks(ma, mb, collide):
    """Collider::mayCollide (Collision.cpp:21-43; masks Particle.hpp:339-349): two particles may collide iff their masks share a bit and none of the shared
    bits lies in DemField.loneMask (0b0010 by default: boundary particles, which all carry it, do not collide with each other)."""
    S = two_spheres(sc.LAW_IDEALELPL, gap=-1e-3, mat=sc.frict_mat())
    S.mask[:] = (ma, mb); S.params.dt = 1e-7; S.params.dynDtPeriod = 0
    O = Oracle(S); O.step(2)
    assert (O.num_contacts(False) == 1) == collide and (O.num_contacts(True) == 1) == collide


def test_gradv_work_two_spheres_compressed_across_the_cell():
    """ContactLoop.cpp:243-254: with trackEnergy and evalStress the work of the velocity gradient on the contact stress is booked as "gradV".  Two blocked spheres
    in a cell shrinking along x at a constant rate, positions following the cell (homoDeform "pos"): the contact force is kn*uN along x, so
    dW = -gradV_xx * (F_x * branch_x / V) * dt * V summed over the steps = minus the elastic energy stored in the contact: the entry closes the energy balance
    (elast + gradV = const for this frictionless, motionless pair)."""
    import math
    S = sc.SceneSpec()
    S.mats = [sc.frict_mat(density=1000., young=1e6, tanPhi=0.)]
    L, r = 1.0, .199
    S.pos = np.array([[.2, .5, .5], [.8, .5, .5]]); S.radius = np.array([r, r])      # apart by .6 inside the cell, by .4 across its x boundary
    S.flags = np.array([sc.DOF_ALL, sc.DOF_ALL], dtype=np.uint32)
    rate = -.2
    G = np.diag([rate, 0., 0.])
    S.params = sc.Params(periodic=1, damping=0., dynDtPeriod=0, dt=1e-3, verletDist=.05)
    S.cell = dict(hSize=np.eye(3) * L, gradV=G, nextGradV=G, homoDeform=sc.HOMO_POS)      # positions follow the cell affinely (the formula assumes an affine branch)
    O = Oracle(S.finalize()); O.track_energy(True); O.eval_stress(True)
    O.step(400)
    c = O.contacts(True)
    assert len(c["ids"]) >= 1 and (c["uN"] < 0).all()
    kn = c["phys"][0, 0]
    elastic = sum(.5 * kn * u * u for u in c["uN"])
    assert O.energy()["elast"] == pytest.approx(elastic, rel=1e-12)
    # the step-wise trapezoid sum lags the spring energy by about one step's increment
    assert O.gradv_work() == pytest.approx(-elastic, rel=2e-2)


# ---------------------------------------------------------------------------------------------------------------
# clumps of spheres with mass: ClumpData::forceTorqueFromMembers / applyToMembers (Clump.cpp:272-307), Leapfrog's aspherical integrator
# (Leapfrog.cpp:299-331)
# ---------------------------------------------------------------------------------------------------------------
def _dumbbells(n=1, omega=(0., 0., 0.), vel=(0., 0., 0.), gravity=(0., 0., 0.), damping=0., r=.05, sep=.1, asym=False):
    S = sc.SceneSpec()
    S.mats = [sc.frict_mat(density=1000., young=1e6, tanPhi=.4)]
    pos, rad, groups = [], [], []
    for i in range(n):
        c = np.array([.5 * i, 0., 1.])
        if asym:      # three unequal spheres in an L: three different principal moments
            pts, rr = [c, c + [sep, 0, 0], c + [0, 1.3 * sep, .2 * sep]], [r, .8 * r, .6 * r]
        else:
            pts, rr = [c - [sep / 2, 0, 0], c + [sep / 2, 0, 0]], [r, r]
        groups.append(list(range(len(pos), len(pos) + len(pts)))); pos += pts; rad += rr
    S.pos, S.radius = np.array(pos), np.array(rad)
    S.params = sc.Params(gravity=gravity, damping=damping, dynDtPeriod=0, dt=1e-4)
    S.finalize()
    return sc.make_clumps(S, groups, vel=[vel] * n, angVel=[omega] * n)


def test_clump_free_rotation_conserves_angular_momentum_and_shape():
    S = _dumbbells(asym=True, omega=(3., -2., 5.), vel=(.1, .2, -.3))
    c0 = S.clumps[0]
    O = Oracle(S)

    def L_global(cl, k=0):
        q = cl["ori"][k]; x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        return R @ (c0["inertia"] * (R.T @ cl["angVel"][k]))
    # Leapfrog.cpp:303-305 initialises DemData.angMom as the LOCAL I*(ori^-1*angVel) and lines 308-309 then use it as the GLOBAL angular momentum:
    # what the reference conserves from the first step on is that vector (it equals the true R*I*R^T*angVel only when ori is the identity).  The oracle restates the code, not the intent.
    x, y, z, w = c0["ori"]
    R0 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    L0 = c0["inertia"] * (R0.T @ c0["angVel"])
    d0 = np.linalg.norm(S.pos[1] - S.pos[0])
    O.step(5000)
    cl, s = O.clumps(), O.spheres()
    assert np.allclose(L_global(cl), L0, rtol=0, atol=1e-3 * np.abs(L0).max())               # torque-free: the angular momentum stands still (ori is at n+1, angVel at n+1/2: O(dt*|w|) apart)
    assert np.linalg.norm(s["pos"][1] - s["pos"][0]) == pytest.approx(d0, rel=1e-12)         # rigid
    assert np.allclose(cl["pos"][0], c0["pos"] + 5000 * 1e-4 * np.array([.1, .2, -.3]), rtol=1e-12)
    assert not np.allclose(cl["angVel"][0], c0["angVel"], rtol=1e-3)                          # an asymmetric top: the angular velocity itself wanders
    assert np.allclose(s["vel"][2], cl["vel"][0] + np.cross(cl["angVel"][0], s["pos"][2] - cl["pos"][0]), rtol=1e-12, atol=1e-14)


def test_clump_falls_like_a_point_mass_and_lands_on_a_wall():
    S = _dumbbells(gravity=(0, 0, -9.81), damping=0.)
    S.walls = [dict(pos=(0, 0, .9), axis=2, sense=1, mat=0, mask=sc.MASK_BOUNDARY)]
    S.params.damping = .3
    O = Oracle(S)
    O.step(100)       # still in the air: z = z0 - g t^2/2 (leapfrog: first step without gravity, force reset after the integration)
    t = 100 * 1e-4
    assert O.clumps()["pos"][0, 2] == pytest.approx(1. - .5 * 9.81 * t * (t - 1e-4) * (1 - .3), rel=1e-9)      # Cundall damping acts against the motion: factor (1 - damping) while accelerating
    O.step(40000)
    cl, c = O.clumps(), O.contacts(True)
    assert len(c["ids"]) == 2 and abs(cl["vel"][0, 2]) < 1e-3                                    # both spheres rest on the wall
    fz = c["force"][:, 0].sum()                                                                   # normal forces carry the clump's weight
    assert abs(fz) == pytest.approx(S.clumps[0]["mass"] * 9.81, rel=1e-2)


# ---------------------------------------------------------------------------------------------------------------
# Impose (Particle.hpp:130-160, Impose.hpp, Impose.cpp; hooks in Leapfrog.cpp:203-205,238,244,248)
# ---------------------------------------------------------------------------------------------------------------
def _free_spheres(n=1, gravity=(0, 0, 0), damping=0., dt=1e-4, r=.05):
    S = sc.SceneSpec()
    S.mats = [sc.frict_mat(density=1000., young=1e6, tanPhi=.4)]
    S.pos = np.array([[.5 * i, 0., 1.] for i in range(n)]); S.radius = np.full(n, r)
    S.params = sc.Params(gravity=gravity, damping=damping, dynDtPeriod=0, dt=dt)
    return S.finalize()


def test_impose_harmonic_oscillation_follows_a_sine():
    S = _free_spheres(2, gravity=(0, 0, -9.81))
    f, A = 5., .02
    sc.attach_impose(S, [sc.imp_harmonic(f, A, dir=(1, 0, 0))], spheres=[0])                    # all other velocity components zeroed: gravity has no effect
    sc.attach_impose(S, [sc.imp_harmonic(f, A, dir=(0, 3, 0), perpFree=True)], spheres=[1])      # only the y component is prescribed: the sphere falls freely in z
    O = Oracle(S)
    N = 1000; O.step(N); t = N * 1e-4
    s = O.spheres()
    assert s["pos"][0, 0] == pytest.approx(A * math.sin(2 * math.pi * f * t), abs=2e-4) and s["pos"][0, 2] == 1. and s["pos"][0, 1] == 0.
    assert s["vel"][0, 0] == pytest.approx(A * 2 * math.pi * f * math.cos(2 * math.pi * f * (t - 1e-4)), rel=1e-12)     # evaluated with the time of the last step
    assert s["pos"][1, 1] == pytest.approx(A * math.sin(2 * math.pi * f * t), abs=2e-4)
    assert s["pos"][1, 2] == pytest.approx(1. - .5 * 9.81 * t * (t - 1e-4), rel=1e-9)


def test_impose_forces_constant_radial_local6():
    S = _free_spheres(3)
    m = S.mass_inertia()[0]
    sc.attach_impose(S, [sc.imp_constant_force((0, 2., 0))], spheres=[0])
    S.pos[1] = (.5, .3, 1.)
    sc.attach_impose(S, [sc.imp_radial_force((.5, 0, 0), (.5, 0, 5), -1.5)], spheres=[1])        # negative: towards the axis x=.5,y=0
    sc.attach_impose(S, [sc.imp_local6((0, 0, 3., 0, 0, 0), (0, 0, 2, 0, 0, 0)), sc.imp_constant_force((1., 0, 0))], spheres=[2])       # CombinedImpose: later ones see the earlier ones' result
    O = Oracle(S)
    O.step(1)
    f = O.spheres()["force"]
    assert np.array_equal(f[0], [0, 2., 0])
    assert np.allclose(f[1], [0, -1.5, 0], atol=1e-15)
    assert np.array_equal(f[2], [1., 0, 3.])
    O.step(100); t = 100 * 1e-4
    assert O.spheres()["vel"][0, 1] == pytest.approx(2. / m[0] * t, rel=1e-9)


def test_impose_velocity_and_read_force_on_a_pushed_sphere():
    """A sphere driven against a fixed one at constant velocity: sumF is the contact force along dir (of the previous step), dist the path."""
    S = _free_spheres(2)
    S.pos[1] = (.1 + 1e-6, 0, 1.); S.flags[1] = 63
    k = sc.attach_impose(S, [sc.imp_velocity_and_read_force((2, 0, 0), .05, invF=True), sc.imp_read_force()], spheres=[0])
    O = Oracle(S)
    O.step(200)
    st, st2 = O.impose_state(S.impose_sets[k][0]), O.impose_state(S.impose_sets[k][1])
    s, c = O.spheres(), O.contacts(True)
    assert st["dist"] == pytest.approx(200 * 1e-4 * .05, rel=1e-12) and s["pos"][0, 0] == pytest.approx(st["dist"], rel=1e-9)      # a free node: the velocity is imposed before the first translation
    assert len(c["ids"]) == 1 and st["sumF"] > 0
    O.step(1)
    # what step 201 read is the force left by step 200's contact loop
    assert O.impose_state(S.impose_sets[k][0])["sumF"] == pytest.approx(-s["force"][0, 0], rel=1e-12)
    assert np.allclose(O.impose_state(S.impose_sets[k][1])["F"], s["force"][0], rtol=1e-12)
    assert st2["F"][0] == pytest.approx(-st["sumF"], rel=1e-12)


def test_impose_variable_velocity_on_a_wall_and_table_semantics():
    S = _free_spheres(1, gravity=(0, 0, -9.81))
    S.walls = [dict(pos=(0, 0, .9), axis=2, sense=1, mat=0, mask=sc.MASK_BOUNDARY)]
    S.finalize()
    nan = float("nan")
    sc.attach_impose(S, [sc.imp_variable_velocity3d([0., .01, .02], [(0, 0, 0), (nan, 0, .5), (nan, 0, .5)])], walls=[0])
    O = Oracle(S)
    O.step(300)       # t = .03: ramp 0 -> .5 over [0,.01], then constant
    w = O.walls()
    z_expected = .9 + (.5 * .5 * .01 + .5 * .02)
    assert w["pos"][0, 2] == pytest.approx(z_expected, abs=2 * .5 * 1e-4)
    assert O.nFullRuns > 1


def test_impose_circular_orbits_keep_radius_and_turn_by_omega_t():
    """StableCircularOrbit (Impose.cpp:38-48) keeps the node at `radius` around the local z axis and advances its polar angle by omega*dt per step; CircularOrbit
    (Impose.cpp:22-37) prescribes the tangential velocity at the approximate mid-step angle (radius not enforced, drifts slowly); a facet whose three nodes carry a
    StableCircularOrbit turns rigidly (py/triangulated.py:59-64)."""
    S = _free_spheres(2, gravity=(0, 0, -9.81))
    q = (0., math.sin(math.pi / 4), 0., math.cos(math.pi / 4))          # local z -> global x (Quaternion.setFromTwoVectors(UnitZ, UnitX))
    c = np.array([.3, .1, .2]); R, om, dt = .25, 7., 1e-4
    S.pos[0] = c + (.05, R * math.cos(.4), R * math.sin(.4)); S.pos[1] = c + (-.02, 0., .1)
    S.facets = [dict(v=np.array([c + (0., R, 0.), c + (.1, R, 0.), c + (0., 0., R)]), halfThick=0., fakeVel=(0., 0., 0.), mat=0, mask=sc.MASK_BOUNDARY)]
    S.finalize()
    sc.attach_impose(S, [sc.imp_stable_circular_orbit(c, om, R, node_ori=q, rotate=True)], spheres=[0], facets=[0])
    sc.attach_impose(S, [sc.imp_circular_orbit(c, om, node_ori=q)], spheres=[1])
    O = Oracle(S)
    N = 2000; O.step(N)
    s = O.spheres()
    # a free node gets the velocity imposed before it is translated (Leapfrog.cpp:203-205): N steps = N increments of the angle
    rel = s["pos"][0] - c
    assert rel[0] == pytest.approx(.05, abs=1e-13) and math.hypot(rel[1], rel[2]) == pytest.approx(R, rel=1e-13)
    # local frame: x_loc = -z_glob... rotate back with q*: for this q, local (x,y,z) = (-gz, gy, gx)
    th = math.atan2(rel[1], -rel[2])
    th0 = math.atan2(R * math.cos(.4), -R * math.sin(.4))
    assert math.remainder(th - th0 - om * N * dt, 2 * math.pi) == pytest.approx(0., abs=1e-10)
    assert np.allclose(s["angVel"][0], (om, 0, 0), atol=1e-14)           # rotate: node.ori * (0,0,omega)
    assert O.impose_state(0)["dist"] == pytest.approx(om * N * dt, rel=1e-12)      # .angle
    rel1 = s["pos"][1] - c
    assert rel1[0] == pytest.approx(-.02, abs=1e-13) and math.hypot(rel1[1], rel1[2]) == pytest.approx(.1, rel=1e-3)
    f = O.facets()[0]
    for i, r_ in enumerate((R, R, R)):
        assert math.hypot(f[i][1] - c[1], f[i][2] - c[2]) == pytest.approx(r_, rel=1e-13)
    assert np.linalg.norm(f[1] - f[0]) == pytest.approx(.1, rel=1e-12)    # rigid: the edge along the axis keeps its length
    # a fully blocked node is translated first and gets its velocity at the end of the step (Leapfrog.cpp:195-199,248): N steps = N-1 increments
    assert math.remainder(math.atan2(f[0][1] - c[1], -(f[0][2] - c[2])) - math.atan2(R, 0.) - om * (N - 1) * dt, 2 * math.pi) == pytest.approx(0., abs=1e-10)
    assert O.nFullRuns > 1


def test_impose_interpolated_motion_reaches_the_given_poses_at_the_given_times():
    """InterpolatedMotion (Impose.cpp:111-133) on the master node of a facet group: the velocities imposed at the end of a step take the node to the interpolated pose
    one step later, so at t = times[k] + dt the node is at poss[k], oris[k] (to the rounding of pos += vel*dt and of the rotation by angVel*dt)."""
    S = _free_spheres(1)
    c = np.array([.3, .1, .2]); dt = 1e-4
    S.facets = [dict(v=np.array([c + (0., .1, 0.), c + (.1, .1, 0.), c + (0., 0., .1)]), halfThick=0., fakeVel=(0., 0., 0.), mat=0, mask=sc.MASK_BOUNDARY)]
    S.facet_groups = [dict(pos=tuple(c), ori=(0., 0., 0., 1.), vel=(0., 0., 0.), angVel=(0., 0., 0.), facets=[0])]
    S.finalize()
    qa = lambda a, ax: (*(math.sin(a / 2) * np.asarray(ax, dtype=float) / np.linalg.norm(ax)), math.cos(a / 2))
    times = [0., .01, .02, .03]
    poss = [c, c + (.02, 0, 0), c + (.02, .03, 0), c + (.02, .03, -.01)]
    oris = [qa(0., (1, 0, 0)), qa(.5, (1, 0, 0)), qa(1.2, (1, 1, 0)), qa(1.2, (1, 1, 0))]
    sc.attach_impose(S, [sc.imp_interpolated_motion(times, poss, oris)], groups=[0])
    O = Oracle(S)
    rel = np.array(S.facets[0]["v"]) - c

    def rot(q, v):
        x, y, z, w = q; u = np.array([x, y, z]); return v + 2 * w * np.cross(u, v) + 2 * np.cross(u, np.cross(u, v))
    done = 0
    for k in (1, 2, 3):
        n = int(round(times[k] / dt)) + 1 - done; O.step(n); done += n      # one step after times[k]
        f = O.facets()[0]
        want = np.array([poss[k] + rot(oris[k], r) for r in rel])
        assert np.allclose(f, want, rtol=0, atol=1e-12), (k, np.abs(f - want).max())
    O.step(50)                                                              # past the last time point: the pose stays
    assert np.allclose(O.facets()[0], want, rtol=0, atol=1e-12)


def test_facet_contact_through_the_periodic_boundary_equals_the_unwrapped_one():
    """A sphere that touches a facet through the cell boundary (Contact.cellDist != 0; Bo1_Facet_Aabb's periodic branch Facet.cpp:286-297, Cg2_Facet_Sphere with shift2)
    feels exactly the force of the same sphere placed one period away, next to the facet, in an aperiodic scene."""
    L = np.array([.5, .4, .45]); r = .03
    tri = np.array([(.02, .05, .2), (.22, .05, .2), (.02, .3, .2)])

    def scene(periodic):
        S = _free_spheres(1, r=r)
        S.pos[0] = (.08, .12, .2 + .9 * r) if not periodic else (.08 + L[0], .12 - L[1], .2 + .9 * r)      # one period away in x and y
        S.vel[0] = (.1, 0., -.2)
        S.facets = [dict(v=tri, halfThick=0., fakeVel=(0., 0., 0.), mat=0, mask=sc.MASK_BOUNDARY)]
        if periodic:
            S.params.periodic = 1
            S.cell = dict(hSize=np.diag(L), gradV=np.zeros((3, 3)), nextGradV=np.zeros((3, 3)), homoDeform=sc.HOMO_GRADV2)
        return S.finalize()
    A, B = Oracle(scene(False)), Oracle(scene(True))
    for k in range(5):
        A.step(3); B.step(3)                                    # the sphere rebounds within ~20 steps
        ca, cb = A.contacts(True), B.contacts(True)
        assert len(ca["ids"]) == len(cb["ids"]) == 1
        assert np.allclose(ca["force"], cb["force"], rtol=1e-12, atol=0) and ca["uN"][0] == pytest.approx(cb["uN"][0], rel=1e-10)
        assert np.allclose(A.spheres()["force"], B.spheres()["force"], rtol=1e-10, atol=1e-18)
    assert tuple(np.abs(cb["cellDist"][0])) == (1, 1, 0) and ca["force"][0][0] != 0.
