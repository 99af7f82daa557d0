"""Clumps of spheres with mass on the CUDA path (wdem_upload_clumps) against the oracle's restatement of Leapfrog's clump branch
(Leapfrog.cpp:148-255,299-331), ClumpData::forceTorqueFromMembers / applyToMembers (Clump.cpp:272-307), Collider::mayCollide (Collision.cpp:34-39) and
DynDt's clump branch (DynDt.cpp:48-52)."""
import numpy as np
import pytest
from woo_b200 import scenes as sc
from _harness import Oracle

pytestmark = pytest.mark.gpu


def engine_for(S):
    from woo_b200.engine import Engine
    return Engine(S.params).load(S)


def check(O, E, step, exact=True):
    so, sg = O.spheres(), E.spheres()
    co, cg = O.contacts(True), E.contacts(True)
    ko, kg = O.clumps(), E.clumps()
    assert [tuple(x) for x in co["ids"]] == [tuple(x) for x in cg["ids"]], f"real contact set differs at step {step}"
    assert O.nFullRuns == E.scalars()["nFullRuns"], f"collider full runs differ at step {step}"
    assert O.dt == E.scalars()["dt"], f"DynDt differs at step {step}"
    if exact:
        for k in ("pos", "vel", "angVel", "force", "torque"):
            assert np.array_equal(so[k], sg[k]), f"sphere {k} not bit-identical at step {step} (max diff {np.abs(so[k] - sg[k]).max()})"
        for k in ("pos", "ori", "vel", "angVel"):
            assert np.array_equal(ko[k], kg[k]), f"clump {k} not bit-identical at step {step} (max diff {np.abs(ko[k] - kg[k]).max()})"
        for k in ("uN", "force", "contPt", "trsf", "phys"):
            assert np.array_equal(co[k], cg[k]), f"contact {k} not bit-identical at step {step}"
    else:
        scale = max(np.abs(co["force"]).max() if len(co["force"]) else 0., 1e-30)
        assert np.allclose(co["force"], cg["force"], rtol=1e-10, atol=1e-10 * scale)
        assert np.allclose(so["force"], sg["force"], rtol=1e-10, atol=1e-10 * scale)
        assert np.allclose(ko["pos"], kg["pos"], rtol=1e-10, atol=1e-12)
    # orientation of a MEMBER sphere of an aspherical clump follows the master without transcendental functions: exact; free spheres: sin/cos tolerance
    assert np.allclose(so["ori"], sg["ori"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
def test_clump_deposition_bitwise(law):
    """Aspherical clumps (the usual case: principal inertias differ) use no transcendental function: every number, DynDt's time step included, is bit-identical."""
    S = sc.clump_deposition(60, 40, law=law)
    O, E = Oracle(S), engine_for(S)
    per = 500 if law == sc.LAW_IDEALELPL else 1500       # until the aggregates lie in a bed on the floor
    for k in range(16):
        O.step(per); E.step(per)
        check(O, E, per * (k + 1))
    assert O.nFullRuns > 100 and len(O.contacts(True)["ids"]) > 100


def test_impositions_on_clump_master_nodes_bitwise():
    """DemData.impose of a clump's master node (Leapfrog.cpp:203-205,244,248 on the clump node, before ClumpData::applyToMembers): a stirrer clump driven harmonically,
    one pushed by a constant force, one whose rotation about z is prescribed by a table (the members follow rigidly), among free clumps and spheres."""
    S = sc.clump_deposition(30, 40)
    nan = float("nan")
    sc.attach_impose(S, [sc.imp_aligned_harmonic((30., nan, nan), (2e-3, 0, 0))], clumps=[0, 1])
    sc.attach_impose(S, [sc.imp_constant_force((0, 3e-3, 0)), sc.imp_variable_aligned_rotation(2, [(0, 0), (.01, 20.), (.05, 20.)])], clumps=[2])
    sc.attach_impose(S, [sc.imp_local6((0, 0, 0, 0, 0, 0), (1, 1, 0, 0, 0, 0))], clumps=[3])      # no lateral motion: falls straight down
    sc.attach_impose(S, [sc.imp_harmonic(20., 1e-3, dir=(0, 1, 0))], spheres=[len(S.pos) - 1])
    O, E = Oracle(S), engine_for(S)
    x3 = E.clumps()["pos"][3].copy()
    for k in range(16):
        O.step(500); E.step(500)
        check(O, E, 500 * (k + 1))
    c = E.clumps()
    assert c["pos"][3][0] == x3[0] and c["pos"][3][1] == x3[1] and c["pos"][3][2] < x3[2]
    assert O.nFullRuns > 50 and len(O.contacts(True)["ids"]) > 30
    from woo_b200.engine import Engine
    S2 = sc.clump_deposition(5, 0)
    sc.attach_impose(S2, [sc.imp_read_force()], clumps=[0])
    with pytest.raises(RuntimeError, match="read-force kinds"):
        Engine(S2.params).load(S2)


@pytest.mark.parametrize("homo", [sc.HOMO_POS, sc.HOMO_VEL, sc.HOMO_VEL_2ND, sc.HOMO_GRADV2])
def test_clumps_in_a_periodic_cell_bitwise(homo):
    """Rigid aggregates and free spheres in a periodic cell that is compressed and sheared: the clump's master node gets the cell's corrections like any node
    (fluctuation velocity in the damping, the gradV2 velocity update, applyPeriodicCorrections; Leapfrog.cpp:181-215,258-273), members of one clump never collide,
    contacts reach across the cell boundary (cellDist)."""
    S = sc.clump_deposition(40, 60, damping=.1, spin=20.)
    hi = np.array([max(w["pos"][k] for w in S.walls) for k in range(3)], dtype=float); hi[2] = S.pos[:, 2].max() + 3 * S.radius.max()
    S.walls = []
    S.params.periodic = 1; S.params.gravity = (0., 0., 0.); S.params.dynDtPeriod = 20
    g = np.array([[-12., 6., 0], [0, -12., 0], [0, 0, -12.]])
    S.cell = dict(hSize=np.diag(hi * 1.01), gradV=g, nextGradV=g, homoDeform=homo)
    S.vel = S.vel + S.pos @ g.T                                       # the mean field: the particles follow the cell
    for c in S.clumps: c["vel"] = tuple(np.asarray(c["vel"]) + g @ np.asarray(c["pos"]))
    S.finalize()
    O, E = Oracle(S), engine_for(S)
    for k in range(10):
        O.step(250); E.step(250)
        check(O, E, 250 * (k + 1))
        assert np.array_equal(O.cell_hsize(), E.cell_hsize())
    co = O.contacts(True)
    assert len(co["ids"]) > 40 and np.abs(co["cellDist"]).max() > 0, "the test should see contacts across the cell boundary"


def test_clump_first_steps_and_member_kinematics():
    S = sc.clump_deposition(9, 0, dynDtPeriod=0)
    O, E = Oracle(S), engine_for(S)
    for k in range(5):
        O.step(1); E.step(1)
        check(O, E, k + 1)
    s, c = E.spheres(), E.clumps()
    for ci, cl in enumerate(S.clumps):
        for m in cl["members"]:
            assert np.allclose(s["vel"][m], c["vel"][ci] + np.cross(c["angVel"][ci], s["pos"][m] - c["pos"][ci]), rtol=1e-12, atol=1e-14)     # applyToMembers
            assert np.array_equal(s["angVel"][m], c["angVel"][ci])
    assert np.all(np.isfinite(c["angMom"]))                                                        # initialised by the first step
    g = np.array(S.params.gravity)
    assert np.allclose(c["force"], np.array([cl["mass"] for cl in S.clumps])[:, None] * g, rtol=1e-15)     # reset: gravity sits on the master node


def test_members_of_one_clump_never_collide():
    S = sc.clump_deposition(6, 0, dynDtPeriod=0)
    E = engine_for(S)
    E.step(3)
    c = E.contacts(False)
    of = {}
    for ci, cl in enumerate(S.clumps):
        for m in cl["members"]:
            of[int(S.sphere_ids()[m])] = ci
    for a, b in c["ids"]:
        assert not (a in of and b in of and of[a] == of[b])
    # without the clumps the overlapping members do collide
    S2 = sc.clump_deposition(6, 0, dynDtPeriod=0); S2.clumps = []; S2.flags[:] = 0
    E2 = engine_for(S2); E2.step(1)
    assert len(E2.contacts(True)["ids"]) > 0


def test_blocked_and_spherical_clumps():
    """Blocked DOFs on the master node; a clump with equal principal inertias (a 1-clump) takes the spherical integrator (sin/cos: tolerance)."""
    S = sc.clump_deposition(12, 10, dynDtPeriod=0)
    sc.make_clumps(S, [[len(S.pos) - 1], [len(S.pos) - 2]], vel=[(0, 0, -.1), (.1, 0, 0)], angVel=[(3., 1., -2.), (0, 0, 0)])     # 1-clumps: equal principal inertias
    S.clumps[0]["flags"] = 63                                   # fully blocked: the aggregate keeps moving with its prescribed velocity, its spin turns it
    S.clumps[1]["flags"] = 56                                   # rotations blocked
    S.clumps[2]["flags"] = 4 | 8                                # z and rot-x blocked (aspherical path: torque component zeroed)
    S.clumps[3]["flags"] = 2048 | 512                           # DAMPING_SKIP | GRAVITY_SKIP
    O, E = Oracle(S), engine_for(S)
    for k in range(8):
        O.step(250); E.step(250)
        check(O, E, 250 * (k + 1), exact=False)
    kg = E.clumps()
    assert np.array_equal(kg["vel"][0], S.clumps[0]["vel"]) and np.array_equal(kg["angVel"][0], S.clumps[0]["angVel"])      # blocked DOFs keep their velocity
    assert np.array_equal(kg["angVel"][1], S.clumps[1]["angVel"]) and kg["vel"][2, 2] == S.clumps[2]["vel"][2]
    assert np.allclose(kg["pos"][0], S.clumps[0]["pos"] + E.scalars()["time"] * S.clumps[0]["vel"], rtol=1e-9)


def test_clump_errors():
    from woo_b200.engine import Engine
    S = sc.clump_deposition(4, 2)
    cl = S.clumps; S.clumps = []
    E = Engine(S.params).load(S)
    with pytest.raises(RuntimeError, match="CLUMPED flag but no clump"):
        E.step(1)
    S.clumps = cl[:-1]
    with pytest.raises(RuntimeError, match="are members of the uploaded clumps"):
        Engine(S.params).load(S)
    S.clumps = cl
    E = Engine(S.params).load(S); E.step(2)
    with pytest.raises(RuntimeError, match="scenes with clumps"):
        E.remove_spheres([int(S.sphere_ids()[-1])])
    St = sc.clump_deposition(4, 2)
    Et = Engine(St.params).load(St); Et.track_energy(True)
    with pytest.raises(RuntimeError, match="trackEnergy"):
        Et.step(1)
