"""Impose on the CUDA path (wdem_upload_imposes) against the oracle's restatement of the hooks in Leapfrog (Leapfrog.cpp:203-205,238,244,248) and of
Impose.hpp / Impose.cpp.  Time-dependent factors are evaluated by the host with libm on both sides: everything but the read-force sums (atomics) is bit-identical."""
import math
import numpy as np
import pytest
from woo_b200 import scenes as sc
from _harness import Oracle

pytestmark = pytest.mark.gpu


def engine_for(S):
    from woo_b200.engine import Engine
    return Engine(S.params).load(S)


def check(O, E, step):
    so, sg = O.spheres(), E.spheres()
    co, cg = O.contacts(True), E.contacts(True)
    assert [tuple(x) for x in co["ids"]] == [tuple(x) for x in cg["ids"]], f"real contact set differs at step {step}"
    assert O.nFullRuns == E.scalars()["nFullRuns"] and O.dt == E.scalars()["dt"]
    for k in ("pos", "vel", "angVel", "force", "torque"):
        assert np.array_equal(so[k], sg[k]), f"sphere {k} not bit-identical at step {step} (max diff {np.abs(so[k] - sg[k]).max()})"
    for k in ("uN", "force", "contPt", "trsf", "phys"):
        assert np.array_equal(co[k], cg[k]), f"contact {k} not bit-identical at step {step}"
    if len(O.spec.walls):
        assert np.array_equal(O.walls()["pos"], E.walls()["pos"])
    if len(O.spec.facets):
        assert np.array_equal(O.facets(), E.facets())


def _gas(n=300, seed=5, law=sc.LAW_IDEALELPL, gravity=(0, 0, -9.81)):
    S = sc.random_gas(n, box=.5, r=.02, psd_ratio=1.5, seed=seed, law=law)
    S.params.gravity = gravity
    return S


@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
def test_every_sphere_imposition_kind_bitwise(law):
    S = _gas(law=law)
    n = len(S.pos); q = np.array([.1, -.2, .3, .9]); q /= np.linalg.norm(q); nan = float("nan")
    sc.attach_impose(S, [sc.imp_harmonic(20., .01, dir=(1, 1, 0))], spheres=range(0, 10))
    sc.attach_impose(S, [sc.imp_harmonic(15., .02, dir=(0, 1, 2), t0=.003, perpFree=True)], spheres=range(10, 20))
    sc.attach_impose(S, [sc.imp_aligned_harmonic((30., nan, 10.), (.005, 1., .01))], spheres=range(20, 30))
    sc.attach_impose(S, [sc.imp_constant_force((0, .02, .05))], spheres=range(30, 50))
    sc.attach_impose(S, [sc.imp_radial_force((.25, .25, 0), (.25, .25, 1), -.03)], spheres=range(50, 70))
    sc.attach_impose(S, [sc.imp_local6((.1, 0, 0, 0, 0, 2.), (1, 0, 0, 0, 0, 1), ori=q)], spheres=range(70, 80))
    sc.attach_impose(S, [sc.imp_local6((0, .01, 0, 0, 1e-5, 0), (0, 2, 0, 0, 2, 0), ori=q), sc.imp_constant_force((.01, 0, 0))], spheres=range(80, 90))      # CombinedImpose
    sc.attach_impose(S, [sc.imp_velocity_and_read_force((0, 0, 1), -.2, latBlock=False)], spheres=range(90, 100))
    sc.attach_impose(S, [sc.imp_variable_velocity3d([0, .002, .004, .006], [(0, 0, 0), (.3, nan, 0), (-.3, nan, .1), (0, nan, 0)], wrap=True, ori=q)], spheres=range(100, 110))
    sc.attach_impose(S, [sc.imp_variable_velocity3d([0, .01], [(.1, 0, 0), (.1, 0, 0)], angVels=[(0, 0, 5.), (0, 8., nan)], diff=True, t0=.001)], spheres=range(110, 115))
    sc.attach_impose(S, [sc.imp_variable_aligned_rotation(2, [(0, 0), (.005, 50.), (.01, 0)], wrap=True), sc.imp_read_force()], spheres=range(115, 125))
    sc.attach_impose(S, [sc.imp_read_force(node_pos=(.2, .2, .2), node_ori=q)], spheres=range(125, n))
    O, E = Oracle(S), engine_for(S)
    for k in range(8):
        O.step(150); E.step(150)
        check(O, E, 150 * (k + 1))
        for ix, d in enumerate(S.imposes):
            so, sg = O.impose_state(ix), E.impose_state(ix)
            assert so["dist"] == sg["dist"]
            scale = max(abs(so["sumF"]), np.abs(so["F"]).max(), np.abs(so["T"]).max(), 1e-300)
            assert abs(so["sumF"] - sg["sumF"]) <= 1e-10 * scale and np.allclose(so["F"], sg["F"], rtol=0, atol=1e-10 * scale) and np.allclose(so["T"], sg["T"], rtol=0, atol=1e-10 * scale)
    assert len(O.contacts(True)["ids"]) > 20 and O.nFullRuns > 3


def test_imposition_on_walls_piston_and_shaker():
    """A wall driven through a velocity table (piston) and one shaken harmonically: fully blocked nodes, moved by the velocity imposed at the end of the previous step."""
    S = sc.deposition(400, r=5e-3, psd_ratio=1.5, seed=3)
    S.params.dynDtPeriod = 50
    nan = float("nan")
    zwall = [i for i, w in enumerate(S.walls) if w["axis"] == 2][0]
    xwall = [i for i, w in enumerate(S.walls) if w["axis"] == 0][0]
    sc.attach_impose(S, [sc.imp_variable_velocity3d([0, .01, .03], [(nan, nan, 0), (nan, nan, .05), (nan, nan, -.02)])], walls=[zwall])
    sc.attach_impose(S, [sc.imp_harmonic(40., 2e-4, dir=(1, 0, 0))], walls=[xwall])
    O, E = Oracle(S), engine_for(S)
    for k in range(10):
        O.step(200); E.step(200)
        check(O, E, 200 * (k + 1))
    assert O.walls()["pos"][zwall, 2] != S.walls[zwall]["pos"][2] and O.nFullRuns > 5


def test_imposition_on_a_facet_group_vibrating_mesh():
    S = sc.drum(200, div=10, axial=2, vz0=.3, vfluct=.2, omega=0., rotate=True)
    S.facet_groups[0]["angVel"] = (0, 0, 0)
    sc.attach_impose(S, [sc.imp_aligned_harmonic((25., float("nan"), 25.), (1e-3, 0, 2e-3)), sc.imp_variable_aligned_rotation(1, [(0, 0), (.02, 3.)])], groups=[0])
    O, E = Oracle(S), engine_for(S)
    for k in range(6):
        O.step(150); E.step(150)
        check(O, E, 150 * (k + 1))
    assert not np.array_equal(E.facets(), np.array([f["v"] for f in S.facets]).reshape(-1, 3, 3))


def test_interpolated_motion_of_a_facet_group_bitwise():
    """InterpolatedMotion (Impose.cpp:111-133) on the master node of a kinematic facet group: the host evaluates it with libm, as the oracle does — bit-identical."""
    S = sc.drum(200, div=10, axial=2, vz0=.3, vfluct=.2, omega=0., rotate=True)
    S.facet_groups[0]["angVel"] = (0, 0, 0)
    c = np.array(S.facet_groups[0]["pos"], dtype=float)
    qa = lambda a, ax: (*(math.sin(a / 2) * np.asarray(ax, dtype=float) / np.linalg.norm(ax)), math.cos(a / 2))
    times = [0., .004, .008, .02]
    poss = [c, c + (1e-3, 0, 0), c + (1e-3, 0, 2e-3), c]
    oris = [qa(0., (1, 0, 0)), qa(.05, (1, 0, 0)), qa(-.03, (1, 0, .2)), qa(.2, (1, 0, 0))]
    sc.attach_impose(S, [sc.imp_interpolated_motion(times, poss, oris, t0=5e-4)], groups=[0])
    O, E = Oracle(S), engine_for(S)
    for k in range(6):
        O.step(150); E.step(150)
        check(O, E, 150 * (k + 1))
    assert not np.array_equal(E.facets(), np.array([f["v"] for f in S.facets]).reshape(-1, 3, 3))
    S2 = _gas(50)
    sc.attach_impose(S2, [sc.imp_interpolated_motion([0, 1], [(0, 0, 0), (1, 0, 0)], [(0, 0, 0, 1), (0, 0, 0, 1)])], spheres=[0])
    with pytest.raises(RuntimeError, match="facet groups only"):
        engine_for(S2)


def test_orbit_impositions_on_spheres_and_on_the_nodes_of_a_rotating_cylinder():
    """CircularOrbit / StableCircularOrbit (Impose.cpp:22-48) take atan2 / sin / cos of the node position on the device: nodes that carry them follow the oracle to
    rounding, and with them — for a while, contacts amplify rounding — the spheres they touch.  A cylinder whose every node carries a StableCircularOrbit is what
    woo.triangulated.cylinder(angVel=..., realRot=True) builds (py/triangulated.py:59-64)."""
    om = 5.
    S = sc.drum(200, div=10, axial=2, vz0=.3, vfluct=.2, omega=om, rotate=False)
    R = float(np.hypot(S.facets[0]["v"][0][1], S.facets[0]["v"][0][2]))
    for f in S.facets: f["fakeVel"] = (0., 0., 0.)
    q = (0., math.sin(math.pi / 4), 0., math.cos(math.pi / 4))      # local z -> the drum's axis x
    sc.attach_impose(S, [sc.imp_stable_circular_orbit((0., 0., 0.), om, R, node_ori=q)], facets=range(len(S.facets)))
    sc.attach_impose(S, [sc.imp_circular_orbit((0., 0., 0.), 2., node_ori=q, rotate=True)], spheres=range(0, 5))
    sc.attach_impose(S, [sc.imp_stable_circular_orbit((0., 0., 0.), -3., .4 * R, node_ori=q)], spheres=range(5, 10))
    O, E = Oracle(S), engine_for(S)
    v0 = E.facets().copy()
    for k in range(4):
        O.step(100); E.step(100)
        fo, fg = O.facets(), E.facets()
        assert np.allclose(fo, fg, rtol=0, atol=1e-12 * R), f"facet nodes differ at step {100 * (k + 1)}: {np.abs(fo - fg).max()}"
        so, sg = O.spheres(), E.spheres()
        for key in ("pos", "vel"):
            # the spheres that carry an orbit are kinematic: rounding only, at every check; the free ones touch them and the cylinder, and contacts amplify rounding
            # exponentially (2e-7 after 300 steps): tight at the first check, loose afterwards
            assert np.allclose(so[key][:10], sg[key][:10], rtol=0, atol=1e-11 * max(1., np.abs(so[key]).max())), f"orbiting sphere {key} differs at step {100 * (k + 1)}: {np.abs(so[key][:10] - sg[key][:10]).max()}"
            assert np.allclose(so[key], sg[key], rtol=0, atol=(1e-8 if k == 0 else 1e-3) * max(1., np.abs(so[key]).max())), f"sphere {key} differs at step {100 * (k + 1)}: {np.abs(so[key] - sg[key]).max()}"
        assert O.nFullRuns == E.scalars()["nFullRuns"] and O.dt == E.scalars()["dt"]
        if k == 0:
            assert [tuple(x) for x in O.contacts(True)["ids"]] == [tuple(x) for x in E.contacts(True)["ids"]]
        assert O.impose_state(0)["dist"] == E.impose_state(0)["dist"]      # .angle
    rr = np.hypot(fg[..., 1], fg[..., 2])
    assert np.allclose(rr, R, rtol=1e-12) and not np.allclose(fg, v0, atol=1e-3 * R), "the cylinder keeps its radius and has turned"
    assert len(E.contacts(True)["ids"]) > 0 and E.scalars()["nFullRuns"] > 5


def test_imposition_on_a_facet_group_in_a_periodic_cell():
    """A plate that follows a recorded trajectory (InterpolatedMotion on the master node of its clump) and is shaken (AlignedHarmonicOscillations on a second one) inside
    a sheared periodic cell: host-evaluated, bit-identical."""
    S = sc.periodic_gas(300, cell=(.5, .4, .45), r=.02, psd_ratio=2.0, seed=71, gradV=[[0, 1.2, 0], [0, 0, 0], [0, 0, 0]], shear0=.1)
    quad = lambda x0, y0, x1, y1, z: [np.array([(x0, y0, z), (x1, y0, z), (x1, y1, z)]), np.array([(x0, y0, z), (x1, y1, z), (x0, y1, z)])]
    tris = quad(.1, .05, .25, .2, .15) + quad(.3, .2, .45, .35, .3)
    S.facets = [dict(v=t, halfThick=0., fakeVel=(0., 0., 0.), mat=0, mask=sc.MASK_BOUNDARY) for t in tris]
    S.facet_groups = [dict(pos=(.175, .125, .15), ori=(0., 0., 0., 1.), vel=(0., 0., 0.), angVel=(0., 0., 0.), facets=[0, 1]),
                      dict(pos=(.375, .275, .3), ori=(0., 0., 0., 1.), vel=(0., 0., 0.), angVel=(0., 0., 0.), facets=[2, 3])]
    S.finalize()
    qa = lambda a, ax: (*(math.sin(a / 2) * np.asarray(ax, dtype=float) / np.linalg.norm(ax)), math.cos(a / 2))
    c0 = np.array(S.facet_groups[0]["pos"])
    sc.attach_impose(S, [sc.imp_interpolated_motion([0., .002, .004], [c0, c0 + (.01, 0, .02), c0 + (.01, .02, 0)], [qa(0, (0, 0, 1)), qa(.3, (0, 0, 1)), qa(.3, (1, 0, 1))])], groups=[0])
    sc.attach_impose(S, [sc.imp_aligned_harmonic((float("nan"), 200., 150.), (0, 2e-3, 3e-3))], groups=[1])
    O, E = Oracle(S), engine_for(S)
    for k in range(6):
        O.step(25); E.step(25)
        check(O, E, 25 * (k + 1))
    assert not np.array_equal(E.facets(), np.array(tris)) and (O.contacts(True)["ids"][:, 0] >= S.n).sum() > 0


def test_axially_periodic_rotating_drum():
    """A drum without end caps in a cell that is periodic along its axis: the cylinder of woo.triangulated.cylinder(realRot=True) (one StableCircularOrbit on every node,
    py/triangulated.py:59-64) plus Bo1_Facet_Aabb's periodic branch; spheres leave through one end and come back through the other."""
    om = 5.
    S = sc.drum(200, div=10, axial=3, vz0=.3, vfluct=.2, omega=om, rotate=False)
    R = float(np.hypot(S.facets[0]["v"][0][1], S.facets[0]["v"][0][2])); L = float(max(w["pos"][0] for w in S.walls))
    for f in S.facets: f["fakeVel"] = (0., 0., 0.)
    S.walls = []
    S.vel[:, 0] += 1.0                                                 # drift along the axis: through the periodic boundary
    S.params.periodic = 1
    S.cell = dict(hSize=np.diag([L, 3 * R, 3 * R]), gradV=np.zeros((3, 3)), nextGradV=np.zeros((3, 3)), homoDeform=sc.HOMO_GRADV2)
    q = (0., math.sin(math.pi / 4), 0., math.cos(math.pi / 4))
    sc.attach_impose(S, [sc.imp_stable_circular_orbit((0., 0., 0.), om, R, node_ori=q)], facets=range(len(S.facets)))
    S.finalize()
    O, E = Oracle(S), engine_for(S)
    seen = 0
    for k in range(4):
        O.step(100); E.step(100)
        fo, fg = O.facets(), E.facets()
        assert np.allclose(fo, fg, rtol=0, atol=1e-12 * R), f"facet nodes differ at step {100 * (k + 1)}: {np.abs(fo - fg).max()}"
        so, sg = O.spheres(), E.spheres()
        for key in ("pos", "vel"):
            assert np.allclose(so[key], sg[key], rtol=0, atol=(1e-8 if k == 0 else 1e-3) * max(1., np.abs(so[key]).max())), f"sphere {key} differs at step {100 * (k + 1)}: {np.abs(so[key] - sg[key]).max()}"
        assert O.nFullRuns == E.scalars()["nFullRuns"] and O.dt == E.scalars()["dt"]
        co = E.contacts(True)
        if len(co["ids"]): seen = max(seen, int(np.abs(co["cellDist"]).max()))
    assert (E.spheres()["pos"][:, 0] > L).any(), "spheres should have left the cell along the axis"
    assert np.allclose(np.hypot(fg[..., 1], fg[..., 2]), R, rtol=1e-12)


def test_impose_errors():
    from woo_b200.engine import Engine
    S = _gas(50)
    sc.attach_impose(S, [dict(sc.imp_constant_force((0, 0, 1)), kind=42)], spheres=[0])
    with pytest.raises(RuntimeError, match="unknown kind"):
        Engine(S.params).load(S)
    S = _gas(50); S.walls = sc.box_walls((0, 0, 0), (.5, .5, .5)); S.finalize()
    sc.attach_impose(S, [sc.imp_constant_force((0, 0, 1))], walls=[0])
    with pytest.raises(RuntimeError, match="velocity impositions only"):
        Engine(S.params).load(S)
    S = _gas(50); S.walls = sc.box_walls((0, 0, 0), (.5, .5, .5)); S.finalize()
    sc.attach_impose(S, [sc.imp_stable_circular_orbit((0, 0, 0), 1., .1)], walls=[0])
    with pytest.raises(RuntimeError, match="not on walls"):
        Engine(S.params).load(S)
