"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on seeded scenes."""
import numpy as np
import pytest
from woo_b200 import scenes as sc
from _harness import Oracle

pytestmark = pytest.mark.gpu


def engine_for(S):
    from woo_b200.engine import Engine
    return Engine(S.params).load(S)


def compare(S, nsteps, check_every=1, exact=True, label="", make=None):
    O = Oracle(S); E = (make or engine_for)(S)
    done = 0
    while done < nsteps:
        k = min(check_every, nsteps - done)
        O.step(k); E.step(k); done += k
        so, sg = O.spheres(), E.spheres()
        co, cg = O.contacts(True), E.contacts(True)
        # bit-exact: the set of real contacts
        assert [tuple(x) for x in co["ids"]] == [tuple(x) for x in cg["ids"]], f"{label}: real contact set differs at step {done}"
        assert O.nFullRuns == E.scalars()["nFullRuns"], f"{label}: collider full runs differ at step {done}"
        assert O.dt == E.scalars()["dt"]
        if exact:
            for k_ in ("pos", "vel", "angVel", "force", "torque"):
                assert np.array_equal(so[k_], sg[k_]), f"{label}: {k_} not bit-identical at step {done} (max diff {np.abs(so[k_]-sg[k_]).max()})"
            for k_ in ("uN", "force", "contPt", "trsf", "phys"):
                assert np.array_equal(co[k_], cg[k_]), f"{label}: contact {k_} not bit-identical at step {done}"
        else:
            scale = np.abs(co["force"]).max() if len(co["force"]) else 1.
            assert np.allclose(co["force"], cg["force"], rtol=1e-10, atol=1e-10 * scale)
            assert np.allclose(so["force"], sg["force"], rtol=1e-10, atol=1e-10 * scale)
        # orientation: tolerance only (sin/cos differ by ulps between glibc and CUDA)
        assert np.allclose(so["ori"], sg["ori"], rtol=0, atol=1e-12)
    return O, E


@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
def test_two_spheres_head_on(law):
    S = sc.SceneSpec()
    S.mats = [sc.frict_mat()] if law == sc.LAW_IDEALELPL else [sc.hertz_mat()]
    S.pos = np.array([[0, 0, 0], [0, .10001, 0.001]]); S.radius = np.array([.05, .05])
    S.vel = np.array([[0, .5, 0], [0, -.5, 0.1]]); S.angVel = np.array([[1., 2, 3], [-3, 2, 1]])
    S.params = sc.Params(law=law, damping=0., dtSafety=.4, dynDtPeriod=10, en=.5 if law else float("nan"))
    compare(S.finalize(), 600, check_every=50, label="two-spheres")


@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
@pytest.mark.parametrize("psd", [1.0, 4.0])
def test_random_gas_bitwise(law, psd):
    S = sc.random_gas(400, box=.5, r=.02, psd_ratio=psd, seed=3, law=law, en=.7 if law else float("nan"))
    compare(S, 300, check_every=25, label=f"gas law={law} psd={psd}")


def test_deposition_cfg1_small():
    S = sc.deposition(2000, seed=1)
    compare(S, 400, check_every=100, label="deposition-2k")


def test_blocked_dofs_and_skips():
    S = sc.random_gas(200, box=.4, r=.02, seed=5)
    S.flags[::7] = sc.DOF_ALL; S.flags[1::7] = 1 | 8; S.flags[2::7] = sc.GRAVITY_SKIP; S.flags[3::7] = sc.DAMPING_SKIP
    compare(S, 200, check_every=50, label="flags")


def test_law_flags():
    for kw in (dict(noSlip=1), dict(noFrict=1), dict(noRatch=1), dict(trsfRenorm=7), dict(damping=0.), dict(relRollStiff=.3, relTwistStiff=.2, rollTanPhi=.4), dict(applyForces=0)):
        S = sc.random_gas(200, box=.4, r=.02, seed=7)
        for k, v in kw.items():
            setattr(S.params, k, v)
        compare(S, 120, check_every=40, label=str(kw))


def test_verlet_zero_and_fixed():
    for vd in (0., .004):
        S = sc.random_gas(150, box=.35, r=.02, seed=11); S.params.verletDist = vd
        compare(S, 60, check_every=20, label=f"verletDist={vd}")


@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
def test_facet_box(law):
    S = sc.facet_box_scene(300, seed=2, law=law, en=.6 if law else float("nan"))
    compare(S, 200, check_every=25, label=f"facet-box law={law}")


def test_facet_variants():
    for kw in (dict(halfThick=.004), dict(fakeVel=(.3, -.2, .1)), dict(fakeVel=(float("nan"), 0, 0)), dict(random_n21lim=False, div=1)):
        S = sc.facet_box_scene(200, seed=4, **kw)
        compare(S, 120, check_every=30, label=f"facets {kw}")


def test_mixed_walls_facets_ids_first():
    S = sc.facet_box_scene(250, seed=9, div=2)
    S.walls = sc.box_walls((-.05, -.05, -.05), (.45, .45, .45), mat=0, top=True)
    S.boundaries_first = True
    compare(S, 150, check_every=50, label="mixed")


# ---------------------------------------------------------------------------------------------------------------
# periodic cell (B4/I2: spatialOverlapPeri, cellDist, Cell.integrateAndUpdate, GRADV2 Leapfrog)
# ---------------------------------------------------------------------------------------------------------------
def compare_potential(O, E):
    co, cg = O.contacts(False), E.contacts(False)
    assert [tuple(x) for x in co["ids"]] == [tuple(x) for x in cg["ids"]], "potential contact set differs"
    assert np.array_equal(co["cellDist"], cg["cellDist"]), "cellDist differs"


@pytest.mark.parametrize("shift", [(0, 0, 0), (1, 2, 3), (6, 8, -4)])
def test_pbc_celldist_kat(shift):
    # py/tests/test_pbc.py:157-171 on the GPU: C.cellDist == -shift
    S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
    S.pos = np.array([(.5, .5, .5), (.8 + shift[0], .6 + shift[1], .5 + shift[2])]); S.radius = np.array([.2, .2])
    S.params = sc.Params(dt=1e-9, periodic=1, damping=0.)
    S.cell = dict(hSize=np.eye(3), homoDeform=sc.HOMO_GRADV2)
    E = engine_for(S.finalize()); E.step(1)
    c = E.contacts(False)
    assert [tuple(x) for x in c["ids"]] == [(0, 1)] and tuple(c["cellDist"][0]) == tuple(-np.array(shift))


def test_pbc_ambiguous_raises():
    # py/tests/test_pbc.py:145-154
    from woo_b200.engine import WdemError
    S = sc.SceneSpec(); S.mats = [sc.frict_mat()]
    S.pos = np.array([(.3, .5, .5), (.8, .9, .5)]); S.radius = np.array([.4, .2])
    S.params = sc.Params(dt=1e-9, periodic=1, damping=0.)
    S.cell = dict(hSize=np.eye(3), homoDeform=sc.HOMO_GRADV2)
    E = engine_for(S.finalize())
    with pytest.raises(WdemError, match="ambiguous periodic contact"):
        E.step(1)


@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
@pytest.mark.parametrize("homo", [sc.HOMO_GRADV2, sc.HOMO_VEL_2ND, sc.HOMO_VEL, sc.HOMO_POS, sc.HOMO_NONE])
def test_periodic_shear_bitwise(law, homo):
    g = [[0, 2.0, 0], [0, 0, 0], [0, 0, 0]]
    S = sc.periodic_gas(400, cell=(.5, .4, .45), r=.02, psd_ratio=2.0, seed=13, law=law, gradV=g, homoDeform=homo, en=.7 if law else float("nan"))
    O, E = compare(S, 200, check_every=25, label=f"peri-shear law={law} homo={homo}")
    compare_potential(O, E)
    assert np.array_equal(np.asarray(O.cell_hsize()), np.asarray(E.cell_hsize()))
    assert O.kineticEnergy() == pytest.approx(E.kinetic_energy(), rel=1e-12)


def test_periodic_preskewed_compression():
    g = [[-1.0, .5, 0], [0, -1.0, 0], [.3, 0, -1.0]]
    S = sc.periodic_gas(300, cell=(.45, .45, .45), r=.02, psd_ratio=3.0, seed=17, gradV=g, shear0=.3)
    O, E = compare(S, 150, check_every=25, label="peri-skew-compress")
    compare_potential(O, E)


@pytest.mark.parametrize("homo", [sc.HOMO_POS, sc.HOMO_VEL, sc.HOMO_VEL_2ND, sc.HOMO_GRADV2])
def test_periodic_with_moving_walls_in_a_deforming_cell(homo):
    """Walls with a velocity are in DemField.nodes; in a periodic cell Leapfrog's fixed-node branch also applies the cell's velocity / position corrections to them
    (Leapfrog.cpp:206-215,258-273).  A piston closing on a gas in a laterally compressed (non-skewed) cell, every homoDeform mode."""
    S = sc.periodic_gas(300, cell=(.4, .4, .5), r=.02, psd_ratio=2.0, seed=23, walls=True, gradV=[[-.8, 0, 0], [0, -.5, 0], [0, 0, 0]], homoDeform=homo)
    S.walls[1]["vel"] = (0., 0., -1.5); S.walls[0]["vel"] = (0., 0., .4)
    S.finalize()
    O, E = compare(S, 200, check_every=25, label=f"peri moving walls homoDeform={homo}")
    compare_potential(O, E)
    wo, wg = O.walls(), E.walls()
    assert np.array_equal(wo["pos"], wg["pos"]) and wg["pos"][1, 2] < S.walls[1]["pos"][2] - .01
    assert (O.contacts(True)["ids"][:, 0] >= S.n).sum() > 3, "no wall contacts"


@pytest.mark.parametrize("shear0,world", [(0., 1), (.25, 1), (.25, 2)])
def test_periodic_with_static_facets(shear0, world):
    """Facets in a periodic cell (Bo1_Facet_Aabb's periodic branch, Facet.cpp:286-297: boxes of the unsheared node positions; Cg2_Facet_Sphere with shift2): a plate in
    the middle of a sheared cell; one of its triangles sticks out of the cell, so that spheres touch it through the boundary (cellDist != 0)."""
    S = sc.periodic_gas(400, cell=(.5, .4, .45), r=.02, psd_ratio=2.0, seed=61, gradV=[[0, 1.5, 0], [0, 0, 0], [0, 0, 0]], shear0=shear0)
    z = .2
    quad = lambda x0, y0, x1, y1: [np.array([(x0, y0, z), (x1, y0, z), (x1, y1, z)]), np.array([(x0, y0, z), (x1, y1, z), (x0, y1, z)])]
    tris = quad(.1, .05, .3, .3) + quad(.3, .05, .62, .3) + [np.array([(.05, .3, z), (.25, .55, z + .05), (-.1, .45, z - .03)])]      # the second quad and the last triangle leave the cell
    S.facets = [dict(v=t, halfThick=.004 * (k % 2), fakeVel=(0., 0., 0.), mat=0, mask=sc.MASK_BOUNDARY) for k, t in enumerate(tris)]
    S.finalize()
    O, E = compare(S, 100, check_every=25, label=f"peri facets shear0={shear0} world={world}", make=cluster(world) if world > 1 else None)
    compare_potential(O, E)
    co = O.contacts(True)
    fs = co["ids"][:, 0] >= S.n
    assert fs.sum() > 5 and np.abs(co["cellDist"][fs]).max() > 0, "the test should see facet contacts through the cell boundary"


@pytest.mark.parametrize("homo", [sc.HOMO_POS, sc.HOMO_VEL_2ND, sc.HOMO_GRADV2])
def test_periodic_with_a_moving_facet_group(homo):
    """A plate clumped to a master node that translates and rotates (ParticleContainer.addClumped with a central node) inside a periodic cell that is being sheared and
    compressed: the master node is a fully blocked node, so Leapfrog's fixed branch applies the cell's corrections to it (Leapfrog.cpp:206-215,258-273); the members
    follow, their unsheared boxes are refreshed at the full runs their motion triggers."""
    S = sc.periodic_gas(400, cell=(.5, .4, .45), r=.02, psd_ratio=2.0, seed=67, gradV=[[-.5, 1.5, 0], [0, -.3, 0], [0, 0, 0]], shear0=.1, homoDeform=homo)
    z = .2
    quad = lambda x0, y0, x1, y1: [np.array([(x0, y0, z), (x1, y0, z), (x1, y1, z)]), np.array([(x0, y0, z), (x1, y1, z), (x0, y1, z)])]
    tris = quad(.1, .05, .25, .2) + quad(.25, .05, .4, .2) + quad(.4, .05, .55, .2)      # the last one leaves the cell
    S.facets = [dict(v=t, halfThick=0., fakeVel=(0., 0., 0.), mat=0, mask=sc.MASK_BOUNDARY) for t in tris]
    S.facet_groups = [dict(pos=(.32, .12, z), ori=(0., 0., 0., 1.), vel=(.3, -.2, .5), angVel=(1., 0., 4.), facets=list(range(len(tris))))]
    S.finalize()
    O, E = compare(S, 150, check_every=25, label=f"peri moving facets homoDeform={homo}")
    compare_potential(O, E)
    assert np.array_equal(O.facets(), E.facets()) and not np.allclose(E.facets()[0], tris[0], atol=1e-3)
    assert (O.contacts(True)["ids"][:, 0] >= S.n).sum() > 0


def test_periodic_with_walls():
    S = sc.periodic_gas(300, cell=(.4, .4, .5), r=.02, psd_ratio=2.0, seed=19, walls=True)
    S.params.gravity = (0, 0, -9.81)
    O, E = compare(S, 150, check_every=25, label="peri-walls")
    compare_potential(O, E)


# ---------------------------------------------------------------------------------------------------------------
# slab decomposition: several subdomains on this one GPU (in-process transport), bit-identical to the oracle
# ---------------------------------------------------------------------------------------------------------------
def cluster(world, axis=0, scramble=False):
    from woo_b200.domain import LocalCluster
    return lambda S: LocalCluster(S, world, axis=axis, scramble=scramble)


@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
@pytest.mark.parametrize("world", [2, 3])
def test_decomposed_gas(law, world):
    S = sc.random_gas(500, box=.5, r=.02, psd_ratio=3.0, seed=21, law=law, en=.7 if law else float("nan"))
    O, E = compare(S, 300, check_every=25, label=f"decomposed gas world={world} law={law}", make=cluster(world))
    st = E.domain_stats()
    assert sum(d["migrated_in"] for d in st) > 0 and sum(d["rows_in"] for d in st) > 0, "test did not exercise migration with history: %r" % st
    assert sum(d["ghosts"] for d in st) > 0


def test_decomposed_scrambled_start_and_axes():
    for axis in (1, 2):
        S = sc.random_gas(300, box=.4, r=.02, psd_ratio=2.0, seed=23)
        compare(S, 120, check_every=30, label=f"decomposed scrambled axis={axis}", make=cluster(3, axis=axis, scramble=True))


@pytest.mark.parametrize("kind", ["gas", "periodic"])
def test_decomposed_rebalance_moves_cuts_and_keeps_bits(kind):
    """wdem_set_rebalance: lopsided cuts (one subdomain owns most of the scene) move to the quantiles at a full run; the result stays bit-identical to the oracle."""
    from woo_b200.domain import LocalCluster
    if kind == "gas":
        S = sc.random_gas(600, box=.5, r=.02, psd_ratio=3.0, seed=41); lo, hi = S.pos[:, 0].min(), S.pos[:, 0].max()
        cuts = [lo + .1 * (hi - lo), lo + .2 * (hi - lo)]
    else:
        S = sc.periodic_gas(500, cell=(.6, .4, .45), r=.02, psd_ratio=2.0, seed=43, gradV=[[0, 2.0, 0], [0, 0, 0], [0, 0, 0]]); cuts = [.1, .2]

    def make(S_):
        E = LocalCluster(S_, 3, axis=0, cuts=cuts); E.set_rebalance(.10, 2); return E
    O, E = compare(S, 200, check_every=25, label="rebalance " + kind, make=make)
    if kind == "periodic": compare_potential(O, E)
    new_cuts, nreb = E.domain()
    own = [d["owned"] for d in E.domain_stats()]
    assert nreb >= 1 and not np.allclose(new_cuts, cuts), (new_cuts, nreb)
    assert max(own) <= 1.25 * np.mean(own), "still unbalanced after re-balancing: %r" % own
    assert sum(own) == S.n


@pytest.mark.parametrize("kind", ["gas-hertz", "periodic"])
def test_decomposed_add_and_remove_spheres(kind):
    """wdem_add_spheres / wdem_remove_spheres in decomposed runs (the same call on every subdomain): insertion, removal, and a freed id handed out again, all
    bit-identical to the oracle — i.e. to the single-domain engine."""
    import math
    from woo_b200.domain import LocalCluster
    if kind == "gas-hertz":
        S = sc.random_gas(400, box=.5, r=.02, psd_ratio=2.0, seed=181, law=sc.LAW_HERTZ_DMT, en=.6); lo, hi = .05, .45
    else:
        S = sc.periodic_gas(400, cell=(.6, .4, .45), r=.02, psd_ratio=2.0, seed=183, gradV=[[0, 1.5, 0], [0, 0, 0], [0, 0, 0]]); lo, hi = .0, .4
    O = Oracle(S); E = LocalCluster(S, 3, axis=0)
    rng = np.random.default_rng(182)
    dens = S.mats[0][1]

    def check(label):
        so, sg = O.spheres(), E.spheres()
        oo = np.argsort(O.sphere_ids())
        assert len(sg["pos"]) == len(so["pos"]), label
        for k_ in ("pos", "vel", "angVel", "force", "torque"):
            assert np.array_equal(so[k_][oo], sg[k_]), f"{label}: {k_} differs"
        co, cg = O.contacts(True), E.contacts(True)
        assert [tuple(x) for x in co["ids"]] == [tuple(x) for x in cg["ids"]], label
        for k_ in ("uN", "force", "trsf", "stepCreated"):
            assert np.array_equal(co[k_], cg[k_]), f"{label}: contact {k_} differs"
        assert O.nFullRuns == E.scalars()["nFullRuns"] and O.dt == E.scalars()["dt"], label

    O.step(60); E.step(60); check("before")
    for batch in range(2):
        m = 30
        pos = rng.uniform(lo, hi, (m, 3)); rad = .02 * np.exp(rng.uniform(math.log(.5), math.log(2.), m))
        vel = rng.uniform(-1, 1, (m, 3)); ang = rng.uniform(-5, 5, (m, 3))
        mass = dens * (4 / 3.) * math.pi * rad ** 3; inertia = .4 * mass * rad ** 2
        first = O.add_spheres(pos, rad, mass, inertia, np.zeros(m, np.int32), vel=vel, angVel=ang)
        ids = np.arange(first, first + m, dtype=np.int32)
        E.add_spheres(ids, pos, rad, mass, inertia, vel=vel, angVel=ang)
        for k in range(2):
            O.step(40); E.step(40); check(f"after insertion {batch}.{k}")
    live = np.sort(O.sphere_ids())
    gone = np.sort(rng.choice(live, 50, replace=False)).astype(np.int32)
    assert np.isin(O.contacts(True)["ids"], gone).any(), "the removed spheres should have contacts"
    O.remove_spheres(gone); E.remove_spheres(gone)
    # a newcomer takes over the oldest freed id in the same breath (ParticleContainer::findFreeId)
    rad = np.array([.025]); mass = dens * (4 / 3.) * math.pi * rad ** 3; pos = np.array([[.5 * (lo + hi)] * 3]); vel = np.array([[.2, -.1, .3]])
    new_id = O.add_spheres(pos, rad, mass, .4 * mass * rad ** 2, np.zeros(1, np.int32), vel=vel)
    assert new_id == int(gone[0])
    E.add_spheres(np.array([new_id], np.int32), pos, rad, mass, .4 * mass * rad ** 2, vel=vel)
    for k in range(3):
        O.step(40); E.step(40); check(f"after removal {k}")
    st = E.domain_stats()
    assert sum(d["owned"] for d in st) == len(O.sphere_ids())
    from woo_b200.engine import WdemError
    with pytest.raises(WdemError):
        E.add_spheres(np.array([new_id], np.int32), pos, rad, mass, .4 * mass * rad ** 2)      # id in use


def test_decomposed_deposition():
    S = sc.deposition(2000, seed=1)
    compare(S, 400, check_every=100, label="decomposed deposition", make=cluster(4))


def test_decomposed_facets_and_flags():
    S = sc.facet_box_scene(300, seed=2)
    S.flags[::7] = sc.DOF_ALL; S.flags[1::7] = 1 | 8
    compare(S, 150, check_every=25, label="decomposed facets", make=cluster(2))


@pytest.mark.parametrize("world", [2, 3])
def test_decomposed_periodic_shear(world):
    g = [[0, 2.0, 0], [0, 0, 0], [0, 0, 0]]
    S = sc.periodic_gas(500, cell=(.6, .4, .45), r=.02, psd_ratio=2.0, seed=29, gradV=g)
    O, E = compare(S, 200, check_every=25, label=f"decomposed periodic world={world}", make=cluster(world))
    compare_potential(O, E)
    S = sc.periodic_gas(300, cell=(.45, .45, .45), r=.02, psd_ratio=3.0, seed=17, gradV=[[-1.0, .5, 0], [0, -1.0, 0], [.3, 0, -1.0]], shear0=.3)
    compare(S, 100, check_every=25, label="decomposed periodic skew", make=cluster(world, axis=1))


# ---------------------------------------------------------------------------------------------------------------
# end-to-end path (host buffers in, host buffers out): the chunk-pipelined wdem_step_e2e must give the resident path's bits
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,stats", [(37, True), (3000, True), (3000, False)])
def test_step_e2e_pipelined_matches_resident_and_oracle(n, stats, monkeypatch):
    S = sc.random_gas(n, box=.3 if n < 100 else .9, r=.02, psd_ratio=2.0, seed=21)
    O = Oracle(S); A = engine_for(S); B = engine_for(S)
    A.step(30); B.step(30); O.step(30)                      # rebuilds have reordered the slots: upload order != slot order
    s = B.spheres()
    pos, vel, w = s["pos"].copy(), s["vel"].copy(), s["angVel"].copy()
    po, vo = np.empty_like(pos), np.empty_like(vel)
    for k in range(40):
        ek, st = B.step_e2e(1, pos, vel, w, po, vo, stats=stats)      # stats=False: returns while the step's contact pass is still running
        A.step(1); O.step(1)
        sa = A.spheres()
        assert np.array_equal(po, sa["pos"]) and np.array_equal(vo, sa["vel"]), f"e2e outputs differ from the resident path at step {k}"
        assert abs(ek - A.kinetic_energy()) <= 1e-13 * abs(ek)    # block sums run over slots; a collider run inside the step reorders them
        if stats: assert st.n_real == len(A.contacts(True)["ids"])
        pos, po = po, pos; vel, vo = vo, vel
        w = B.spheres()["angVel"].copy()                    # angVel is an input of the call only; read it back for the next one
    so, sb = O.spheres(), B.spheres()
    for k_ in ("pos", "vel", "angVel", "force", "torque"):
        assert np.array_equal(so[k_], sb[k_]), k_
    assert A.scalars() == B.scalars()
    # the sequential fallback gives the same bits
    monkeypatch.setenv("WDEM_E2E_SEQUENTIAL", "1")
    ek2, _ = B.step_e2e(1, pos, vel, w, po, vo); A.step(1)
    assert np.array_equal(po, A.spheres()["pos"]) and abs(ek2 - A.kinetic_energy()) <= 1e-13 * abs(ek2)


def test_step_e2e_owned_on_a_single_domain_engine_in_any_id_order():
    """wdem_step_e2e_owned also serves a single-domain engine: the state travels by Particle.id, here in a scrambled order; outputs come back in that order."""
    S = sc.random_gas(500, box=.5, r=.02, psd_ratio=2.0, seed=151)
    A = engine_for(S); B = engine_for(S)
    A.step(40); B.step(40)
    ids = S.sphere_ids().astype(np.int32); perm = np.random.default_rng(5).permutation(len(ids)); ids_p = np.ascontiguousarray(ids[perm])
    s = B.spheres(); pos, vel, w = (np.ascontiguousarray(s[k][perm]) for k in ("pos", "vel", "angVel"))
    po, vo = np.empty_like(pos), np.empty_like(vel)
    for k in range(30):
        ek, changed, _ = B.step_e2e_owned(ids_p, pos, vel, w, po, vo, stats=(k % 2 == 0))
        A.step(1); sa = A.spheres()
        assert not changed and np.array_equal(po, sa["pos"][perm]) and np.array_equal(vo, sa["vel"][perm]), f"step {k}"
        assert abs(ek - A.kinetic_energy()) <= 1e-13 * abs(ek)
        pos, po = po, pos; vel, vo = vo, vel
        w = np.ascontiguousarray(sa["angVel"][perm])
    sb = B.spheres()
    for k_ in ("pos", "vel", "angVel", "force", "torque"):
        assert np.array_equal(A.spheres()[k_], sb[k_]), k_
    from woo_b200.engine import WdemError
    bad = ids_p.copy(); bad[3] = 10 ** 6
    with pytest.raises(WdemError):
        B.step_e2e_owned(bad, pos, vel, w, po, vo)


@pytest.mark.parametrize("stats", [True, False])
def test_step_e2e_owned_matches_resident_decomposed_run(stats):
    """wdem_step_e2e_owned: every subdomain round-trips the spheres it owns (by id) through host buffers each step; results equal the resident run of the same
    decomposition — and so the oracle — bit for bit, across collider runs and migrations (the host then fetches the new owned set)."""
    import threading
    from woo_b200.domain import LocalCluster, _owned
    S = sc.random_gas(600, box=.5, r=.02, psd_ratio=3.0, seed=141)
    O = Oracle(S); A = LocalCluster(S, 2, axis=0)
    A.step(30); O.step(30)
    bufs = []
    for E in A.E:
        o = _owned(E); bufs.append(dict(ids=o["ids"].copy(), pos=o["pos"].copy(), vel=o["vel"].copy(), w=o["angVel"].copy(), po=np.empty_like(o["pos"]), vo=np.empty_like(o["vel"])))
    refreshed = 0
    for k in range(60):
        res = [None, None]

        def run(r):
            b = bufs[r]; res[r] = A.E[r].step_e2e_owned(b["ids"], b["pos"], b["vel"], b["w"], b["po"], b["vo"], stats=stats)
        th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
        [t.start() for t in th]; [t.join() for t in th]
        O.step(1)
        so = O.spheres()
        for r, E in enumerate(A.E):
            b = bufs[r]
            assert np.array_equal(b["po"], so["pos"][b["ids"]]) and np.array_equal(b["vo"], so["vel"][b["ids"]]), f"outputs of subdomain {r} differ from the oracle at step {k}"
            if res[r][1]:      # spheres migrated: fetch the new owned set
                o = _owned(E); refreshed += 1
                bufs[r] = dict(ids=o["ids"].copy(), pos=o["pos"].copy(), vel=o["vel"].copy(), w=o["angVel"].copy(), po=np.empty_like(o["pos"]), vo=np.empty_like(o["vel"]))
            else:
                b["pos"], b["po"] = b["po"], b["pos"]; b["vel"], b["vo"] = b["vo"], b["vel"]
                b["w"] = so["angVel"][b["ids"]].copy()          # angVel is an input only
    assert refreshed > 0, "the test did not see a migration"
    sg = A.spheres()
    for k_ in ("pos", "vel", "angVel", "force", "torque"):
        assert np.array_equal(O.spheres()[k_], sg[k_]), k_
    assert O.nFullRuns == A.scalars()["nFullRuns"]
    from woo_b200.engine import WdemError
    b = bufs[0]
    with pytest.raises(WdemError):
        A.E[0].step_e2e_owned(b["ids"][:-1], b["pos"][:-1], b["vel"][:-1], b["w"][:-1], b["po"][:-1], b["vo"][:-1])      # not the whole owned set


# ---------------------------------------------------------------------------------------------------------------
# wdem_upload_contacts / wdem_set_scene: a run resumed from downloaded state + contact history continues bit for bit.
# The candidate set of the resumed engine differs (its collider runs at other steps), which must not matter: real contacts, their history and the
# per-particle summation order do not depend on it.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["gas-linear-rolling", "gas-hertz", "facets", "periodic", "gas-dyndt"])
def test_resume_from_uploaded_contacts(kind):
    from woo_b200.engine import Engine
    import copy
    if kind == "gas-dyndt":      # DynDt keeps firing across the resume: its PeriodicEngine counters travel through wdem_set_periodic_state, dt through the config
        S = sc.random_gas(400, box=.5, r=.02, psd_ratio=3.0, seed=36)
    elif kind == "gas-linear-rolling":
        S = sc.random_gas(400, box=.5, r=.02, psd_ratio=3.0, seed=31); S.params.relRollStiff = .3; S.params.relTwistStiff = .2; S.params.rollTanPhi = .4
    elif kind == "gas-hertz":
        S = sc.random_gas(400, box=.5, r=.02, psd_ratio=2.0, seed=32, law=sc.LAW_HERTZ_DMT, en=.7)
    elif kind == "facets":
        S = sc.facet_box_scene(300, seed=33)
    else:
        S = sc.periodic_gas(300, seed=34, gradV=[[0, 1., 0], [0, 0, 0], [0, 0, 0]], shear0=.1)
    S.params.dynDtPeriod = 40 if kind == "gas-dyndt" else 0
    O = Oracle(S); A = engine_for(S)
    A.step(150); O.step(150)
    s, c, sca = A.spheres(), A.contacts(True), A.scalars()
    assert len(c["ids"]) > 10
    P = copy.copy(S.params); P.dt = sca["dt"]
    B = Engine(P)
    B.spec = S
    B.upload_materials(S.mats)
    m, I = S.mass_inertia()
    B.upload_spheres(S.sphere_ids(), s["pos"], S.radius, m, I, S.mat, S.mask, S.flags, s["vel"], s["angVel"], ori=s["ori"])
    B._chk(B.L.wdem_upload_forces(B.h, s["force"].ctypes.data_as(__import__("ctypes").c_void_p), s["torque"].ctypes.data_as(__import__("ctypes").c_void_p)))
    B.upload_walls(S.wall_ids(), S.walls); B.upload_facets(S.facet_ids(), S.facets)
    if S.cell is not None:
        B.set_cell(A.cell_hsize(), S.cell.get("gradV"), S.cell.get("nextGradV"), S.cell.get("homoDeform", 4))
    B.set_scene(sca["step"], sca["time"])
    if kind == "gas-dyndt":
        assert A.periodic_state() == (120, 80)
        B.set_periodic_state(*A.periodic_state())
    B.upload_contacts(c)
    for k in range(4):
        A.step(40); B.step(40); O.step(40)
        assert A.scalars()["dt"] == B.scalars()["dt"]
        sa, sb, so = A.spheres(), B.spheres(), O.spheres()
        for k_ in ("pos", "vel", "angVel", "force", "torque"):
            assert np.array_equal(sa[k_], sb[k_]), f"{kind}: {k_} of the resumed run differs after {40 * (k + 1)} steps"
            assert np.array_equal(so[k_], sb[k_]), f"{kind}: {k_} differs from the oracle"
        ca, cb = A.contacts(True), B.contacts(True)
        assert [tuple(x) for x in ca["ids"]] == [tuple(x) for x in cb["ids"]]
        for k_ in ("uN", "force", "torque", "trsf", "contPt", "physRaw", "stepCreated", "cellDist"):
            assert np.array_equal(ca[k_], cb[k_]), f"{kind}: contact {k_} of the resumed run differs"
    assert A.scalars()["step"] == B.scalars()["step"] and A.scalars()["time"] == B.scalars()["time"]


def test_upload_contacts_rejects_bad_input():
    from woo_b200.engine import Engine, WdemError
    S = sc.random_gas(50, box=.3, r=.02, seed=35)
    A = engine_for(S); A.step(30); c = A.contacts(True)
    with pytest.raises(WdemError):
        A.upload_contacts(c)                                   # after the first step
    B = engine_for(S)
    bad = {k: v.copy() for k, v in c.items()}
    ss = np.nonzero(bad["ids"][:, 0] >= len(S.walls))[0] if S.boundaries_first else np.nonzero(bad["ids"][:, 0] < S.n)[0]
    if len(ss):
        bad["ids"][ss[0]] = bad["ids"][ss[0]][::-1]            # id1 > id2
        with pytest.raises(WdemError):
            B.upload_contacts(bad)
    bad = {k: v.copy() for k, v in c.items()}; bad["ids"][0, 1] = 10 ** 6
    with pytest.raises(WdemError):
        B.upload_contacts(bad)


# ---------------------------------------------------------------------------------------------------------------
# ContactLoop.evalStress: stress = sum F (x) dPos / V over real contacts, periodic scenes
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2])
def test_periodic_stress_matches_oracle(world):
    S = sc.periodic_gas(500, seed=51, gradV=[[0, .5, 0], [0, 0, 0], [0, 0, -.3]], shear0=.15, psd_ratio=2.0)
    O = Oracle(S)
    if world == 1:
        E = engine_for(S); stress = E.stress
    else:
        from woo_b200.domain import LocalCluster
        E = LocalCluster(S, world, axis=0); stress = lambda: sum(e.stress() for e in E.E)
    for k in range(3):
        O.step(60); E.step(60)
        so, sg = O.stress(), stress()
        scale = np.abs(so).max()
        assert scale > 0
        # the sum over contacts runs in table order on the device and in canonical order in the oracle: 1e-10 relative (north_star tolerance)
        assert np.allclose(sg, so, rtol=1e-10, atol=1e-10 * scale), f"stress differs after {60 * (k + 1)} steps:\n{sg}\n{so}"
    with pytest.raises(Exception):
        engine_for(sc.random_gas(50, box=.3, r=.02, seed=52)).stress()      # aperiodic: not defined (ContactLoop.cpp:123)


# Scene.trackEnergy + ContactLoop.evalStress in a periodic scene: the work of the velocity gradient, -tr(gradV.midStress) dt midVol per step, is the "gradV" energy
# (ContactLoop.cpp:243-254)
def test_gradv_work_matches_oracle():
    S = sc.periodic_gas(500, seed=53, gradV=[[-.8, .5, 0], [0, -.6, 0], [0, 0, -.7]], shear0=.1, psd_ratio=2.0)      # a cell being compressed and sheared
    O = Oracle(S); O.track_energy(True); O.eval_stress(True)
    from woo_b200.engine import Engine
    E = Engine(S.params, evalStress=1).load(S); E.track_energy(True)
    for k in range(3):
        O.step(80); E.step(80)
        wo, wg = O.gradv_work(), E.gradv_work()
        assert wo != 0 and wg == pytest.approx(wo, rel=1e-9), f"gradV work {wg} vs {wo} after {80 * (k + 1)} steps"
        eo, eg = O.energy(), E.energy()
        for name in eo:
            assert eg[name] == pytest.approx(eo[name], rel=1e-9, abs=1e-12 * max(abs(v) for v in eo.values()))
    assert np.array_equal(O.spheres()["pos"], E.spheres()["pos"])          # bookkeeping only: the trajectory is untouched


# ---------------------------------------------------------------------------------------------------------------
# moving walls (piston): a wall uploaded with a velocity is in DemField.nodes and is translated by Leapfrog (Particle.cpp:638-640, Leapfrog.cpp:195-199,275-278);
# its force is the sum over the step's contacts
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
def test_moving_wall_piston(law, world):
    S = sc.random_gas(300, box=.4, r=.02, psd_ratio=2.0, seed=61, law=law, en=.6 if law else float("nan"))
    top = [w for w in S.walls if w["axis"] == 2 and w["sense"] == -1]
    assert len(top) == 1
    top[0]["vel"] = (0., 0., -.8)                                  # the lid comes down and compresses the gas
    S.walls[0]["vel"] = (.05, 0., 0.) if S.walls[0]["axis"] != 0 else (0., .05, 0.)    # a wall sliding in its own plane: only the contact kinematics see it
    O = Oracle(S)
    if world == 1:
        E = engine_for(S); walls = E.walls
    else:   # every subdomain moves its copy of the walls; the force on a wall is the sum of the subdomains' shares (not bitwise: the order of the sum differs)
        E = cluster(world)(S)
        def walls():
            parts = [e.walls() for e in E.E]
            return dict(pos=parts[0]["pos"], force=sum(p["force"] for p in parts), torque=sum(p["torque"] for p in parts))
    z0 = walls()["pos"][S.walls.index(top[0]), 2]
    for k in range(5):
        O.step(60); E.step(60)
        wo, wg = O.walls(), walls()
        assert np.array_equal(wo["pos"], wg["pos"]), "wall positions differ"
        so, sg = O.spheres(), E.spheres()
        for k_ in ("pos", "vel", "angVel", "force", "torque"):
            assert np.array_equal(so[k_], sg[k_]), f"{k_} differs after {60 * (k + 1)} steps"
        co, cg = O.contacts(True), E.contacts(True)
        assert [tuple(x) for x in co["ids"]] == [tuple(x) for x in cg["ids"]]
        assert O.nFullRuns == E.scalars()["nFullRuns"]
        mv = [i for i, w in enumerate(S.walls) if any(w.get("vel", (0, 0, 0)))]
        if world == 1:
            assert np.array_equal(wo["force"][mv], wg["force"][mv]) and np.array_equal(wo["torque"][mv], wg["torque"][mv]), "force on the moving walls differs"
        else:
            sc_ = max(np.abs(wo["force"][mv]).max(), 1e-300)
            assert np.allclose(wo["force"][mv], wg["force"][mv], rtol=1e-10, atol=1e-12 * sc_) and np.allclose(wo["torque"][mv], wg["torque"][mv], rtol=1e-10, atol=1e-12 * sc_)
        # walls at rest are not in DemField.nodes: nothing resets their force, it accumulates over the steps (ContactLoop.cpp:270-276); per-step partial
        # sums on the device, so equal to the oracle's running sum to rounding only
        rest = [i for i in range(len(S.walls)) if i not in mv]
        sr = max(np.abs(wo["force"][rest]).max(), 1e-300)
        assert sr > 0 and np.allclose(wo["force"][rest], wg["force"][rest], rtol=1e-9, atol=1e-9 * sr), "accumulated force on the walls at rest differs"
        assert np.allclose(wo["torque"][rest], wg["torque"][rest], rtol=1e-9, atol=1e-9 * max(np.abs(wo["torque"][rest]).max(), 1e-300))
    zi = S.walls.index(top[0])
    assert walls()["pos"][zi, 2] < z0 - .8 * 299 * E.scalars()["dt"] * .5 and np.abs(walls()["force"][zi]).max() > 0


def test_gpu_matches_reference_python_models():
    """The CUDA path against golden vectors made by the reference's own py/models.py (tests/golden/ref_models.json)."""
    from test_oracle_kat import check_against_ref_models
    check_against_ref_models(engine_for, lambda E: E.contacts(True))


# ---------------------------------------------------------------------------------------------------------------
# Scene.trackEnergy: EnergyTracker entries, device vs oracle.  Sums run over thousands of terms in different orders (atomics): 1e-9 relative.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["linear", "linear-noslip", "hertz", "periodic", "decomposed"])
def test_track_energy_matches_oracle(kind):
    if kind == "linear":
        S = sc.random_gas(400, box=.5, r=.02, psd_ratio=3.0, seed=71)
        S.flags[::9] = sc.DOF_ALL; S.flags[1::9] = 256; S.flags[2::9] = sc.GRAVITY_SKIP; S.flags[3::9] = sc.DAMPING_SKIP; S.flags[4::9] = 1 | 16
    elif kind == "linear-noslip":
        S = sc.random_gas(400, box=.5, r=.02, psd_ratio=2.0, seed=72); S.params.noSlip = 1
    elif kind == "hertz":
        S = sc.random_gas(400, box=.5, r=.02, psd_ratio=2.0, seed=73, law=sc.LAW_HERTZ_DMT, en=.6)
    elif kind == "periodic":
        S = sc.periodic_gas(400, seed=74, gradV=[[0, 1., 0], [0, 0, 0], [0, 0, 0]], shear0=.1, psd_ratio=2.0)
    else:
        S = sc.random_gas(500, box=.5, r=.02, psd_ratio=3.0, seed=75)
    O = Oracle(S); O.track_energy(True)
    if kind == "decomposed":
        from woo_b200.domain import LocalCluster
        E = LocalCluster(S, 2, axis=0)
        for e in E.E:
            e.track_energy(True)
        energy = lambda: {k: sum(e.energy()[k] for e in E.E) for k in E.E[0].energy()}
    else:
        E = engine_for(S); E.track_energy(True); energy = E.energy
    seen = set()
    for k in range(4):
        O.step(80); E.step(80)
        eo, eg = O.energy(), energy()
        scale = max(abs(v) for v in eo.values())
        for name in eo:
            assert eg[name] == pytest.approx(eo[name], rel=1e-9, abs=1e-12 * scale), f"{kind}: energy[{name}] after {80 * (k + 1)} steps: {eg[name]} vs {eo[name]}"
            if eo[name] != 0:
                seen.add(name)
        # tracking does not touch the trajectory
        assert np.array_equal(O.spheres()["pos"], E.spheres()["pos"])
    want = {"linear": {"kinetic", "grav", "nonviscDamp", "elast", "plast"}, "linear-noslip": {"kinetic", "elast", "broken"}, "hertz": {"kinetic", "elast", "viscN", "viscT", "plast"},
            "periodic": {"kinetic", "elast"}, "decomposed": {"kinetic", "elast", "plast"}}[kind]
    assert want <= seen, f"{kind}: entries never became non-zero: {want - seen}"


# ---------------------------------------------------------------------------------------------------------------
# kinematic facet groups: a clump of facets whose blocked master node carries an angular velocity — a drum that really rotates
# (examples/mill.py:38-42; Clump.cpp:293-307, Leapfrog.cpp:195-199,275-297)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2])
def test_rotating_drum_facet_group(world):
    S = sc.drum(600, div=16, axial=4, vz0=1., vfluct=.5, omega=8., rotate=True)
    S.facet_groups[0]["vel"] = (.02, 0., 0.)                     # and creeps along its axis
    f0 = np.array([f["v"] for f in S.facets])
    O, E = compare(S, 8000, check_every=2000, label=f"rotating drum world={world}", make=cluster(world) if world > 1 else None)
    fo = O.facets()
    fg = E.facets() if world == 1 else E.E[0].facets()
    assert np.array_equal(fo, fg), "facet node positions differ from the oracle"
    assert np.abs(fo - f0).max() > .05 * np.abs(f0).max(), "the drum did not move"
    ids = O.contacts(True)["ids"]
    assert (ids[:, 0] >= S.n + len(S.walls)).sum() > 30, "no facet contacts formed"
    # the turning wall drags the bed along: the spheres' mean angular momentum about the axis has the sign of the rotation
    s = O.spheres(); ybar, zbar = s["pos"][:, 1].mean(), s["pos"][:, 2].mean()
    assert ((s["pos"][:, 1] - ybar) * s["vel"][:, 2] - (s["pos"][:, 2] - zbar) * s["vel"][:, 1]).mean() > 0


# ---------------------------------------------------------------------------------------------------------------
# wdem_add_spheres (the insertion half of Inlets): ParticleContainer.add in mid-run without re-uploading the scene
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
def test_add_spheres_mid_run(law):
    import math
    S = sc.random_gas(300, box=.5, r=.02, psd_ratio=2.0, seed=81, law=law, en=.6 if law else float("nan"))
    O = Oracle(S); E = engine_for(S)
    rng = np.random.default_rng(82)
    dens = S.mats[0][1]
    for batch in range(3):
        O.step(70); E.step(70)
        m = 40
        pos = rng.uniform(.05, .45, (m, 3)); rad = .02 * np.exp(rng.uniform(math.log(.4), math.log(2.5), m))      # also smaller and larger than anything so far
        vel = rng.uniform(-1, 1, (m, 3)); ang = rng.uniform(-5, 5, (m, 3))
        mass = dens * (4 / 3.) * math.pi * rad ** 3; inertia = .4 * mass * rad ** 2
        first = O.add_spheres(pos, rad, mass, inertia, np.zeros(m, np.int32), vel=vel, angVel=ang)
        ids = np.arange(first, first + m, dtype=np.int32)
        E.add_spheres(ids, pos, rad, mass, inertia, vel=vel, angVel=ang)
        for k in range(3):
            O.step(40); E.step(40)
            # the oracle lists spheres in the order they were added, and so does the engine (upload order continues)
            so = O.spheres(); no = len(so["pos"])
            sg = E.spheres()
            assert len(sg["pos"]) == no == S.n + (batch + 1) * m
            for k_ in ("pos", "vel", "angVel", "force", "torque"):
                assert np.array_equal(so[k_], sg[k_]), f"{k_} differs after batch {batch}, check {k}"
            co, cg = O.contacts(True), E.contacts(True)
            assert [tuple(x) for x in co["ids"]] == [tuple(x) for x in cg["ids"]] and np.array_equal(co["force"], cg["force"]) and np.array_equal(co["trsf"], cg["trsf"])
            assert O.nFullRuns == E.scalars()["nFullRuns"] and O.dt == E.scalars()["dt"]
    assert np.array_equal(co["stepCreated"], cg["stepCreated"])
    from woo_b200.engine import WdemError
    with pytest.raises(WdemError):
        E.add_spheres(ids[:1], pos[:1], rad[:1], mass[:1], inertia[:1])          # id already in use


# ---------------------------------------------------------------------------------------------------------------
# wdem_remove_spheres (the deletion half: Outlets): ParticleContainer.remove in mid-run.  Checked twice: against the oracle (which takes the removed
# particle out of the game but keeps its slot) and through the resume path: removing spheres in a running engine == a fresh engine started from the same state without them.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
def test_remove_spheres_mid_run(law):
    import copy, ctypes
    from woo_b200.engine import Engine, WdemError
    S = sc.random_gas(400, box=.5, r=.02, psd_ratio=2.5, seed=95, law=law, en=.6 if law else float("nan")); S.params.dynDtPeriod = 0
    X = engine_for(S); X.step(150)
    O = Oracle(S); O.step(150)
    s, c, sca = X.spheres(), X.contacts(True), X.scalars()
    ids_all = S.sphere_ids()
    rng = np.random.default_rng(96)
    gone = np.sort(rng.choice(ids_all, 60, replace=False)).astype(np.int32)
    assert np.isin(c["ids"], gone).any(), "the removed spheres should have contacts"
    keep = ~np.isin(ids_all, gone)
    X.remove_spheres(gone); O.remove_spheres(gone)
    with pytest.raises(WdemError):
        X.contacts(True)                                  # the table still refers to the old slots until the next step
    # a sphere added in the same breath (removal + insertion before the next collider run)
    m_, I_ = S.mass_inertia()
    new_id = int(gone[0])                                 # the id ParticleContainer.add hands out next: the oldest freed one (findFreeId, ParticleContainer.cpp:23-36)
    newp = dict(pos=np.array([[.25, .25, .45]]), rad=np.array([.03]), vel=np.array([[0., 0., -2.]]))
    mass_new = S.mats[0][1] * (4 / 3.) * np.pi * newp["rad"] ** 3
    X.add_spheres(np.array([new_id], np.int32), newp["pos"], newp["rad"], mass_new, .4 * mass_new * newp["rad"] ** 2, vel=newp["vel"])
    assert O.add_spheres(newp["pos"], newp["rad"], mass_new, .4 * mass_new * newp["rad"] ** 2, np.zeros(1, np.int32), vel=newp["vel"]) == new_id
    # the fresh engine: remaining spheres with their state, forces and the contacts among them, plus the new sphere
    pr = copy.copy(S.params); pr.dt = sca["dt"]
    Y = Engine(pr); Y.spec = S
    Y.upload_materials(S.mats)
    cat = lambda a, b: np.concatenate([a, b])
    Y.upload_spheres(cat(ids_all[keep], [new_id]), cat(s["pos"][keep], newp["pos"]), cat(S.radius[keep], newp["rad"]), cat(m_[keep], mass_new), cat(I_[keep], .4 * mass_new * newp["rad"] ** 2),
                     cat(S.mat[keep], [0]), cat(S.mask[keep], [S.mask[0]]), cat(S.flags[keep], [0]), cat(s["vel"][keep], newp["vel"]), cat(s["angVel"][keep], np.zeros((1, 3))),
                     ori=cat(s["ori"][keep], [[0, 0, 0, 1.]]))
    f = np.ascontiguousarray(cat(s["force"][keep], np.zeros((1, 3)))); t = np.ascontiguousarray(cat(s["torque"][keep], np.zeros((1, 3))))
    Y._chk(Y.L.wdem_upload_forces(Y.h, f.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p)))
    Y.upload_walls(S.wall_ids(), S.walls)
    Y.set_scene(sca["step"], sca["time"])
    alive = ~np.isin(c["ids"], gone).any(axis=1)
    Y.upload_contacts({k: v[alive] for k, v in c.items()})
    for k in range(3):
        X.step(50); Y.step(50); O.step(50)
        sx, sy, so = X.spheres(), Y.spheres(), O.spheres()
        assert len(sx["pos"]) == S.n - len(gone) + 1 == len(so["pos"])
        # the engine lists spheres in upload order (the new one last), the oracle by container slot (the new one took over a freed slot): compare by id
        ox, oo = np.argsort(cat(ids_all[keep], [new_id])), np.argsort(O.sphere_ids())
        for k_ in ("pos", "vel", "angVel", "force", "torque"):
            assert np.array_equal(sx[k_], sy[k_]), f"{k_} differs {50 * (k + 1)} steps after the removal"
            assert np.array_equal(sx[k_][ox], so[k_][oo]), f"{k_} differs from the oracle {50 * (k + 1)} steps after the removal"
        assert X.scalars()["nFullRuns"] == O.nFullRuns
        cx, cy = X.contacts(True), Y.contacts(True)
        assert np.array_equal(cx["ids"], cy["ids"]) and not np.isin(cx["ids"], gone[1:]).any()          # gone[0] is back: the newcomer reuses that id
        for k_ in ("uN", "force", "trsf", "stepCreated", "physRaw"):
            assert np.array_equal(cx[k_], cy[k_]), k_
    with pytest.raises(WdemError):
        X.remove_spheres(gone[1:2])                       # already gone (gone[0] is back in the game under its old id)


# A freed id is handed out again (ParticleContainer::findFreeId): the newcomer must not inherit anything from the removed particle's contacts, even when it lands
# right where the old one was and touches the same partners before the next collider run (round-1 advisor finding).  Checked against a FRESH engine started from
# the same state, not against the oracle: the reference's incremental sweep keeps the bounds of a vanished particle where they were (InsertionSortCollider.cpp:479-486)
# and only reacts to inversions, so a particle re-inserted near its predecessor gets NO contact with the old neighbours there until some bound inverts — a quirk of
# the sweep (SURVEY A.18 class) that the from-scratch candidate search does not reproduce.
@pytest.mark.parametrize("law", [sc.LAW_IDEALELPL, sc.LAW_HERTZ_DMT])
def test_id_reuse_after_remove_starts_with_fresh_contacts(law):
    import math, copy, ctypes
    from woo_b200.engine import Engine
    S = sc.random_gas(300, box=.45, r=.02, psd_ratio=2.0, seed=131, law=law, en=.6 if law else float("nan")); S.params.dynDtPeriod = 0
    X = engine_for(S)
    X.step(120)
    c, s, sca = X.contacts(True), X.spheres(), X.scalars()
    ss = c["ids"][(c["ids"][:, 0] < S.n) & (np.abs(c["force"][:, 1:]).max(axis=1) > 0)]          # sphere-sphere contacts that carry shear history
    cnt = np.bincount(ss.reshape(-1), minlength=S.n)
    g = int(np.argmax(cnt)); assert cnt[g] >= 2, "need a sphere with several sheared contacts"
    partners = set(int(x) for x in ss[(ss == g).any(axis=1)].reshape(-1)) - {g}
    X.remove_spheres(np.array([g], np.int32))
    pos = s["pos"][g:g + 1].copy(); rad = S.radius[g:g + 1] * .97; vel = np.array([[.3, -.2, .1]]); ang = np.array([[5., 0., -3.]])
    mass = S.mats[0][1] * (4 / 3.) * math.pi * rad ** 3; inertia = .4 * mass * rad ** 2
    X.add_spheres(np.array([g], np.int32), pos, rad, mass, inertia, vel=vel, angVel=ang)          # the same id again, where the old sphere was
    # the fresh engine: the remaining spheres with their state, forces and the contacts among them, plus the newcomer (upload order as in X: kept, then new)
    keep = np.arange(S.n) != g; ids_all = S.sphere_ids(); m_, I_ = S.mass_inertia(); cat = lambda a, b: np.concatenate([a, b])
    pr = copy.copy(S.params); pr.dt = sca["dt"]
    Y = Engine(pr); Y.spec = S
    Y.upload_materials(S.mats)
    Y.upload_spheres(cat(ids_all[keep], [g]), cat(s["pos"][keep], pos), cat(S.radius[keep], rad), cat(m_[keep], mass), cat(I_[keep], inertia), cat(S.mat[keep], [0]),
                     cat(S.mask[keep], [S.mask[0]]), cat(S.flags[keep], [0]), cat(s["vel"][keep], vel), cat(s["angVel"][keep], ang), ori=cat(s["ori"][keep], [[0, 0, 0, 1.]]))
    f = np.ascontiguousarray(cat(s["force"][keep], np.zeros((1, 3)))); t = np.ascontiguousarray(cat(s["torque"][keep], np.zeros((1, 3))))
    Y._chk(Y.L.wdem_upload_forces(Y.h, f.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p)))
    Y.upload_walls(S.wall_ids(), S.walls)
    Y.set_scene(sca["step"], sca["time"])
    alive = ~(c["ids"] == g).any(axis=1)
    Y.upload_contacts({k: v[alive] for k, v in c.items()})
    step0 = sca["step"]
    for k in range(1, 4):
        n = 1 if k == 1 else 30
        X.step(n); Y.step(n)
        sx, sy = X.spheres(), Y.spheres()
        for k_ in ("pos", "vel", "angVel", "force", "torque"):
            assert np.array_equal(sx[k_], sy[k_]), f"{k_} differs from the fresh engine after id reuse (check {k})"
        cx, cy = X.contacts(True), Y.contacts(True)
        assert [tuple(x) for x in cx["ids"]] == [tuple(x) for x in cy["ids"]]
        for k_ in ("uN", "force", "trsf", "stepCreated", "physRaw"):
            assert np.array_equal(cx[k_], cy[k_]), f"contact {k_} differs from the fresh engine after id reuse"
        if k == 1:
            mine = (cx["ids"] == g).any(axis=1)
            touched = set(int(x) for x in cx["ids"][mine].reshape(-1)) - {g}
            assert touched & partners, "the newcomer should touch partners of the removed sphere"
            assert (cx["stepCreated"][mine] == step0).all() and np.all(cx["force"][mine][:, 1:] == 0.), "contacts of the newcomer must be fresh: created now, no shear history"
    assert X.scalars()["nFullRuns"] >= 1
