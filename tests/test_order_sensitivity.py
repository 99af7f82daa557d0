"""Parity must not hinge on the one thing that cannot be pinned here: the order in which Eigen adds the three terms of dot()/squaredNorm(),
Matrix3*Vector3 rows and Matrix3::trace() (Eigen is not installed; DESIGN.md section 2 restates it from the Eigen sources).  The order is a
compile-time switch shared by woo_b200/csrc/wdem_math.cuh and oracle/dem_oracle.cpp (WD_SUM3_ORDER / WD_TRACE_ORDER / WD_MATPROD_ORDER);
`alt` = every order flipped (libwdem_alt.so / libdem_oracle_alt.so).  north_star's two levels are checked between the two orders:

* from an identical Scene state, one step: the set of real contacts is identical and per-particle / per-contact forces agree within 1e-10
  relative (measured: ~1e-15);
* multi-thousand-step runs (trajectories diverge chaotically, as they do between any two builds of the reference): coordination number and
  packing fraction within 1 %, kinetic energy statistically (window mean; 1500-sphere scenes fluctuate by more than 1 % on their own).

CPU part: oracle vs oracle-alt.  GPU part (-m gpu): CUDA engine vs CUDA-alt from a developed state moved through wdem_upload_contacts, and
CUDA-alt == oracle-alt bit for bit (the switch means the same thing on both sides)."""
import ctypes as C
import numpy as np
import pytest
from woo_b200 import scenes as sc
from woo_b200 import observables as ob
from _harness import Oracle

REL_FORCE_TOL = 1e-10   # north_star: "per-contact normal/shear forces and per-particle force/torque after one step must match within 1e-10 relative"


def cfg_scenes():
    """cfg1..cfg5 of BASELINE.json at oracle-friendly sizes + the dense-gas stress scenes of both laws (contacts exist from step 0)."""
    S1 = sc.deposition(1500, seed=1, spacing_factor=2.05); S1.vel[:] = np.random.default_rng(1).uniform(-.3, .3, S1.vel.shape) + [0, 0, -.5]
    return {
        "cfg1": S1,
        "cfg2": sc.deposition(1200, r=1e-3, psd_ratio=4.0, seed=2, layers=8, spacing_factor=1.6),
        "cfg3": sc.hertz_deposition(1200, spacing_factor=2.05, vz0=.3, vfluct=.5, aspect=(4, 1, 1)),
        "cfg4": sc.periodic_shear(1500, vfluct=.05, solid=.6, rate=50.),
        "cfg5": sc.drum(800, div=16, axial=4, vz0=1., vfluct=.5),
        "gas-linear": sc.random_gas(1500, box=.75, r=.02, psd_ratio=3.0, seed=1),
        "gas-hertz": sc.random_gas(800, box=.6, r=.02, psd_ratio=2.0, seed=2, law=sc.LAW_HERTZ_DMT, en=.7),
        "facets": sc.facet_box_scene(1000, seed=33),
        "peri-skewed": sc.periodic_gas(1200, seed=34, gradV=[[0, 1., 0], [0, 0, 0], [0, 0, 0]], shear0=.1),
    }


def one_step_agreement(a, b, ca, cb, label, rmax):
    assert [tuple(x) for x in ca["ids"]] == [tuple(x) for x in cb["ids"]], f"{label}: the real-contact set depends on the reduction order"
    scale = max(np.abs(a["force"]).max(), np.abs(ca["force"]).max() if len(ca["force"]) else 0., 1e-300)
    tscale = max(np.abs(a["torque"]).max(), scale * rmax)     # purely normal forces give torques that are round-off of |F|*r: scale by that
    rel = dict(force=np.abs(a["force"] - b["force"]).max() / scale, torque=np.abs(a["torque"] - b["torque"]).max() / tscale,
               contact_force=(np.abs(ca["force"] - cb["force"]).max() / scale) if len(ca["force"]) else 0.)
    for k, v in rel.items():
        assert v <= REL_FORCE_TOL, f"{label}: {k} moves by {v:.2e} relative under the other reduction order (tolerance {REL_FORCE_TOL})"
    return rel


@pytest.mark.parametrize("name", ["gas-linear", "gas-hertz", "facets", "peri-skewed", "cfg4", "cfg2"])
def test_oracle_one_step_from_identical_state(name):
    S = cfg_scenes()[name]
    A, B = Oracle(S), Oracle(S, variant="alt")
    for k in range(3):           # steps 1..3 from the same initial state (contacts are created in step 1, updated with history afterwards)
        A.step(1); B.step(1)
        ca = A.contacts(True)
        rel = one_step_agreement(A.spheres(), B.spheres(), ca, B.contacts(True), f"{name} step {k + 1}", S.radius.max())
    assert len(ca["ids"]) > 50, "scene has too few contacts to say anything"
    assert A.nFullRuns == B.nFullRuns


def test_oracle_long_run_statistics_small_system():
    """Thousands of steps: the two orders follow different (ulp-perturbed, chaotically diverging) trajectories of the same packing.  At the
    1500-sphere size the oracle can afford, coordination number and late kinetic energy fluctuate by a few per cent between ANY two such
    trajectories, so only the robust observables get the 1 % bound here; the 1 % criterion proper is checked at 60k-100k spheres on the GPU
    (test_gpu_long_run_observables_between_orders)."""
    S = cfg_scenes()["cfg1"]
    A, B = Oracle(S), Oracle(S, variant="alt")
    A.step(1000); B.step(1000)       # spheres are still raining down: kinetic energy is a bulk quantity
    assert B.kineticEnergy() == pytest.approx(A.kineticEnergy(), rel=1e-2)
    A.step(2000); B.step(2000)
    sa, sb = A.spheres(), B.spheres()
    assert not np.array_equal(sa["pos"], sb["pos"]), "the two orders should have diverged at ulp level by now"
    assert sb["pos"][:, 2].mean() == pytest.approx(sa["pos"][:, 2].mean(), rel=1e-2)          # potential energy of the bed
    za, zb = (ob.coord_number(S.sphere_ids(), s["pos"], O.contacts(True)["ids"]) for s, O in ((sa, A), (sb, B)))
    assert za > 3. and zb == pytest.approx(za, rel=5e-2)


# ---------------------------------------------------------------------------------------------------------------------------------
# GPU part
# ---------------------------------------------------------------------------------------------------------------------------------
def _resume(S, A, variant):
    """A fresh engine (default or alt library) holding exactly the state of engine A: spheres, forces, contact history, scene counters."""
    from woo_b200.engine import Engine
    s, c, sca = A.spheres(), A.contacts(True), A.scalars()
    B = Engine(S.params, variant=variant)
    B.spec = S
    B.upload_materials(S.mats)
    m, I = S.mass_inertia()
    B.upload_spheres(S.sphere_ids(), s["pos"], S.radius, m, I, S.mat, S.mask, S.flags, s["vel"], s["angVel"], ori=s["ori"])
    B._chk(B.L.wdem_upload_forces(B.h, s["force"].ctypes.data_as(C.c_void_p), s["torque"].ctypes.data_as(C.c_void_p)))
    B.upload_walls(S.wall_ids(), S.walls); B.upload_facets(S.facet_ids(), S.facets)
    if S.cell is not None:
        B.set_cell(A.cell_hsize(), S.cell.get("gradV"), S.cell.get("nextGradV"), S.cell.get("homoDeform", 4))
    B.set_config(dt=sca["dt"])
    B.set_scene(sca["step"], sca["time"])
    B.upload_contacts(c)
    return B


@pytest.mark.gpu
@pytest.mark.parametrize("name,warm", [("cfg1", 1500), ("cfg2-bed20k", 200), ("cfg3", 1500), ("cfg4", 300), ("cfg5", 2500), ("gas-linear", 60), ("gas-hertz", 60), ("facets", 60)])
def test_gpu_one_step_from_identical_developed_state(name, warm):
    from woo_b200.engine import Engine
    if name == "cfg2-bed20k":      # the committed pre-settled PSD 1:4 bed that bench.py tiles to 1M: ~40k real contacts with long shear histories
        import bench
        S = bench.load_sample()
    else:
        S = cfg_scenes()[name]
    S.params.dynDtPeriod = 0          # PeriodicEngine counters are not part of the exchanged state
    A = Engine(S.params).load(S)
    A.step(warm)
    assert len(A.contacts(True)["ids"]) > 50, "no developed contact network to test on"
    D, X = _resume(S, A, ""), _resume(S, A, "alt")
    D.step(1); X.step(1); A.step(1)
    sd, cd = D.spheres(), D.contacts(True)
    # the resumed default engine reproduces the uninterrupted one bit for bit: the comparison below really starts from identical state.  (Not asserted for
    # periodic cells: only REAL contacts travel through wdem_upload_contacts, and a potential contact that survived many collider runs of a sheared cell keeps the
    # cellDist it was created with, which a fresh candidate search need not reproduce.)
    if S.cell is None:
        assert np.array_equal(sd["force"], A.spheres()["force"]) and np.array_equal(cd["force"], A.contacts(True)["force"])
    one_step_agreement(sd, X.spheres(), cd, X.contacts(True), f"{name} (GPU, developed state)", S.radius.max())


@pytest.mark.gpu
@pytest.mark.parametrize("name,steps", [("gas-linear", 100), ("gas-hertz", 100), ("facets", 100), ("peri-skewed", 150), ("cfg4", 200)])
def test_gpu_alt_equals_oracle_alt_bitwise(name, steps):
    from woo_b200.engine import Engine
    S = cfg_scenes()[name]
    O = Oracle(S, variant="alt"); E = Engine(S.params, variant="alt").load(S)
    O.step(steps); E.step(steps)
    so, sg, co, cg = O.spheres(), E.spheres(), O.contacts(True), E.contacts(True)
    assert [tuple(x) for x in co["ids"]] == [tuple(x) for x in cg["ids"]]
    for k in ("pos", "vel", "angVel", "force", "torque"):
        assert np.array_equal(so[k], sg[k]), f"{name}: {k} (alt order) differs between the CUDA path and the oracle"
    for k in ("uN", "force", "trsf", "contPt"):
        assert np.array_equal(co[k], cg[k]), f"{name}: contact {k} (alt order) differs"
    # and the alt order is really a different arithmetic: it differs from the default oracle somewhere at ulp level
    D = Oracle(S); D.step(steps)
    assert not np.array_equal(D.spheres()["force"], so["force"])


def _bed_observables(S, E):
    s, c = E.spheres(), E.contacts(True)
    z = ob.coord_number(S.sphere_ids(), s["pos"], c["ids"])
    if S.cell is not None:
        return dict(z=z)
    # DemFuncs::porosity (Funcs.cpp:104-113) over a fixed inner box: away from the side walls, lower 60 % of the initial lattice height
    lo = np.array([w["pos"][w["axis"]] for w in S.walls if w["sense"] == 1][:3]); hi = np.array([w["pos"][w["axis"]] for w in S.walls if w["sense"] == -1][:2] + [0.])
    rm = S.radius.max(); zq = np.sort(s["pos"][:, 2])[int(.6 * S.n)]
    box = (lo + 3 * rm, np.array([hi[0] - 3 * rm, hi[1] - 3 * rm, zq]))
    return dict(z=z, phi=ob.packing_fraction(s["pos"], S.radius, (box[0], box[1])), height=float(s["pos"][:, 2].mean()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg4"])
def test_gpu_long_run_observables_between_orders(name):
    """north_star: "multi-thousand-step runs must match the reference's kinetic energy, coordination number and packing fraction within 1 %"
    — between the two candidate Eigen orders, at a size where these are bulk quantities."""
    from woo_b200.engine import Engine
    if name == "cfg1":
        S = sc.deposition(60000, seed=1, spacing_factor=2.05, layers=12); S.vel[:] = np.random.default_rng(1).uniform(-.3, .3, S.vel.shape) + [0, 0, -.5]
        flow, total = 1500, 12000
    elif name == "cfg3":
        S = sc.hertz_deposition(60000, spacing_factor=2.05, vz0=.3, vfluct=.5, aspect=(4, 1, .25)); flow, total = 1500, 12000
    else:
        S = sc.periodic_shear(60000, vfluct=.05, solid=.6, rate=50.); flow, total = 500, 3000
    A = Engine(S.params).load(S); B = Engine(S.params, variant="alt").load(S)
    A.step(flow); B.step(flow)
    eka, ekb = A.kinetic_energy(), B.kinetic_energy()
    A.step(total - flow); B.step(total - flow)
    oa, ob_ = _bed_observables(S, A), _bed_observables(S, B)
    print(f"{name}: Ek(flow) {eka:.6g} vs {ekb:.6g}; end {oa} vs {ob_}")
    assert not np.array_equal(A.spheres()["pos"], B.spheres()["pos"])
    assert ekb == pytest.approx(eka, rel=1e-2)
    assert oa["z"] > 2.5
    for k in oa:
        assert ob_[k] == pytest.approx(oa[k], rel=1e-2), f"{name}: {k} {oa[k]:.5g} vs {ob_[k]:.5g}"
