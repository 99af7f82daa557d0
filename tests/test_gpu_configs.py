"""BASELINE.json configs 1-5 as parity cases at sizes the oracle finishes in seconds (bench.py times configs[1] at full size), single-domain and
slab-decomposed, plus the long-run observables of SURVEY §8c (kinetic energy, coordination number, packing fraction; the 1 % criterion)."""
import numpy as np
import pytest
from woo_b200 import scenes as sc
from woo_b200 import observables as ob
from _harness import Oracle
from test_gpu_parity import compare, cluster, compare_potential

pytestmark = pytest.mark.gpu


def test_cfg1_deposition_long_run_observables():
    # cfg1 in small: monodisperse deposition in a box of 5 walls, linear law, DynDt(100); 6000 steps
    S = sc.deposition(1500, seed=1, spacing_factor=2.05)
    S.vel[:] = np.random.default_rng(1).uniform(-.3, .3, S.vel.shape) + [0, 0, -.5]
    O, E = compare(S, 6000, check_every=2000, label="cfg1 long run")
    so, sg = O.spheres(), E.spheres()
    co, cg = O.contacts(True), E.contacts(True)
    lo = np.zeros(3); hi = np.array([w["pos"][w["axis"]] for w in S.walls if w["sense"] == -1] + [so["pos"][:, 2].max() + S.radius[0]])
    ek_o, ek_g = O.kineticEnergy(), E.kinetic_energy()
    z_o, z_g = ob.coord_number(S.sphere_ids(), so["pos"], co["ids"]), ob.coord_number(S.sphere_ids(), sg["pos"], cg["ids"])
    f_o, f_g = ob.packing_fraction(so["pos"], S.radius, (lo, hi)), ob.packing_fraction(sg["pos"], S.radius, (lo, hi))
    assert len(co["ids"]) > 2 * S.n and z_o > 3.5, "the bed did not form (coordination %.2f)" % z_o
    # north_star tolerance for multi-thousand-step runs: 1 % (here the trajectories are in fact bit-identical, so these are equalities)
    assert ek_g == pytest.approx(ek_o, rel=1e-2) and z_g == pytest.approx(z_o, rel=1e-2) and f_g == pytest.approx(f_o, rel=1e-2)
    assert 0.3 < f_o < 0.75


def test_cfg1_at_its_named_size_10k_spheres_20k_steps():
    """BASELINE.json configs[0] as named: 10 000 spheres, 20 000 steps of Leapfrog + collider + ContactLoop(linear) + DynDt — the one configuration the reference itself is
    timed on (obsolete/perf/inclined.py:16-49).  About a minute of oracle on the box's host cores; the whole trajectory is bit-identical."""
    S = sc.deposition(10000, seed=1)          # the scene bench.py's configs block times as cfg1
    assert S.n == 10000
    O, E = compare(S, 20000, check_every=5000, label="cfg1 10k x 20k")
    co = O.contacts(True)
    assert len(co["ids"]) > 2 * S.n and O.nFullRuns > 100
    so, sg = O.spheres(), E.spheres()
    assert ob.coord_number(S.sphere_ids(), so["pos"], co["ids"]) == ob.coord_number(S.sphere_ids(), sg["pos"], E.contacts(True)["ids"])


@pytest.mark.parametrize("world", [1, 3])
def test_cfg2_psd_bed(world):
    # the committed pre-settled 20k-sphere PSD 1:4 bed that bench.py tiles to 1M (tests/golden/cfg2_sample_20k.npz): dense, ~4 contacts per sphere
    import bench
    S = bench.load_sample()
    assert S is not None and S.n == 20000
    O, E = compare(S, 300, check_every=100, label=f"cfg2 bed world={world}", make=cluster(world) if world > 1 else None)
    assert len(O.contacts(True)["ids"]) > 1.8 * S.n


@pytest.mark.parametrize("world", [1, 2])
def test_cfg3_hertz_deposition_slabs(world):
    S = sc.hertz_deposition(1200, spacing_factor=2.05, vz0=.3, vfluct=.5, aspect=(4, 1, 1))
    O, E = compare(S, 4000, check_every=1000, label=f"cfg3 world={world}", make=cluster(world) if world > 1 else None)
    assert len(O.contacts(True)["ids"]) > .8 * S.n


@pytest.mark.parametrize("world", [1, 4])
def test_cfg4_periodic_simple_shear(world):
    S = sc.periodic_shear(1500, vfluct=.05, solid=.6, rate=50.)
    O, E = compare(S, 600, check_every=200, label=f"cfg4 world={world}", make=cluster(world) if world > 1 else None)
    compare_potential(O, E)
    h = E.cell_hsize()
    assert h[0, 1] > 0 and np.array_equal(h, O.cell_hsize())          # the cell is being sheared
    assert len(O.contacts(True)["ids"]) > 2 * S.n


@pytest.mark.parametrize("world", [1, 2])
def test_cfg5_drum_facets_fakevel(world):
    S = sc.drum(800, div=16, axial=4, vz0=1., vfluct=.5)
    O, E = compare(S, 5000, check_every=1000, label=f"cfg5 world={world}", make=cluster(world) if world > 1 else None)
    ids = O.contacts(True)["ids"]
    assert (ids[:, 0] >= S.n + len(S.walls)).sum() > 50, "no facet contacts formed"
