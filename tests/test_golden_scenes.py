"""Committed end states of small seeded scenes (tests/golden/scene_*.npz, scripts/make_golden_scenes.py): the oracle must still reproduce them bit for bit
(CPU), and so must the CUDA path through the C-ABI (GPU)."""
import os, sys
import numpy as np
import pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from make_golden_scenes import golden_scenes, FIRST_STEPS
from _harness import Oracle

SCENES = golden_scenes()


def check(E, name, contacts, spheres, nfull, dt):
    g = np.load(os.path.join(ROOT, "tests", "golden", "scene_%s.npz" % name))
    E.step(int(g["steps"]))
    s, c = spheres(E), contacts(E)
    assert [tuple(x) for x in c["ids"]] == [tuple(x) for x in g["ids"]], "real contact set differs from the golden state"
    for k in ("pos", "vel", "angVel", "force", "torque"):
        assert np.array_equal(s[k], g[k]), k
    for k, gk in (("uN", "uN"), ("force", "cforce"), ("trsf", "trsf"), ("cellDist", "cellDist")):
        assert np.array_equal(c[k], g[gk]), "contact " + k
    assert nfull(E) == int(g["nFullRuns"]) and dt(E) == float(g["dt"])


def check_first(E, name, contacts, spheres, dt):
    """tests/golden/first_<scene>.npz: the states right after the first steps from the initial state (the fixtures scripts/replay_in_woo.py holds a real Woo build against)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "first_%s.npz" % name))
    done = 0
    for k in FIRST_STEPS:
        E.step(k - done); done = k
        s, c = spheres(E), contacts(E)
        assert [tuple(x) for x in c["ids"]] == [tuple(x) for x in g["ids_%d" % k]], "real contact set differs after %d steps" % k
        for q in ("pos", "vel", "angVel", "force", "torque"):
            assert np.array_equal(s[q], g["%s_%d" % (q, k)]), "%s after %d steps" % (q, k)
        assert np.array_equal(c["uN"], g["uN_%d" % k]) and np.array_equal(c["force"], g["cforce_%d" % k]) and dt(E) == float(g["dt_%d" % k])


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_reproduces_first_steps(name):
    check_first(Oracle(SCENES[name][0]()), name, lambda O: O.contacts(True), lambda O: O.spheres(), lambda O: O.dt)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_gpu_reproduces_first_steps(name):
    from woo_b200.engine import Engine
    S = SCENES[name][0]()
    check_first(Engine(S.params).load(S), name, lambda E: E.contacts(True), lambda E: E.spheres(), lambda E: E.scalars()["dt"])


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_reproduces_golden(name):
    check(Oracle(SCENES[name][0]()), name, lambda O: O.contacts(True), lambda O: O.spheres(), lambda O: O.nFullRuns, lambda O: O.dt)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_gpu_reproduces_golden(name):
    from woo_b200.engine import Engine
    S = SCENES[name][0]()
    check(Engine(S.params).load(S), name, lambda E: E.contacts(True), lambda E: E.spheres(), lambda E: E.scalars()["nFullRuns"], lambda E: E.scalars()["dt"])
