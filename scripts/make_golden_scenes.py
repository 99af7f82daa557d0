"""Golden end states of small seeded scenes, produced by the CPU oracle (oracle/dem_oracle.cpp) at the commit that introduced them.  They pin the oracle
against silent drift between rounds (the reference itself cannot run here) and give the GPU tests a fixture that does not depend on the
oracle being rebuilt identically.   python scripts/make_golden_scenes.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from woo_b200 import scenes as sc
from _harness import Oracle


FIRST_STEPS = (1, 5)     # tests/golden/first_<scene>.npz: states right after the first steps from the initial state (north_star's one-step criterion; scripts/replay_in_woo.py)


def golden_scenes():
    """name -> (SceneSpec factory, steps); shared with tests/test_golden_scenes.py"""
    return {
        "gas_linear_psd3": (lambda: sc.random_gas(300, box=.45, r=.02, psd_ratio=3.0, seed=101), 250),
        "gas_hertz": (lambda: sc.random_gas(300, box=.45, r=.02, psd_ratio=2.0, seed=102, law=sc.LAW_HERTZ_DMT, en=.7), 250),
        "deposition": (lambda: sc.deposition(1500, seed=103), 2500),
        "facet_box": (lambda: sc.facet_box_scene(250, seed=104, halfThick=.003), 600),
        "rotating_drum": (lambda: sc.drum(400, div=12, axial=3, vz0=.8, vfluct=.4, omega=7., rotate=True), 4000),
        "periodic_shear": (lambda: sc.periodic_gas(300, seed=105, gradV=[[0, 1., 0], [0, 0, 0], [0, 0, 0]], shear0=.1, psd_ratio=2.0), 500),
    }


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    for name, (make, steps) in golden_scenes().items():
        S = make(); O = Oracle(S)
        first = {}; done = 0
        for k in FIRST_STEPS:
            O.step(k - done); done = k
            s, c = O.spheres(), O.contacts(True)
            first.update({"pos_%d" % k: s["pos"], "vel_%d" % k: s["vel"], "angVel_%d" % k: s["angVel"], "force_%d" % k: s["force"], "torque_%d" % k: s["torque"],
                          "ids_%d" % k: c["ids"], "uN_%d" % k: c["uN"], "cforce_%d" % k: c["force"], "dt_%d" % k: O.dt})
        np.savez_compressed(os.path.join(out, "first_%s.npz" % name), **first)
        O.step(steps - done)
        s, c = O.spheres(), O.contacts(True)
        np.savez_compressed(os.path.join(out, "scene_%s.npz" % name), steps=steps, pos=s["pos"], vel=s["vel"], angVel=s["angVel"], force=s["force"], torque=s["torque"],
                            ids=c["ids"], uN=c["uN"], cforce=c["force"], trsf=c["trsf"], cellDist=c["cellDist"], dt=O.dt, nFullRuns=O.nFullRuns)
        print(name, steps, "steps,", len(c["ids"]), "real contacts, nFullRuns", O.nFullRuns)
