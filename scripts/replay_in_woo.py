#!/usr/bin/env python
"""Replay the committed golden scenes in a REAL Woo build and diff against the golden states — the step that turns "parity unpinned" into "pinned".

It cannot run in the container this repository was developed in (Woo needs Eigen/Boost/TBB to build; none is installed).  On any machine
with Woo (`pip install woo` / a distro package / a source build; `import woo` must work) and numpy:

    python scripts/replay_in_woo.py                 # all scenes of tests/golden/, report on stdout, JSON in replay_in_woo_report.json
    python scripts/replay_in_woo.py gas_hertz       # one scene
    python scripts/replay_in_woo.py --dump out/     # also save Woo's own end states as npz (to be committed as reference-made fixtures)

What is compared (north_star's two levels):
  * after the FIRST steps from the identical initial state (tests/golden/first_<scene>.npz: states after 1 and after 5 steps): the set of real contacts must be
    identical, per-particle force/torque and per-contact force within 1e-10 relative; "bitwise" is reported too: it tells whether the Eigen reduction orders
    restated in oracle/dem_oracle.cpp (WD_SUM3_ORDER / WD_TRACE_ORDER / WD_MATPROD_ORDER) are the ones this Woo build uses;
  * at the END of the run (tests/golden/scene_<scene>.npz, hundreds to thousands of steps): bitwise if the orders match, else kinetic energy, coordination number
    and mean height within 1 %.
The scenes are rebuilt from woo_b200/scenes.py (pure numpy, no GPU needed) by scripts/make_golden_scenes.golden_scenes(); ids follow S.dem.par.add order
(spheres, then walls, then facets), which is also what the oracle and the CUDA engine assume.  S.deterministic=True makes the reference apply forces in
creation order; per-particle sums may then differ from the engine's ascending-partner-id order in the last bits — reported as max relative difference.
"""
import json, math, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_woo_scene(spec):
    """SceneSpec -> woo.core.Scene with the reference's own engines for the path (py/utils.py:94-129 defaultEngines, restricted to sphere/wall/facet functors)."""
    import woo, woo.core, woo.dem, woo.utils
    from minieigen import Vector3, Matrix3, Quaternion
    P = spec.params
    S = woo.core.Scene(fields=[woo.dem.DemField(gravity=Vector3(*P.gravity), loneMask=int(P.loneMask), distFactor=float(P.distFactor))])
    S.deterministic = True
    S.dtSafety = float(P.dtSafety)
    if not math.isnan(P.dt):
        S.dt = float(P.dt)
    mats = []
    for kind, density, young, ktDivKn, tanPhi, surfEnergy, alpha in spec.mats:
        if kind == 0:
            mats.append(woo.dem.FrictMat(density=density, young=young, ktDivKn=ktDivKn, tanPhi=tanPhi))
        else:
            mats.append(woo.dem.HertzMat(density=density, young=young, ktDivKn=ktDivKn, tanPhi=tanPhi, surfEnergy=surfEnergy, alpha=alpha))
    if spec.cell is not None:
        S.periodic = True
        H = np.asarray(spec.cell["hSize"], dtype=float).reshape(3, 3)
        S.cell.setBox(Vector3(1, 1, 1)); S.cell.hSize = Matrix3(*H.reshape(9).tolist())
        S.cell.homoDeform = {0: "none", 1: "pos", 2: "vel", 3: "vel2", 4: "gradV2"}[int(spec.cell.get("homoDeform", 4))]
        g = np.asarray(spec.cell.get("gradV", np.zeros((3, 3))), dtype=float).reshape(9)
        ng = np.asarray(spec.cell.get("nextGradV", g), dtype=float).reshape(9)
        S.cell.gradV = Matrix3(*g.tolist()); S.cell.nextGradV = Matrix3(*ng.tolist())

    def spheres():
        for i in range(spec.n):
            p = woo.utils.sphere(Vector3(*spec.pos[i]), float(spec.radius[i]), mat=mats[int(spec.mat[i])], mask=int(spec.mask[i]))
            p.vel = Vector3(*spec.vel[i]); p.angVel = Vector3(*spec.angVel[i])
            fl = int(spec.flags[i])
            if fl & 63:
                p.blocked = "".join(c for b, c in zip((1, 2, 4, 8, 16, 32), "xyzXYZ") if fl & b)
            p.shape.nodes[0].dem.gravitySkip = bool(fl & 512); p.shape.nodes[0].dem.dampingSkip = bool(fl & 2048)
            S.dem.par.add(p)

    def boundaries():
        for w in spec.walls:
            b = woo.utils.wall(Vector3(*w["pos"]), axis=int(w["axis"]), sense=int(w["sense"]), mat=mats[int(w.get("mat", 0))], mask=int(w.get("mask", 3)))
            if any(w.get("vel", (0, 0, 0))):
                b.vel = Vector3(*w["vel"])              # a moving wall enters DemField.nodes (Particle::guessMoving)
            S.dem.par.add(b)
        facets = []
        for f in spec.facets:
            v = np.asarray(f["v"], dtype=float).reshape(3, 3)
            fk = f.get("fakeVel", (0, 0, 0))
            b = woo.utils.facet([Vector3(*v[0]), Vector3(*v[1]), Vector3(*v[2])], fakeVel=Vector3(*fk), halfThick=float(f.get("halfThick", 0.)),
                                mat=mats[int(f.get("mat", 0))], mask=int(f.get("mask", 3)))
            nl = f.get("n21lim")
            if nl is not None and not all(math.isnan(x) for x in nl):
                b.shape.n21lim = Vector3(*[float(x) for x in nl])
            vv = np.asarray(f.get("vvel", np.zeros(9)), dtype=float).reshape(3, 3)
            for k in range(3):
                if vv[k].any():
                    b.shape.nodes[k].dem.vel = Vector3(*vv[k])
            facets.append(b)
        grouped = set()
        for g in getattr(spec, "facet_groups", []):     # examples/mill.py:38-42
            node = woo.core.Node(pos=Vector3(*g["pos"]), ori=Quaternion((g["ori"][3], g["ori"][0], g["ori"][1], g["ori"][2])))
            S.dem.par.addClumped([facets[i] for i in g["facets"]], centralNode=node)
            node.dem.vel = Vector3(*g["vel"]); node.dem.angVel = Vector3(*g["angVel"])
            grouped |= set(g["facets"])
        for i, b in enumerate(facets):
            if i not in grouped:
                S.dem.par.add(b)

    if spec.boundaries_first:
        boundaries(); spheres()
    else:
        spheres(); boundaries()
    if P.law == 0:
        cp2 = woo.dem.Cp2_FrictMat_FrictPhys()
        law = woo.dem.Law2_L6Geom_FrictPhys_IdealElPl(noSlip=bool(P.noSlip), noBreak=bool(P.noBreak), noFrict=bool(P.noFrict), relRollStiff=P.relRollStiff,
                                                      relTwistStiff=P.relTwistStiff, rollTanPhi=P.rollTanPhi)
    else:
        cp2 = woo.dem.Cp2_HertzMat_HertzPhys(poisson=P.poisson, en=P.en)
        law = woo.dem.Law2_L6Geom_HertzPhys_DMT(noAttraction=bool(P.noAttraction))
    cg = [woo.dem.Cg2_Sphere_Sphere_L6Geom(), woo.dem.Cg2_Facet_Sphere_L6Geom(), woo.dem.Cg2_Wall_Sphere_L6Geom()]
    for c in cg:
        c.noRatch = bool(P.noRatch); c.trsfRenorm = int(P.trsfRenorm); c.iniLensTouch = bool(P.iniLensTouch)
    S.engines = [woo.dem.Leapfrog(damping=P.damping, reset=bool(P.reset)),
                 woo.dem.InsertionSortCollider([woo.dem.Bo1_Sphere_Aabb(), woo.dem.Bo1_Facet_Aabb(), woo.dem.Bo1_Wall_Aabb(halfThick=P.wallHalfThick)], verletDist=P.verletDist, label="collider"),
                 woo.dem.ContactLoop(cg, [cp2], [law], applyForces=bool(P.applyForces))] + \
                ([woo.dem.DynDt(stepPeriod=int(P.dynDtPeriod), maxRelInc=P.maxRelInc)] if P.dynDtPeriod > 0 else [])
    return S


def woo_state(S, spec):
    """The arrays of a golden file, read back from the Woo scene (spheres by id; real contacts sorted by (id1,id2) with boundary first / lower id first)."""
    import woo, woo.dem
    ids = spec.sphere_ids()
    g = lambda f: np.array([list(f(S.dem.par[int(i)])) for i in ids], dtype=float)
    out = dict(pos=g(lambda p: p.pos), vel=g(lambda p: p.vel), angVel=g(lambda p: p.angVel), force=g(lambda p: p.f), torque=g(lambda p: p.t))
    rows = []
    for c in S.dem.con:
        if not c.real:
            continue
        a, b = c.id1, c.id2
        f = np.array(list(c.phys.force), dtype=float); t = np.array(c.geom.trsf).reshape(3, 3)
        # canonical orientation of the fixtures: boundary particle first, else the lower id first.  The reference already orders wall/facet before sphere
        # (functor order); a sphere-sphere contact created as (hi,lo) describes the same contact seen from the other side: normal and force flip sign.
        if isinstance(S.dem.par[a].shape, woo.dem.Sphere) and isinstance(S.dem.par[b].shape, woo.dem.Sphere) and a > b:
            rows.append(((b, a), c.geom.uN, None, None))      # frame-dependent data of flipped contacts is not compared
        else:
            rows.append(((a, b), c.geom.uN, f, t.T.reshape(9)))          # trsf: columns (local axes) stored one after another, as in the fixtures
    rows.sort(key=lambda r: r[0])
    out["ids"] = np.array([r[0] for r in rows], dtype=np.int32).reshape(-1, 2)
    out["uN"] = np.array([r[1] for r in rows], dtype=float)
    out["cforce"] = np.array([r[2] if r[2] is not None else [np.nan] * 3 for r in rows], dtype=float).reshape(-1, 3)
    out["dt"] = S.dt
    return out


def coordination(n, ids):
    cnt = np.bincount(np.asarray(ids).reshape(-1), minlength=n + 1)[:n]
    N0, N1 = int((cnt == 0).sum()), int((cnt == 1).sum())
    return (cnt.sum() - N1) / max(n - N0 - N1, 1)


def compare(tag, got, gold, n):
    rep = {"real_contacts": [int(len(got["ids"])), int(len(gold["ids"]))]}
    rep["contact_set_identical"] = bool(got["ids"].shape == gold["ids"].shape and np.array_equal(got["ids"], gold["ids"]))
    scale = max(np.abs(gold["force"]).max(), 1e-300)
    for k in ("pos", "vel", "angVel", "force", "torque"):
        d = np.abs(got[k] - gold[k]).max() if len(gold[k]) else 0.
        ref = max(np.abs(gold[k]).max(), 1e-300)
        rep[k] = {"bitwise": bool(np.array_equal(got[k], gold[k])), "max_rel": float(d / ref)}
    if rep["contact_set_identical"] and len(gold["ids"]):
        ok = ~np.isnan(got["cforce"]).any(axis=1)
        rep["contact_force_max_rel"] = float(np.abs(got["cforce"][ok] - gold["cforce"][ok]).max() / scale) if ok.any() else None
        rep["uN_bitwise"] = bool(np.array_equal(got["uN"], gold["uN"]))
    rep["dt_bitwise"] = bool(got["dt"] == float(gold["dt"]))
    rep["within_1e-10"] = bool(rep["contact_set_identical"] and rep["force"]["max_rel"] <= 1e-10 and rep["torque"]["max_rel"] <= 1e-10 and (rep.get("contact_force_max_rel") or 0.) <= 1e-10)
    ek = lambda s: float((s["vel"] ** 2).sum())
    rep["observables"] = {"sum_v2": [ek(got), ek(gold)], "coordination": [float(coordination(n, got["ids"][got["ids"] < n])), float(coordination(n, gold["ids"][gold["ids"] < n]))],
                          "mean_z": [float(got["pos"][:, 2].mean()), float(gold["pos"][:, 2].mean())]}
    print("%-34s contacts %s %s | force rel %.2e %s | pos %s" % (tag, rep["real_contacts"], "IDENTICAL" if rep["contact_set_identical"] else "DIFFER", rep["force"]["max_rel"],
                                                              "<=1e-10" if rep["within_1e-10"] else "", "bitwise" if rep["pos"]["bitwise"] else "max rel %.2e" % rep["pos"]["max_rel"]))
    return rep


def main():
    try:
        import woo   # noqa: F401
    except ImportError:
        print("This script needs a real Woo installation (import woo failed).  See the docstring.", file=sys.stderr)
        return 2
    from make_golden_scenes import golden_scenes, FIRST_STEPS
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    report = {}
    gdir = os.path.join(ROOT, "tests", "golden")
    for name, (make, steps) in golden_scenes().items():
        if args and name not in args:
            continue
        spec = make()
        S = build_woo_scene(spec)
        first = np.load(os.path.join(gdir, "first_%s.npz" % name))
        rep = {}; done = 0
        for k in FIRST_STEPS:
            S.run(k - done, True); done = k
            got = woo_state(S, spec)
            rep["after_%d_steps" % k] = compare("%s after %d steps" % (name, k), got, {q[:-len("_%d" % k)]: first[q] for q in first.files if q.endswith("_%d" % k)}, spec.n)
        S.run(steps - done, True)
        got = woo_state(S, spec)
        gold = np.load(os.path.join(gdir, "scene_%s.npz" % name))
        rep["end"] = compare("%s end (%d steps)" % (name, steps), got, gold, spec.n)
        rep["end"]["nFullRuns"] = [int(S.lab.collider.nFullRuns), int(gold["nFullRuns"])]
        report[name] = rep
        if dump:
            os.makedirs(dump, exist_ok=True)
            np.savez_compressed(os.path.join(dump, "woo_scene_%s.npz" % name), steps=steps, **got)
    json.dump(report, open("replay_in_woo_report.json", "w"), indent=1)
    pinned = all(r["after_%d_steps" % FIRST_STEPS[0]]["within_1e-10"] for r in report.values())
    print("PARITY PINNED (contacts identical, forces within 1e-10 after the first step)" if pinned else "PARITY NOT REACHED: see replay_in_woo_report.json")
    return 0 if pinned else 1


if __name__ == "__main__":
    sys.exit(main())
