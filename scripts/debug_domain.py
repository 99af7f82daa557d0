"""Step a LocalCluster and a single-domain Engine side by side; report the first divergence (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from woo_b200 import scenes as sc
from woo_b200.engine import Engine
from woo_b200.domain import LocalCluster, owner_of
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = sc.random_gas(500, box=.5, r=.02, psd_ratio=3.0, seed=21, law=0)
E = Engine(S.params).load(S); K = LocalCluster(S, world)
print("cuts", K.cuts)
for step in range(1, 60):
    E.step(1); K.step(1)
    s1, s2 = E.spheres(), K.spheres()
    p1, p2 = E.contacts(False), K.contacts(False)
    r1, r2 = E.contacts(True), K.contacts(True)
    P1, P2 = set(map(tuple, p1["ids"])), set(map(tuple, p2["ids"]))
    R1, R2 = set(map(tuple, r1["ids"])), set(map(tuple, r2["ids"]))
    bad = {k: int(np.any(s1[k] != s2[k], axis=1).sum()) for k in ("pos", "vel", "angVel", "force", "torque")}
    print(step, "fullruns", E.scalars()["nFullRuns"], K.scalars()["nFullRuns"], "pot", len(P1), len(P2), "real", len(R1), len(R2), "dup", len(p2["ids"]) - len(P2), "bad", bad, K.domain_stats())
    if P1 != P2 or R1 != R2 or any(bad.values()):
        own = owner_of(S, K.cuts, 0)
        print(" pot only single:", sorted(P1 - P2)[:10]); print(" pot only cluster:", sorted(P2 - P1)[:10])
        print(" real only single:", sorted(R1 - R2)[:10]); print(" real only cluster:", sorted(R2 - R1)[:10])
        for (a, b) in (sorted(P1 ^ P2) + sorted(R1 ^ R2))[:6]:
            xa = s1["pos"][a, 0] if a < S.n else None; xb = s1["pos"][b, 0] if b < S.n else None
            print("  pair", a, b, "x", xa, xb, "r", S.radius[a] if a < S.n else None, S.radius[b] if b < S.n else None)
        ids = np.nonzero(np.any(s1["force"] != s2["force"], axis=1))[0][:5]
        for i in ids:
            print("  sphere", i, "x", s1["pos"][i, 0], "F single", s1["force"][i], "cluster", s2["force"][i])
            for (x, k) in [(tuple(x), k) for k, x in enumerate(r1["ids"]) if i in x]:
                k2 = [tuple(y) for y in r2["ids"]].index(x) if x in R2 else None
                if k2 is None: print("   missing in cluster", x); continue
                for f in ("uN", "force", "contPt", "trsf", "phys"):
                    if not np.array_equal(r1[f][k], r2[f][k2]): print("   contact", x, f, r1[f][k], r2[f][k2])
        break
