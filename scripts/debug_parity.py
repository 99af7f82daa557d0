import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from woo_b200 import scenes as sc
from woo_b200.engine import Engine
from _harness import Oracle
S = sc.random_gas(400, box=.5, r=.02, psd_ratio=4.0, seed=3, law=0)
O = Oracle(S); E = Engine(S.params).load(S)
for step in range(1, 40):
    O.step(1); E.step(1)
    co, cg = O.contacts(False), E.contacts(False)
    so_, sg_ = set(map(tuple, co["ids"])), set(map(tuple, cg["ids"]))
    ro, rg = O.contacts(True), E.contacts(True)
    sro, srg = set(map(tuple, ro["ids"])), set(map(tuple, rg["ids"]))
    so, sg = O.spheres(), E.spheres()
    print(step, "pot O/G", len(so_), len(sg_), "real O/G", len(sro), len(srg), "fullruns", O.nFullRuns, E.scalars()["nFullRuns"], "posdiff", np.abs(so["pos"] - sg["pos"]).max(), "dt", O.dt, E.scalars()["dt"])
    if sro != srg or np.abs(so['pos'] - sg['pos']).max()>0 or np.abs(so['force'] - sg['force']).max()>0 or np.abs(so['torque'] - sg['torque']).max()>0:
        print(" only oracle pot:", sorted(so_ - sg_)[:10]); print(" only gpu pot:", sorted(sg_ - so_)[:10])
        print(" only oracle real:", sorted(sro - srg)[:10]); print(" only gpu real:", sorted(srg - sro)[:10])
        for (a, b) in sorted(so_ - sg_)[:3]:
            if a < 400 and b < 400:
                print("  pair", a, b, "r", S.radius[a], S.radius[b], "pos", sg["pos"][a], sg["pos"][b], "dist", np.linalg.norm(sg["pos"][a] - sg["pos"][b]))
        bad=np.nonzero(np.any(so['force']!=sg['force'],axis=1)|np.any(so['torque']!=sg['torque'],axis=1)|np.any(so['pos']!=sg['pos'],axis=1))[0]
        print(" bad particles", bad[:10])
        for k in ('force','torque','vel','angVel','pos'):
            print(k, np.abs(so[k]-sg[k]).max())
        i=bad[0]
        print(" particle",i,"force O",so['force'][i],"G",sg['force'][i], "torque O", so['torque'][i], "G", sg['torque'][i])
        inc=[(tuple(x),k) for k,x in enumerate(ro['ids']) if i in x]
        print(" oracle incident:",[x for x,_ in inc])
        for x,k in inc:
            kg=[tuple(y) for y in rg['ids']].index(x) if x in srg else None
            if kg is None: print("  missing on gpu",x); continue
            for f in ('uN','force','contPt','trsf','phys'):
                if not np.array_equal(ro[f][k],rg[f][kg]): print("  contact",x,f,"O",ro[f][k],"G",rg[f][kg])
        break
