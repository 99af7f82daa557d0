"""torchrun --nproc-per-node N scripts/mgpu_check.py [scene] [steps]: N subdomains over NCCL must reproduce the single-GPU engine bit for bit.
Rank 0 runs the single-domain engine on its own GPU with the same scene and compares positions, velocities, forces and the real contact set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from woo_b200 import scenes as sc
from woo_b200.engine import Engine
from woo_b200.domain import DistributedEngine

scene = sys.argv[1] if len(sys.argv) > 1 else "gas"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
if scene == "gas":
    S = sc.random_gas(4000, box=1.0, r=.02, psd_ratio=3.0, seed=31, law=sc.LAW_HERTZ_DMT, en=.7)
elif scene == "peri":
    S = sc.periodic_gas(4000, cell=(1.2, .8, .9), r=.02, psd_ratio=2.0, seed=33, gradV=[[0, 2., 0], [0, 0, 0], [0, 0, 0]])
else:
    S = sc.deposition(20000, seed=1)
D = DistributedEngine(S, axis=0, device=local)
done = 0
while done < steps:
    k = min(50, steps - done); st = D.step(k); done += k
o = D.owned(); con = D.E.contacts(True)
parts = [None] * world
dist.all_gather_object(parts, dict(o=o, ids=con["ids"], force=con["force"], stats=D.domain_stats(), ms=st.device_ms / k))
ok = True
if rank == 0:
    E = Engine(S.params, device=local).load(S); E.step(steps)
    ref = E.spheres(); rc = E.contacts(True)
    ids = np.concatenate([p["o"]["ids"] for p in parts]); order = np.argsort(ids)
    ok = len(ids) == S.n and np.array_equal(ids[order], S.sphere_ids())
    for k_ in ("pos", "vel", "angVel", "force", "torque"):
        got = np.concatenate([p["o"][k_] for p in parts])[order]
        same = np.array_equal(got, ref[k_]); ok &= same
        print(k_, "bit-identical" if same else "DIFFERS max %g" % np.abs(got - ref[k_]).max())
    cid = np.concatenate([p["ids"] for p in parts]); key = cid[:, 0].astype(np.int64) << 32 | cid[:, 1]
    co = np.argsort(key)
    same = np.array_equal(cid[co], rc["ids"]) and np.array_equal(np.concatenate([p["force"] for p in parts])[co], rc["force"]); ok &= same
    print("real contacts", len(cid), "vs", len(rc["ids"]), "identical" if same else "DIFFER")
    print("domain stats", [p["stats"] for p in parts], "ms/step", [round(p["ms"], 4) for p in parts])
    print("MGPU OK" if ok else "MGPU FAIL", scene, "world", world, flush=True)
t = torch.tensor([1 if ok else 0], device="cuda"); dist.broadcast(t, 0)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
