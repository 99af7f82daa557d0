"""N>1 paths.  GPU: torchrun with one subdomain per GPU over NCCL must match the single-GPU engine bit for bit (skipped with <2 GPUs).
CPU: world_size-2 gloo test of the host-side partition logic (cuts, ownership, id-buffer broadcast)."""
import os, subprocess, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["gas", "peri", "deposition"])
def test_nccl_subdomains_match_single_gpu(scene):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)), "--master-addr", "127.0.0.1", "--master-port", "29641",
           os.path.join(ROOT, "scripts", "mgpu_check.py"), scene, "200"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from woo_b200 import scenes as sc
from woo_b200.domain import quantile_cuts, owner_of
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for S, axis in ((sc.random_gas(1000, seed=5, psd_ratio=3.0), 0), (sc.periodic_gas(1000, seed=6, gradV=[[0, 1., 0], [0, 0, 0], [0, 0, 0]], shear0=.2), 1)):
    cuts = quantile_cuts(S, world, axis)
    c = torch.from_numpy(cuts.copy()); dist.broadcast(c, 0)
    assert np.array_equal(c.numpy(), cuts), "ranks computed different cuts"
    mine = owner_of(S, cuts, axis) == rank
    cnt = torch.tensor([int(mine.sum())]); dist.all_reduce(cnt)
    assert int(cnt) == S.n, "spheres lost or duplicated by the partition"
    assert abs(int(mine.sum()) - S.n / world) <= 1, "quantile cuts are not balanced"
    owned = torch.zeros(S.n, dtype=torch.int32); owned[torch.from_numpy(np.nonzero(mine)[0])] = 1; dist.all_reduce(owned)
    assert bool((owned == 1).all())
# the 128-byte communicator id travels as a uint8 tensor
idb = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(idb, 0)
assert bytes(idb.numpy().tobytes()) == bytes(range(128))
dist.destroy_process_group()
print("GLOO OK", rank)
'''


def test_partition_logic_gloo_world2(tmp_path):
    w = tmp_path / "worker.py"; w.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29643", str(w), ROOT]
    env = dict(os.environ); env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.count("GLOO OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]
