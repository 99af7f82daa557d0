#!/usr/bin/env python
"""bench.py — particle-steps/s of the DEM hot path on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one Scene step (Leapfrog -> collider -> ContactLoop -> DynDt) over the whole packing.
Workload: BASELINE.json configs[1] — 1M polydisperse spheres (PSD 1:4, log-uniform radii) per GPU, dense bed settling under gravity in a box of
5 walls, linear IdealElPl law; synthetic: the committed pre-settled 20k-sphere bed (tests/golden/cfg2_sample_20k.npz, made by
scripts/make_cpu_sample.py with woo_b200.scenes.deposition seed 2) tiled 10x5 per GPU and re-settled in the untimed preparation phase.
For N>1 it is ONE box N times as long, cut into N x-slabs (weak scaling): per step every rank refreshes the pos/vel/angVel of its ghosts
and at collider full runs spheres migrate with their contact history, all over NCCL inside libwdem.so (DESIGN.md "Multi-GPU").
value  = particle-steps/s with the state resident in HBM (CUDA events on the engine's stream, max over ranks).
e2e    = same metric through the C-ABI with pinned HOST buffers: per step H2D of pos/vel/angVel, one step, D2H of pos/vel (+ids/angVel
         when decomposed) and the kinetic energy.
--impl reference: the CPU restatement oracle (the reference itself cannot be built here: Eigen/Boost/TBB missing) on all
host threads, on a bounded sample (pre-settled 20k-sphere packing of the same PSD/material, tests/golden/cfg2_sample_20k.npz).
"""
import argparse, json, os, subprocess, sys, threading, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_FULL = 1_000_000
LAYERS = 40
SETTLE_STEPS = int(os.environ.get("WDEM_BENCH_SETTLE", "6000"))
SAMPLE = os.path.join(ROOT, "tests", "golden", "cfg2_sample_20k.npz")
# algorithmic bytes (SURVEY.md §8d / DESIGN.md)
B_P, B_C, B_Q, B_R, B_PAIR = 348, 552, 72, 200, 496
B_C_KERNEL = 480   # contact kernel's share of B_C (552 minus the 72 B read by the reduce kernel)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


LAW = "linear"     # --law hertz: the same bed under Cp2_HertzMat_HertzPhys + Law2_L6Geom_HertzPhys_DMT with cfg3's material (extra data point, not the default line)


def workload(n, seed):
    from woo_b200 import scenes as sc
    if LAW == "hertz":   # BASELINE.json configs[2] material and law: HertzMat(rho=2500, E=1e8, tanPhi=.5), poisson=.2, en=.5, noAttraction, damping 0
        return sc.deposition(n, r=1e-3, psd_ratio=4.0, seed=seed, law=sc.LAW_HERTZ_DMT, layers=LAYERS, damping=0., young=1e8, density=2500., en=.5)
    return sc.deposition(n, r=1e-3, psd_ratio=4.0, seed=seed, law=sc.LAW_IDEALELPL, layers=LAYERS, damping=.2, young=1e7, density=2500.)


def workload_tiled(n, tiles_x=None):
    """cfg2 at full size without hours of settling: the committed pre-settled 20k-sphere PSD 1:4 bed (tests/golden/cfg2_sample_20k.npz,
    made by scripts/make_cpu_sample.py with the same generator/material/law) is tiled in x and y inside one box of 5 walls; the former
    side walls of the tiles become internal planes across which spheres of neighbouring tiles touch; the bed re-settles in the untimed
    preparation phase (--settle steps)."""
    from woo_b200 import scenes as sc
    d = np.load(SAMPLE)
    m = len(d["radius"])
    tiles = max(1, n // m)
    tx = tiles_x or int(np.ceil(np.sqrt(2 * tiles)))
    while tiles % tx:
        tx += 1
    ty = tiles // tx
    L = d["hi"] - d["lo"]
    off = np.array([[i * L[0], j * L[1], 0.] for i in range(tx) for j in range(ty)])
    S = workload(m, seed=2)
    S.pos = (d["pos"][None, :, :] + off[:, None, :]).reshape(-1, 3)
    S.radius = np.tile(d["radius"], tiles); S.vel = np.tile(d["vel"], (tiles, 1)); S.angVel = np.tile(d["angVel"], (tiles, 1))
    S.mat = S.mask = S.flags = None
    lo = d["lo"]; hi = np.array([lo[0] + tx * L[0], lo[1] + ty * L[1], d["hi"][2]])
    S.walls = sc.box_walls(tuple(lo), tuple(hi), mat=0)
    S.params.dt = float(d["dt"]) if LAW == "linear" else float("nan")    # another law: start from dtSafety * pWaveDt, DynDt takes over
    return S.finalize(), dict(tiles=tiles, tiles_x=tx, tiles_y=ty)


def workload_slab(n, rank, world):
    """The part of the N-GPU scene that rank `rank` uploads: the tiled bed of workload_tiled() repeated `world` times along x inside ONE box
    of 5 walls; slab boundaries are the tile-group boundaries.  Sphere ids are scene-wide (walls take ids 0..4)."""
    from woo_b200 import scenes as sc
    d = np.load(SAMPLE)
    m = len(d["radius"]); tiles = max(1, n // m)
    tx = int(np.ceil(np.sqrt(2 * tiles)))
    while tiles % tx:
        tx += 1
    ty = tiles // tx
    L = d["hi"] - d["lo"]
    off = np.array([[(rank * tx + i) * L[0], j * L[1], 0.] for i in range(tx) for j in range(ty)])
    S = workload(m, seed=2)
    S.pos = (d["pos"][None, :, :] + off[:, None, :]).reshape(-1, 3)
    S.radius = np.tile(d["radius"], tiles); S.vel = np.tile(d["vel"], (tiles, 1)); S.angVel = np.tile(d["angVel"], (tiles, 1))
    S.mat = S.mask = S.flags = None
    lo = d["lo"]; hi = np.array([lo[0] + world * tx * L[0], lo[1] + ty * L[1], d["hi"][2]])
    S.walls = sc.box_walls(tuple(lo), tuple(hi), mat=0)
    S.params.dt = float(d["dt"]) if LAW == "linear" else float("nan")
    nloc = m * tiles; nb = len(S.walls)
    S.boundaries_first = True
    S.ids = (nb + rank * nloc + np.arange(nloc)).astype(np.int32)
    S.global_max_id = nb + world * nloc - 1
    cuts = np.array([lo[0] + k * tx * L[0] for k in range(1, world)], dtype=np.float64)
    return S.finalize(), dict(tiles=tiles * world, tiles_x=tx * world, tiles_y=ty, cuts=cuts)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from woo_b200.engine import Engine
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n = args.n
    D = None
    if world > 1:
        # one scene, slab-decomposed along x: one subdomain per GPU, halo exchange + migration over NCCL (woo_b200/domain.py)
        from woo_b200.domain import DistributedEngine
        S, tiling = workload_slab(n, rank, world); n = S.n
        D = DistributedEngine(S, axis=0, cuts=tiling["cuts"], device=local, sel=np.arange(S.n))
        E = D.E
    elif args.lattice or not os.path.exists(SAMPLE):
        S = workload(n, seed=2 + rank); tiling = None
        E = Engine(S.params, device=local).load(S)
    else:
        S, tiling = workload_tiled(n); n = S.n
        E = Engine(S.params, device=local).load(S)
    # settle (untimed preparation of the dense packing; includes many collider full runs)
    t0 = time.time()
    done = 0
    while done < args.settle:
        k = min(2000, args.settle - done); st_ = E.step(k); done += k
        if rank == 0:
            print("settle %d: rebuilds %d/%d, %.3f ms/step, real %d, Ek %.4g" % (done, st_.n_rebuilds, k, st_.device_ms / k, st_.n_real, E.kinetic_energy()), file=sys.stderr, flush=True)
    st = E.stats
    settle_s = time.time() - t0
    # warm-up
    E.step(max(args.warmup, 3))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local); clocks.start()
    # ---- timed region 1: K steps, state resident in HBM; per-phase CUDA events on the engine's stream
    E.set_profiling(True)
    barrier()
    rt = None
    if args.ncu_range:      # ncu --profile-from-start off: only the timed region's launches are listed (profiles/*launches*)
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12"); rt.cudaProfilerStart()
    w0 = time.time()
    st = E.step(args.steps)
    barrier()
    wall = time.time() - w0
    if rt is not None:
        rt.cudaProfilerStop()
    E.set_profiling(False)
    dev_ms = st.device_ms
    stats = st.asdict()
    # ---- timed region 2: end to end through the C-ABI with pinned host buffers
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    ke = min(args.steps, args.e2e_steps)
    if D is None:
        s = E.spheres()
        tp, tv, tw = pin(s["pos"]), pin(s["vel"]), pin(s["angVel"])
        po, vo = torch.empty_like(tp).pin_memory(), torch.empty_like(tv).pin_memory()
        for _ in range(3):
            E.step_e2e(1, tp.numpy(), tv.numpy(), tw.numpy(), po.numpy(), vo.numpy()); tp, po = po, tp; tv, vo = vo, tv
        barrier()
        e0 = time.time()
        for _ in range(ke):
            ek, _st = E.step_e2e(1, tp.numpy(), tv.numpy(), tw.numpy(), po.numpy(), vo.numpy())
            tp, po = po, tp; tv, vo = vo, tv     # the step's output buffers are the next step's host-side input (no host copy)
        barrier()
        e2e_s = time.time() - e0
        h2d, d2h = int(n * 72), int(n * 48 + 8)
    else:
        # decomposed: every rank round-trips the spheres it owns, addressed by id (the owned set changes when spheres migrate)
        o = D.owned(); m = len(o["ids"]); cap = int(m * 1.05) + 4096
        buf = lambda w, dt=np.float64: torch.from_numpy(np.zeros((cap, w) if w > 1 else cap, dtype=dt)).pin_memory()
        ii, pi_, vi, wi = buf(1, np.int32), buf(3), buf(3), buf(3)
        io, po, vo, wo = buf(1, np.int32), buf(3), buf(3), buf(3)
        ii.numpy()[:m] = o["ids"]; pi_.numpy()[:m] = o["pos"]; vi.numpy()[:m] = o["vel"]; wi.numpy()[:m] = o["angVel"]

        bufs = [[ii, pi_, vi, wi], [io, po, vo, wo]]

        def one(m):
            a, b = bufs
            m2, ek, _st = D.step_e2e(1, a[0].numpy()[:m], a[1].numpy()[:m], a[2].numpy()[:m], a[3].numpy()[:m], b[0].numpy(), b[1].numpy(), b[2].numpy(), b[3].numpy())
            bufs.reverse()          # outputs become the next step's inputs (no host copy)
            return m2
        for _ in range(3):
            m = one(m)
        barrier()
        e0 = time.time()
        for _ in range(ke):
            m = one(m)
        barrier()
        e2e_s = time.time() - e0
        h2d, d2h = int(m * 76), int(m * 76 + 8)
    ck = clocks.stop()
    # ---- reduce over ranks (max of times)
    t = torch.tensor([dev_ms, e2e_s * 1e3, wall * 1e3], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([stats["sum_real"], stats["sum_potential_evaluated"], stats["n_rebuilds"], stats["n_launches"], stats["n_real"], stats["n_contacts"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, wall_ms = [float(x) for x in t.cpu()]
    sum_real, sum_pot, n_reb, n_launch, n_real, n_con = [float(x) for x in cnt.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    K = args.steps
    ntot = n * world
    value = ntot * K / (dev_ms * 1e-3)
    # dominant kernel: fused contact kernel (Cg2+Cp2+Law2), per launch = per step (rank 0's numbers)
    ms_contact = stats["ms_contact"] / K
    bytes_contact = (B_C_KERNEL * stats["sum_real"] + B_Q * stats["sum_potential_evaluated"]) / K
    ach = bytes_contact / (ms_contact * 1e-3) / 1e9 if ms_contact > 0 else 0.
    step_bytes = (n * B_P * K + B_C * stats["sum_real"] + B_Q * stats["sum_potential_evaluated"] + stats["n_rebuilds"] * (n * B_R + B_PAIR * stats["n_contacts"])) / K
    out = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: 1M polydisperse spheres (PSD 1:4) settled under gravity, Leapfrog+collider+ContactLoop(Cg2_Sphere_Sphere/Wall_Sphere_L6Geom, Cp2_FrictMat_FrictPhys, Law2_L6Geom_FrictPhys_IdealElPl)+DynDt" if LAW == "linear" else
                   "cfg2 bed under cfg3's law and material: 1M polydisperse spheres (PSD 1:4), HertzMat(E=1e8), Cp2_HertzMat_HertzPhys(poisson=.2, en=.5)+Law2_L6Geom_HertzPhys_DMT, damping 0, DynDt",
                   "n_spheres_per_gpu": n, "initial_state": ("tiled pre-settled 20k bed (tests/golden/cfg2_sample_20k.npz) %dx%d, re-settled" % (tiling["tiles_x"], tiling["tiles_y"])) if tiling else "jittered lattice", "settle_steps": args.settle, "real_contacts": n_real, "candidate_contacts": n_con, "rebuilds_in_timed_region": n_reb,
                   "l2_policy": "inputs larger than L2 (per-step working set %.0f MB > 126 MB)" % (step_bytes / 1e6), "parallelism": ("%d x-slabs of one box, 1 per GPU; ghosts refreshed per step and spheres migrated at collider runs over NCCL" % world) if world > 1 else "single GPU",
                   "settle_wall_s": round(settle_s, 1)},
        "roofline": {"bound": "hbm", "kernel": "k_contact_ss<IdealElPl>" if LAW == "linear" else "k_contact_ss<HertzDMT>", "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak, "traffic": (ncu_traffic() or {}).get("dram_bytes_per_launch"), "traffic_source": (ncu_traffic() or {}).get("source"),
                     "algorithmic_bytes_per_launch": bytes_contact, "ms_per_launch": ms_contact,
                     "step": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (stats["device_ms"] / K * 1e-3) / 1e9, "frac": step_bytes / (stats["device_ms"] / K * 1e-3) / 1e9 / peak},
                     "phase_ms_per_step": {k: stats[k] / K for k in ("ms_integrate", "ms_rebuild", "ms_contact", "ms_reduce", "ms_dyndt")}},
        "e2e": {"value": ntot * ke / (e2e_ms * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": ke,
                "api": "wdem_step_e2e (C-ABI, pinned host buffers)" if D is None else "wdem_set_owned + wdem_step + wdem_download_owned per rank (C-ABI, pinned host buffers)"},
        "gpu_launches": int(n_launch), "clocks": ck, "wall_ms_per_step": wall_ms / K,
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(args.cpu_steps)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def load_sample():
    from woo_b200 import scenes as sc
    if not os.path.exists(SAMPLE):
        return None
    d = np.load(SAMPLE)
    S = workload(len(d["radius"]), seed=2)
    S.pos, S.radius, S.vel, S.angVel = d["pos"], d["radius"], d["vel"], d["angVel"]
    S.walls = sc.box_walls(tuple(d["lo"]), tuple(d["hi"]), mat=0)
    S.params.dt = float(d["dt"])
    return S.finalize()


def cpu_baseline(steps, warmup=3):
    """The CPU oracle (kind 'port': the reference cannot be compiled here) on all host threads, bounded sample."""
    from _harness import Oracle, lib
    S = load_sample()
    sample = "pre-settled 20k-sphere PSD 1:4 packing (tests/golden/cfg2_sample_20k.npz), same material/law as the 1M workload"
    if S is None:
        S = workload(20000, seed=2); sample = "20k-sphere PSD 1:4 lattice (unsettled: golden sample missing)"
    S.params.deterministic = 0      # the reference's default: parallel force application (ContactLoop.cpp:195-198)
    O = Oracle(S)
    cores = lib().orc_num_threads()
    O.step(warmup)
    t0 = time.time(); O.step(steps); dt = time.time() - t0
    return {"value": S.n * steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample, "steps": steps,
            "real_contacts": int(O.num_contacts(True)), "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 3)
    # bounded: the CPU arm times min(K, cpu_steps) steps per "step sample"
    k = min(K, args.cpu_steps)
    cb = cpu_baseline(k, W)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": "particle-steps/s", "value": cb["value"], "unit": "particle-steps/s", "n_gpus": world, "steps": k, "warmup": W,
           "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "cfg2 sample: " + cb["sample"]},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n", type=int, default=N_FULL, help="spheres per GPU (default: the 1M workload)")
    ap.add_argument("--settle", type=int, default=SETTLE_STEPS, help="untimed preparation steps before warm-up")
    ap.add_argument("--lattice", action="store_true", help="start from the jittered lattice instead of the tiled pre-settled bed")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--cpu-steps", type=int, default=200)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--law", default="linear", choices=["linear", "hertz"], help="hertz: extra data point with cfg3's law and material on the same bed")
    ap.add_argument("--ncu-range", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    global LAW
    LAW = args.law
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
