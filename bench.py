#!/usr/bin/env python
"""bench.py — particle-steps/s of the DEM hot path on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one Scene step (Leapfrog -> collider -> ContactLoop -> DynDt) over the whole packing.
Headline workload: BASELINE.json configs[1] — 1M polydisperse spheres (PSD 1:4, log-uniform radii) per GPU, dense bed settling under gravity in a box of
5 walls, linear IdealElPl law; synthetic: the committed pre-settled 20k-sphere bed (tests/golden/cfg2_sample_20k.npz, made by
scripts/make_cpu_sample.py with woo_b200.scenes.deposition seed 2) tiled 10x5 per GPU and re-settled in the untimed preparation phase.
For N>1 it is ONE box N times as long, cut into N x-slabs (weak scaling): per step every rank refreshes the pos/vel/angVel of its ghosts
and at collider full runs spheres migrate with their contact history (woo_b200/csrc/wdem_domain.cuh, DESIGN.md "Multi-GPU").
value  = particle-steps/s with the state resident in HBM (CUDA events on the engine's stream, max over ranks).
e2e    = same metric through the C-ABI with pinned HOST buffers: per step H2D of pos/vel/angVel (+ids when decomposed), one step, D2H of pos/vel
         and the kinetic energy (wdem_step_e2e / wdem_step_e2e_owned).
configs = the other BASELINE.json configs at their named sizes, a few seconds each, untimed by the headline: cfg1 (10k, with the CPU port on the SAME 10k),
         cfg3 (4M Hertz/DMT, STRONG-scaled over the N GPUs: the north-star target), cfg4 (periodic simple shear, 1M per GPU, ring-closed slabs), cfg5 (2M drum of facets).
mgpu_parity (N>1) = before anything is timed, the real multi-process path (NCCL migration + CUDA-IPC peer mailboxes) reproduces a single-GPU engine bit for bit
         on three scenes, on these very ranks.
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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, windows):
        """windows: [(t0, t1)] of the timed regions; only samples taken inside them count."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if len(r) >= 7 and any(a <= t <= b for a, b in windows)]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
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


def _tiling(n):
    d = np.load(SAMPLE)
    m = len(d["radius"]); tiles = max(1, n // m)
    tx = int(np.ceil(np.sqrt(2 * tiles)))
    while tiles % tx:
        tx += 1
    return d, m, tiles, tx, tiles // tx


def workload_tiled(n):
    """cfg2 at full size without hours of settling: the committed pre-settled 20k-sphere PSD 1:4 bed (tests/golden/cfg2_sample_20k.npz,
    made by scripts/make_cpu_sample.py with the same generator/material/law) is tiled in x and y inside one box of 5 walls; the former
    side walls of the tiles become internal planes across which spheres of neighbouring tiles touch; the bed re-settles in the untimed
    preparation phase (--settle steps)."""
    return workload_slab(n, 0, 1)


def workload_slab(n, rank, world):
    """The part of the N-GPU scene that rank `rank` uploads: the tiled bed repeated `world` times along x inside ONE box
    of 5 walls; slab boundaries are the tile-group boundaries.  Sphere ids are scene-wide (walls take ids 0..4)."""
    from woo_b200 import scenes as sc
    d, m, tiles, tx, ty = _tiling(n)
    L = d["hi"] - d["lo"]
    off = np.array([[(rank * tx + i) * L[0], j * L[1], 0.] for i in range(tx) for j in range(ty)])
    S = workload(m, seed=2)
    S.pos = (d["pos"][None, :, :] + off[:, None, :]).reshape(-1, 3)
    S.radius = np.tile(d["radius"], tiles); S.vel = np.tile(d["vel"], (tiles, 1)); S.angVel = np.tile(d["angVel"], (tiles, 1))
    S.mat = S.mask = S.flags = None
    lo = d["lo"]; hi = np.array([lo[0] + world * tx * L[0], lo[1] + ty * L[1], d["hi"][2]])
    S.walls = sc.box_walls(tuple(lo), tuple(hi), mat=0)
    S.params.dt = float(d["dt"]) if LAW == "linear" else float("nan")    # another law: start from dtSafety * pWaveDt, DynDt takes over
    nloc = m * tiles; nb = len(S.walls)
    if world > 1:
        S.boundaries_first = True
        S.ids = (nb + rank * nloc + np.arange(nloc)).astype(np.int32)
        S.global_max_id = nb + world * nloc - 1
    cuts = np.array([lo[0] + k * tx * L[0] for k in range(1, world)], dtype=np.float64)
    return S.finalize(), dict(tiles=tiles * world, tiles_x=tx * world, tiles_y=ty, cuts=cuts)


def bind_to_gpu_numa(local, world):
    """e2e at N>1: every rank pushes its state through host memory each step; keep the rank's threads (and so its pinned buffers, first touch)
    on the cores next to its GPU instead of letting all ranks crowd one NUMA node."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:00.0/local_cpulist" % (dom, bus)
        cpus = []
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if len(cpus) >= 1:
            os.sched_setaffinity(0, cpus)
            return "%d cpus near GPU %d" % (len(cpus), local)
    except Exception as e:   # noqa: BLE001 - affinity is an optimisation only
        return "default affinity (%s)" % type(e).__name__
    return "default affinity"


# ---------------------------------------------------------------------------------------------------------------------------------
# N>1: the real multi-process path checked against a single-GPU engine before anything is timed
# ---------------------------------------------------------------------------------------------------------------------------------
def mgpu_parity(local, steps=200):
    """Rank 0 also runs the single-domain engine on its GPU; positions, velocities, forces, torques and the real-contact set of the
    N subdomains (NCCL migration at full runs, peer-memory halo per step) must be bit-identical to it."""
    import torch.distributed as dist
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    from woo_b200.domain import DistributedEngine
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {}
    cases = {
        "gas_hertz_4k": lambda: sc.random_gas(4000, box=1.0, r=.02, psd_ratio=3.0, seed=31, law=sc.LAW_HERTZ_DMT, en=.7),
        "periodic_shear_4k": lambda: sc.periodic_gas(4000, cell=(1.2, .8, .9), r=.02, psd_ratio=2.0, seed=33, gradV=[[0, 2., 0], [0, 0, 0], [0, 0, 0]]),
        "deposition_20k": lambda: sc.deposition(20000, seed=1),
    }
    # the same gas with lopsided cuts, re-balancing switched on, and spheres inserted and removed in mid-run (wdem_add_spheres / wdem_remove_spheres on every rank)
    cases["gas_inlet_outlet_rebalance_4k"] = cases["gas_hertz_4k"]

    def inlet_outlet(S, eng, k):
        """after 50 steps: 60 new spheres; after 100: 80 spheres leave and a newcomer takes over the oldest freed id (ParticleContainer::findFreeId)"""
        import math
        rng = np.random.default_rng(77 + k); dens = S.mats[0][1]
        if k == 0:
            m = 60; pos = rng.uniform(.05, .95, (m, 3)); rad = .02 * np.exp(rng.uniform(math.log(.5), math.log(2.), m)); mass = dens * (4 / 3.) * math.pi * rad ** 3
            eng.add_spheres(np.arange(S.n + len(S.walls), S.n + len(S.walls) + m, dtype=np.int32), pos, rad, mass, .4 * mass * rad ** 2, vel=rng.uniform(-1, 1, (m, 3)))
        elif k == 1:
            gone = np.sort(rng.choice(S.sphere_ids(), 80, replace=False)).astype(np.int32)
            eng.remove_spheres(gone)
            rad = np.array([.03]); mass = dens * (4 / 3.) * math.pi * rad ** 3
            eng.add_spheres(gone[:1], np.array([[.5, .5, .5]]), rad, mass, .4 * mass * rad ** 2, vel=np.array([[.3, .2, -.1]]))

    for name, make in cases.items():
        S = make()
        io = name.startswith("gas_inlet")
        lo, hi = float(S.pos[:, 0].min()), float(S.pos[:, 0].max())
        D = DistributedEngine(S, axis=0, device=local, cuts=[lo + (hi - lo) * .5 * (c + 1) / world for c in range(world - 1)] if io else None)
        if io:
            D.set_rebalance(.1, 2)
        done = 0
        while done < steps:
            k = min(50, steps - done); D.step(k); done += k
            if io and done in (50, 100):
                inlet_outlet(S, D, done // 50 - 1)
        o = D.owned(); con = D.E.contacts(True)
        parts = [None] * world
        dist.all_gather_object(parts, dict(o=o, ids=con["ids"], force=con["force"], stats=D.domain_stats(), peer=bool(D.E.L.wdem_peer_active(D.E.h)), dom=D.domain() if io else None))
        D.E.close()
        if rank == 0:
            E = Engine(S.params, device=local).load(S)
            if io:      # the single-domain engine goes through the same insertions and removals; its download is in upload order, the subdomains' by id
                E.step(50); inlet_outlet(S, E, 0); E.step(50); inlet_outlet(S, E, 1); E.step(steps - 100)
            else:
                E.step(steps)
            ref = E.spheres(); rc = E.contacts(True)
            ids = np.concatenate([p["o"]["ids"] for p in parts]); order = np.argsort(ids)
            if io:      # upload order of the single-domain engine: the initial spheres, then the 60 new ones; removal closes the order up; the newcomer comes last
                rng1 = np.random.default_rng(78); gone = np.sort(rng1.choice(S.sphere_ids(), 80, replace=False)).astype(np.int32)
                ref_ids = np.concatenate([S.sphere_ids(), np.arange(S.n + len(S.walls), S.n + len(S.walls) + 60)]).astype(np.int32)
                ref_ids = np.concatenate([ref_ids[~np.isin(ref_ids, gone)], gone[:1]])
                ro = np.argsort(ref_ids, kind="stable"); ok = len(ids) == len(ref_ids) == len(ref["pos"]) and np.array_equal(ref_ids[ro], ids[order])
                ref = {k_: v[ro] for k_, v in ref.items()}
            else:
                ok = len(ids) == S.n and np.array_equal(ids[order], S.sphere_ids())
            E.close()
            diff = []
            for k_ in ("pos", "vel", "angVel", "force", "torque"):
                got = np.concatenate([p["o"][k_] for p in parts])[order]
                if not (ok and np.array_equal(got, ref[k_])):
                    diff.append(k_)
            cid = np.concatenate([p["ids"] for p in parts]); key = cid[:, 0].astype(np.int64) << 32 | cid[:, 1]
            co = np.argsort(key)
            csame = np.array_equal(cid[co], rc["ids"]) and np.array_equal(np.concatenate([p["force"] for p in parts])[co], rc["force"])
            out[name] = {"ok": bool(ok and not diff and csame), "bit_identical": not diff, "real_contacts": int(len(rc["ids"])), "contacts_identical": bool(csame),
                         "migrated": int(sum(p["stats"]["migrated_in"] for p in parts)), "ghosts": int(sum(p["stats"]["ghosts"] for p in parts)),
                         "peer_memory_halo": parts[0]["peer"], "steps": steps, "differs": diff}
            if io:
                out[name].update({"owned_per_rank": [int(p["stats"]["owned"]) for p in parts], "rebalanced": int(parts[0]["dom"][1]), "inserted": 61, "removed": 80})
    return out


# ---------------------------------------------------------------------------------------------------------------------------------
# the other BASELINE configs at their named sizes
# ---------------------------------------------------------------------------------------------------------------------------------
def _hertz_bed(n, layers=16, aspect_xy=8.0, seed=3):
    """cfg3: monodisperse r=1 mm HertzMat(rho=2500, E=1e8, tanPhi=.5), poisson .2, en .5, noAttraction, damping 0, as a lattice `layers`
    spheres high in a box 8:1 in plan (x = slab axis), thrown downwards so that the bed forms within the preparation phase."""
    import math
    from woo_b200 import scenes as sc
    rng = np.random.default_rng(seed)
    r = 1e-3; sp = 2.05 * r
    ny = max(1, int(round(math.sqrt(n / (layers * aspect_xy))))); nx = int(math.ceil(n / (layers * ny)))
    p = sc.lattice(nx, ny, layers, sp, (0, 0, 0), 0.05 * r, rng)[:n]
    S = sc.SceneSpec()
    S.mats = [sc.hertz_mat(density=2500., young=1e8, tanPhi=.5, surfEnergy=0., alpha=0.)]
    S.pos, S.radius = p, np.full(len(p), r)
    S.vel = np.tile([0., 0., -.3], (len(p), 1)) * rng.uniform(.5, 1.5, (len(p), 1)) + rng.uniform(-.2, .2, (len(p), 3))
    S.walls = sc.box_walls((0, 0, 0), (nx * sp, ny * sp, layers * sp), mat=0)
    S.params = sc.Params(gravity=(0, 0, -9.81), damping=0., law=sc.LAW_HERTZ_DMT, verletDist=-.05, dtSafety=.9, dynDtPeriod=100, en=.5, poisson=.2, noAttraction=1)
    return S.finalize()


def _time_case(make, n_label, settle, steps, world, local, axis=0):
    """settle (untimed) then `steps` timed steps of one scene, single-domain or cut into `world` slabs; device time = max over ranks."""
    import torch
    import torch.distributed as dist
    from woo_b200.engine import Engine
    t0 = time.time()
    S = make()
    if world > 1:
        from woo_b200.domain import DistributedEngine
        D = DistributedEngine(S, axis=axis, device=local); E = D.E
    else:
        E = Engine(S.params, device=local).load(S)
    done = 0
    while done < settle:
        k = min(1000, settle - done); E.step(k); done += k
    E.step(3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    E.set_profiling(True)
    st = E.step(steps)
    E.set_profiling(False)
    d = st.asdict()
    t = torch.tensor([d["device_ms"]], dtype=torch.float64, device="cuda")
    c = torch.tensor([d["n_spheres"], d["sum_real"], d["sum_potential_evaluated"], d["n_real"], d["n_contacts"], d["n_rebuilds"] if world == 1 else 0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ms = float(t.item()); ntot, sum_real, sum_pot, n_real, n_con, _ = [float(x) for x in c.cpu()]
    n_reb = d["n_rebuilds"]
    E.close()
    peak, _ = peaks()
    step_bytes = (ntot * B_P * steps + B_C * sum_real + B_Q * sum_pot + n_reb * (ntot * B_R + B_PAIR * n_con)) / steps
    return {"n_spheres": int(ntot), "n_facets": len(S.facets), "value": ntot * steps / (ms * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms / steps, "steps": steps,
            "settle_steps": settle, "real_contacts": int(n_real), "candidate_contacts": int(n_con), "rebuilds": int(n_reb),
            "step_roofline_frac": step_bytes / (ms / steps * 1e-3) / 1e9 / (peak * world),
            "phase_ms_per_step_rank0": {k: d[k] / steps for k in ("ms_integrate", "ms_rebuild", "ms_contact", "ms_reduce", "ms_dyndt")},
            "wall_s_total": round(time.time() - t0, 1)}


def bench_configs(args, rank, world, local):
    from woo_b200 import scenes as sc
    out = {}
    sc3 = args.cfg3_n
    r = _time_case(lambda: _hertz_bed(sc3), sc3, args.cfg_settle, 200, world, local)
    r["workload"] = "cfg3: %d monodisperse r=1mm spheres, HertzMat(E=1e8), Cp2_HertzMat_HertzPhys(poisson=.2,en=.5)+Law2_L6Geom_HertzPhys_DMT(noAttraction), damping 0, 16-layer bed in a box of 5 walls; TOTAL size fixed: %s" % (
        sc3, "strong-scaled over %d x-slabs (quantile cuts)" % world if world > 1 else "single GPU")
    r["scaling"] = "strong"
    out["cfg3_hertz_4M"] = r
    n4 = args.cfg4_n * world
    r = _time_case(lambda: sc.periodic_shear(n4, solid=.6, rate=1., vfluct=.01), n4, min(args.cfg_settle, 3000), 200, world, local)
    r["workload"] = "cfg4: %d monodisperse spheres in a periodic cell under simple shear (gradV2, rate 1/s), linear law; %s" % (
        n4, "%d ring-closed x-slabs, 1M per GPU (weak)" % world if world > 1 else "single GPU")
    r["scaling"] = "weak"
    out["cfg4_periodic_shear"] = r
    if world == 1:
        n5 = args.cfg5_n
        r = _time_case(lambda: sc.drum(n5, div=256, axial=64, vz0=.5, vfluct=.2), n5, min(args.cfg_settle, 3000), 200, 1, local)
        r["workload"] = "cfg5: %d spheres r=2mm in a horizontal drum of 32768 static facets (fakeVel = surface speed at 5 rad/s) + 2 end walls, linear law" % n5
        out["cfg5_drum_2M"] = r
        out["cfg1_deposition_10k"] = bench_cfg1(args, local)
    return out


def bench_cfg1(args, local):
    """cfg1: the one config the reference itself is timed on (obsolete/perf/inclined.py): 10k monodisperse spheres, 20000 steps from the lattice.
    GPU: the whole 20000-step run.  CPU port: the first `cpu_steps` steps of the SAME scene, beside the GPU's time for the same window."""
    from woo_b200 import scenes as sc
    from woo_b200.engine import Engine
    from _harness import Oracle, lib
    n, total, cpu_steps = 10000, 20000, args.cfg1_cpu_steps
    S = sc.deposition(n, seed=1)
    E = Engine(S.params, device=local).load(S)
    E.step(50)                                            # warm-up: allocations, first collider runs
    E.close()
    E = Engine(S.params, device=local).load(S)
    st = E.step(cpu_steps); ms_head = st.device_ms
    st2 = E.step(total - cpu_steps); ms_tail = st2.device_ms
    real = int(st2.n_real); E.close()
    lib().orc_set_num_threads(os.cpu_count() or 1)
    O = Oracle(S)
    t0 = time.time(); O.step(cpu_steps); cpu_s = time.time() - t0
    return {"workload": "cfg1: 10000 monodisperse r=5mm spheres, gravity deposition in a box of 5 walls, linear law, DynDt(100); 20000 steps from the jittered lattice",
            "n_spheres": n, "steps": total, "value": n * total / ((ms_head + ms_tail) * 1e-3), "unit": "particle-steps/s", "ms_per_step": (ms_head + ms_tail) / total,
            "value_settled_tail": n * (total - cpu_steps) / (ms_tail * 1e-3), "real_contacts": real, "rebuilds": int(st.n_rebuilds + st2.n_rebuilds),
            "same_window": {"steps": cpu_steps, "gpu_value": n * cpu_steps / (ms_head * 1e-3), "cpu_port_value": n * cpu_steps / cpu_s, "cpu_threads": lib().orc_num_threads(),
                            "note": "first %d steps of the same 10k scene on both sides (free fall: a collider full run almost every step)" % cpu_steps}}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from woo_b200.engine import Engine
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa(local, world) if world > 1 else "n/a"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    clocks = ClockSampler(local); clocks.start()
    parity = None
    if world > 1 and not args.no_parity:
        parity = mgpu_parity(local)
        if rank == 0:
            print("mgpu_parity: %s" % json.dumps(parity), file=sys.stderr, flush=True)
    n = args.n
    D = None
    if world > 1:
        # one scene, slab-decomposed along x: one subdomain per GPU (woo_b200/domain.py)
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
    w1 = time.time()
    wall = w1 - w0
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
        # stats=False: each call returns when that step's pos/vel/kinetic energy are in the host buffers; the contact pass preparing the next step's forces
        # overlaps the next call's upload (include/wdem.h).  The closing barrier() synchronises the device, so the last step's contact pass is inside the timed region.
        for _ in range(3):
            E.step_e2e(1, tp.numpy(), tv.numpy(), tw.numpy(), po.numpy(), vo.numpy(), stats=False); tp, po = po, tp; tv, vo = vo, tv
        barrier()
        e0 = time.time()
        for _ in range(ke):
            ek, _st = E.step_e2e(1, tp.numpy(), tv.numpy(), tw.numpy(), po.numpy(), vo.numpy(), stats=False)
            tp, po = po, tp; tv, vo = vo, tv     # the step's output buffers are the next step's host-side input (no host copy)
        barrier()
        e1 = time.time()
        h2d, d2h = int(n * 72), int(n * 48 + 8)
    else:
        # decomposed: every rank round-trips the spheres it owns, addressed by id, through wdem_step_e2e_owned (the chunk pipeline of the single-GPU path per subdomain;
        # stats=False: the contact pass overlaps the next call's upload).  When spheres migrate the owned set changes and the rank fetches it again (wdem_download_owned).
        def fresh():
            o = D.owned(); m_ = len(o["ids"])
            pinz = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return dict(m=m_, ids=pinz(o["ids"]), pos=pinz(o["pos"]), vel=pinz(o["vel"]), w=pinz(o["angVel"]), po=pinz(np.empty_like(o["pos"])), vo=pinz(np.empty_like(o["vel"])))
        B = fresh()

        def one():
            nonlocal B
            ek, changed, _st = D.step_e2e_owned(B["ids"].numpy(), B["pos"].numpy(), B["vel"].numpy(), B["w"].numpy(), B["po"].numpy(), B["vo"].numpy(), stats=False)
            if changed:
                B = fresh()
            else:
                B["pos"], B["po"] = B["po"], B["pos"]; B["vel"], B["vo"] = B["vo"], B["vel"]      # outputs become the next step's inputs (no host copy)
        for _ in range(3):
            one()
        barrier()
        e0 = time.time()
        for _ in range(ke):
            one()
        barrier()
        e1 = time.time()
        m = B["m"]
        h2d, d2h = int(m * 76), int(m * 48 + 8)
    e2e_s = e1 - e0
    ck = clocks.stop([(w0, w1), (e0, e1)])
    E.close()
    # ---- reduce over ranks (max of times)
    t = torch.tensor([dev_ms, e2e_s * 1e3, wall * 1e3], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([stats["sum_real"], stats["sum_potential_evaluated"], stats["n_rebuilds"], stats["n_launches"], stats["n_real"], stats["n_contacts"], h2d, d2h], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, wall_ms = [float(x) for x in t.cpu()]
    sum_real, sum_pot, n_reb, n_launch, n_real, n_con, h2d, d2h = [float(x) for x in cnt.cpu()]
    h2d, d2h = int(h2d), int(d2h)      # whole job: summed over the ranks, like `value`
    configs = None
    if not args.no_configs:
        configs = bench_configs(args, rank, world, local)
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
    law_s = ("Leapfrog+collider+ContactLoop(Cg2_Sphere_Sphere/Wall_Sphere_L6Geom, Cp2_FrictMat_FrictPhys, Law2_L6Geom_FrictPhys_IdealElPl)+DynDt" if LAW == "linear" else
             "HertzMat(E=1e8), Cp2_HertzMat_HertzPhys(poisson=.2, en=.5)+Law2_L6Geom_HertzPhys_DMT, damping 0, DynDt (cfg3's law and material on the cfg2 bed)")
    tr = ncu_traffic() or {}
    out = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: %d polydisperse spheres (PSD 1:4) per GPU settled under gravity, %s" % (n, law_s),
                   "n_spheres_per_gpu": n, "n_spheres_total": ntot, "law": LAW,
                   "initial_state": ("tiled pre-settled 20k bed (tests/golden/cfg2_sample_20k.npz) %dx%d, re-settled" % (tiling["tiles_x"], tiling["tiles_y"])) if tiling else "jittered lattice", "settle_steps": args.settle, "real_contacts": n_real, "candidate_contacts": n_con, "rebuilds_in_timed_region": n_reb / world,      # every subdomain does its full runs in the same steps
                   "l2_policy": "inputs larger than L2 (per-step working set %.0f MB > 126 MB)" % (step_bytes / 1e6), "parallelism": ("%d x-slabs of one box, 1 per GPU; ghosts refreshed per step over NVLink peer memory, spheres migrated at collider runs over NCCL" % world) if world > 1 else "single GPU",
                   "settle_wall_s": round(settle_s, 1), "cpu_affinity": affinity},
        "roofline": {"bound": "hbm", "kernel": "k_contact_ss<IdealElPl>" if LAW == "linear" else "k_contact_ss<HertzDMT>", "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                     "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"), "traffic_real_contacts": tr.get("real_contacts"),
                     "algorithmic_bytes_per_launch": bytes_contact, "ms_per_launch": ms_contact,
                     "step": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (stats["device_ms"] / K * 1e-3) / 1e9, "frac": step_bytes / (stats["device_ms"] / K * 1e-3) / 1e9 / peak},
                     "phase_ms_per_step": {k: stats[k] / K for k in ("ms_integrate", "ms_rebuild", "ms_contact", "ms_reduce", "ms_dyndt")}},
        "e2e": {"value": ntot * ke / (e2e_ms * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": ke,
                "api": "wdem_step_e2e (C-ABI, pinned host buffers, stats=NULL: a step's contact pass overlaps the next step's upload)" if D is None else "wdem_step_e2e_owned per rank (C-ABI, pinned host buffers, ids + pos/vel/angVel in, pos/vel out, stats=NULL)"},
        "gpu_launches": int(n_launch), "clocks": ck, "wall_ms_per_step": wall_ms / K,
    }
    if parity is not None:
        out["mgpu_parity"] = parity
    if configs is not None:
        out["configs"] = configs
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


def cpu_baseline(steps, warmup=3, large=True):
    """The CPU oracle (kind 'port': the reference cannot be compiled here) on ALL host threads (also under torchrun, which exports OMP_NUM_THREADS=1),
    on a bounded sample: the pre-settled 20k bed (cache-resident) and, beside it, a 200k-sphere tiling of the same bed (out of the last-level cache)."""
    from _harness import Oracle, lib
    cores = os.cpu_count() or 1
    lib().orc_set_num_threads(cores)
    S = load_sample()
    sample = "pre-settled 20k-sphere PSD 1:4 packing (tests/golden/cfg2_sample_20k.npz), same material/law as the 1M workload"
    if S is None:
        S = workload(20000, seed=2); sample = "20k-sphere PSD 1:4 lattice (unsettled: golden sample missing)"
    S.params.deterministic = 0      # the reference's default: parallel force application (ContactLoop.cpp:195-198)
    O = Oracle(S)
    O.step(warmup)
    t0 = time.time(); O.step(steps); dt = time.time() - t0
    out = {"value": S.n * steps / dt, "unit": "particle-steps/s", "cores": lib().orc_num_threads(), "kind": "port", "sample": sample, "sample_spheres": S.n, "steps": steps,
           "real_contacts": int(O.num_contacts(True)), "ms_per_step": dt / steps * 1e3,
           "note": "soft baseline: restatement, not the reference build; its collider sweep is serial where the reference's is chunk-parallel (InsertionSortCollider.cpp:124-187)"}
    if large and os.path.exists(SAMPLE):
        del O
        S2, _ = workload_tiled(200000); S2.params.deterministic = 0
        O2 = Oracle(S2); O2.step(warmup)
        k2 = max(5, steps // 10)
        t0 = time.time(); O2.step(k2); dt2 = time.time() - t0
        out["large_sample"] = {"value": S2.n * k2 / dt2, "sample_spheres": S2.n, "steps": k2, "ms_per_step": dt2 / k2 * 1e3, "real_contacts": int(O2.num_contacts(True)),
                               "sample": "the same bed tiled to 200k spheres (out of LLC)"}
    return out


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
           "config": {"workload": "cfg2 sample: " + cb["sample"], "n_spheres_total": cb["sample_spheres"], "same_config_as_gpu_arm": False,
                      "note": "per-particle-step normalised; the GPU arm runs 1M spheres per GPU, the CPU arm a 20k (and a 200k) sample of the same bed"},
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
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1/cfg3/cfg4/cfg5 block")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the multi-process parity check")
    ap.add_argument("--cfg3-n", type=int, default=4_000_000)
    ap.add_argument("--cfg4-n", type=int, default=1_000_000, help="cfg4 spheres per GPU")
    ap.add_argument("--cfg5-n", type=int, default=2_000_000)
    ap.add_argument("--cfg-settle", type=int, default=6000)
    ap.add_argument("--cfg1-cpu-steps", type=int, default=1000)
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
