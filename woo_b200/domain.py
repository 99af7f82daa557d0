"""Host side of the slab decomposition: one Engine handle per subdomain.

* ``LocalCluster``  — all subdomains in this process, one host thread each (in-process transport; several subdomains may share a GPU).
* ``DistributedEngine`` — one subdomain per process under torchrun; halo exchange / migration over NCCL inside the C-ABI library, the
  128-byte NCCL id is broadcast with torch.distributed.

Both take a whole-scene SceneSpec; every subdomain uploads the boundaries/materials and the spheres of its slab.  Results are assembled
by particle id, so they can be compared with a single-domain run bit for bit (forces are reduced per particle in ascending-id order of the
contact partner, independent of the decomposition).
"""
import ctypes as C
import threading
import numpy as np
from . import capi
from .capi import ptr
from .engine import Engine, WdemError


def slab_coord(S, axis):
    """Coordinate along which slabs are cut: position (aperiodic) or fraction of the cell along the unsheared axis (periodic)."""
    if not S.params.periodic:
        return S.pos[:, axis]
    H = np.asarray(S.cell["hSize"], dtype=np.float64).reshape(3, 3)
    size = np.linalg.norm(H, axis=0)
    u = S.pos @ np.linalg.inv(H / size).T
    f = u[:, axis] / size[axis]
    return f - np.floor(f)


def quantile_cuts(S, world, axis):
    """world-1 boundaries with equal sphere counts per slab (SURVEY §8e "particle-count quantiles")."""
    x = np.sort(slab_coord(S, axis))
    return np.array([.5 * (x[len(x) * k // world - 1] + x[len(x) * k // world]) for k in range(1, world)], dtype=np.float64)


def owner_of(S, cuts, axis):
    return np.searchsorted(np.asarray(cuts), slab_coord(S, axis), side="right")


def _load_part(E, S, sel):
    E.spec = S
    E.upload_materials(S.mats)
    m, I = S.mass_inertia()
    ids = S.sphere_ids()
    E.upload_spheres(ids[sel], S.pos[sel], S.radius[sel], m[sel], I[sel], S.mat[sel], S.mask[sel], S.flags[sel], S.vel[sel], S.angVel[sel])
    E.upload_walls(S.wall_ids(), S.walls)
    E.upload_facets(S.facet_ids(), S.facets)
    E.upload_facet_groups(len(S.facets), getattr(S, "facet_groups", []))
    if S.cell is not None:
        E.set_cell(S.cell["hSize"], S.cell.get("gradV"), S.cell.get("nextGradV"), S.cell.get("homoDeform", 4))


def _set_domain(E, axis, cuts, max_id):
    cuts = np.ascontiguousarray(cuts, dtype=np.float64)
    E._chk(E.L.wdem_set_domain(E.h, C.c_int(axis), C.c_int(len(cuts)), ptr(cuts), C.c_int64(int(max_id))))


def _owned(E):
    n = E.L.wdem_num_owned(E.h)
    o = {k: np.zeros((n, 3)) for k in ("pos", "vel", "angVel", "force", "torque")}
    o["ori"] = np.zeros((n, 4)); o["ids"] = np.zeros(n, np.int32)
    E._chk(E.L.wdem_download_owned(E.h, ptr(o["ids"]), ptr(o["pos"]), ptr(o["ori"]), ptr(o["vel"]), ptr(o["angVel"]), ptr(o["force"]), ptr(o["torque"])))
    return o


def _set_owned(E, ids, pos=None, vel=None, angVel=None):
    E._chk(E.L.wdem_set_owned(E.h, C.c_int64(len(ids)), ptr(ids), ptr(pos), ptr(vel), ptr(angVel)))


def _domain_stats(E):
    v = [C.c_int64() for _ in range(5)]
    E._chk(E.L.wdem_domain_stats(E.h, *[C.byref(x) for x in v]))
    return dict(zip(("owned", "ghosts", "migrated_out", "migrated_in", "rows_in"), [x.value for x in v]))


def _set_rebalance(E, tolerance, min_full_runs=1):
    E._chk(E.L.wdem_set_rebalance(E.h, C.c_double(float(tolerance)), C.c_int(int(min_full_runs))))


def _get_domain(E, world):
    cuts = np.zeros(max(world - 1, 1)); nr = C.c_int64()
    E._chk(E.L.wdem_get_domain(E.h, ptr(cuts), C.byref(nr)))
    return cuts[:world - 1], nr.value


def _max_id(S):
    return S.n + len(S.walls) + len(S.facets) - 1


class LocalCluster:
    """``world`` subdomains of one scene inside this process."""

    def __init__(self, S, world, axis=0, cuts=None, devices=None, scramble=False):
        self.S, self.world, self.axis = S, world, axis
        self.L = capi.load()
        self.group = C.c_void_p(self.L.wdem_local_group_create(C.c_int(world)))
        self.cuts = quantile_cuts(S, world, axis) if cuts is None else np.asarray(cuts, dtype=np.float64)
        own = owner_of(S, self.cuts, axis)
        if scramble:      # deliberately wrong initial placement: the first collider run must migrate everything home
            own = (own + 1 + np.arange(S.n) % max(world - 1, 1)) % world
        devices = devices or [0] * world
        self.ids = np.sort(np.asarray(S.sphere_ids(), dtype=np.int32))      # live particle ids (add_spheres / remove_spheres keep it current)
        self.E = []
        for r in range(world):
            E = Engine(S.params, device=devices[r])
            E._chk(self.L.wdem_comm_init_local(E.h, self.group, C.c_int(r)))
            _set_domain(E, axis, self.cuts, _max_id(S))
            _load_part(E, S, np.nonzero(own == r)[0])
            self.E.append(E)

    def close(self):
        for E in self.E:
            E.close()
        self.E = []
        if self.group:
            self.L.wdem_local_group_destroy(self.group); self.group = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _all(self, fn):
        out, err = [None] * self.world, [None] * self.world

        def run(r):
            try:
                out[r] = fn(self.E[r])
            except Exception as e:   # noqa: BLE001 - re-raised below
                err[r] = e
        th = [threading.Thread(target=run, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None and "another subdomain failed" not in str(e):
                raise e
        for e in err:
            if e is not None:
                raise e
        return out

    def add_spheres(self, ids, pos, radius, mass, inertia, **kw):
        """wdem_add_spheres on every subdomain with the same list (each keeps what falls into its slab)."""
        for E in self.E:
            E.add_spheres(ids, pos, radius, mass, inertia, **kw)
        self.ids = np.sort(np.concatenate([self.ids, np.asarray(ids, dtype=np.int32)]))

    def remove_spheres(self, ids):
        """wdem_remove_spheres on every subdomain with the same list."""
        for E in self.E:
            E.remove_spheres(ids)
        self.ids = self.ids[~np.isin(self.ids, ids)]

    def set_rebalance(self, tolerance, min_full_runs=1):
        """Move the cuts to the quantiles of the slab coordinate at a full run whenever the largest subdomain exceeds (1+tolerance) x the mean (wdem_set_rebalance)."""
        for E in self.E:
            _set_rebalance(E, tolerance, min_full_runs)

    def domain(self):
        return _get_domain(self.E[0], self.world)

    def step(self, n=1):
        self.stats = self._all(lambda E: E.step(n))
        return self.stats

    def spheres(self):
        parts = [_owned(E) for E in self.E]
        ids = np.concatenate([p["ids"] for p in parts])
        order = np.argsort(ids, kind="stable")
        assert len(ids) == len(self.ids) and np.array_equal(ids[order], self.ids), "spheres lost or duplicated across subdomains"
        return {k: np.concatenate([p[k] for p in parts])[order] for k in parts[0] if k != "ids"}

    def contacts(self, real_only=False):
        parts = [E.contacts(real_only) for E in self.E]
        ids = np.concatenate([p["ids"] for p in parts])
        key = ids[:, 0].astype(np.int64) << 32 | ids[:, 1].astype(np.int64)
        order = np.argsort(key, kind="stable")
        return {k: np.concatenate([p[k] for p in parts])[order] for k in parts[0]}

    def scalars(self):
        s = [E.scalars() for E in self.E]
        assert all(x == s[0] for x in s), "subdomains disagree on dt/step/verletDist/nFullRuns: %r" % (s,)
        return s[0]

    def kinetic_energy(self):
        return sum(E.kinetic_energy() for E in self.E)

    def cell_hsize(self):
        return self.E[0].cell_hsize()

    def domain_stats(self):
        return [_domain_stats(E) for E in self.E]


class DistributedEngine:
    """One subdomain per process (torchrun: RANK/WORLD_SIZE/LOCAL_RANK); torch.distributed must be initialised by the caller."""

    def __init__(self, S, axis=0, cuts=None, device=None, sel=None):
        import torch
        import torch.distributed as dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        device = torch.cuda.current_device() if device is None else device
        self.E = E = Engine(S.params, device=device)
        L = E.L
        idbuf = C.create_string_buffer(128)
        if self.rank == 0:
            E._chk(L.wdem_comm_unique_id(idbuf))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda(device); dist.broadcast(t, 0); t = t.cpu()
        else:
            dist.broadcast(t, 0)
        E._chk(L.wdem_comm_init_nccl(E.h, C.c_int(self.rank), C.c_int(self.world), bytes(t.numpy().tobytes())))
        self.cuts = quantile_cuts(S, self.world, axis) if cuts is None else np.asarray(cuts, dtype=np.float64)
        _set_domain(E, axis, self.cuts, getattr(S, "global_max_id", None) or _max_id(S))
        if sel is None:
            sel = np.nonzero(owner_of(S, self.cuts, axis) == self.rank)[0]
        _load_part(E, S, sel)

    def add_spheres(self, ids, pos, radius, mass, inertia, **kw):
        """Collective: the same list on every rank (wdem_add_spheres)."""
        self.E.add_spheres(ids, pos, radius, mass, inertia, **kw)

    def remove_spheres(self, ids):
        """Collective: the same list on every rank (wdem_remove_spheres)."""
        self.E.remove_spheres(ids)

    def set_rebalance(self, tolerance, min_full_runs=1):
        _set_rebalance(self.E, tolerance, min_full_runs)

    def domain(self):
        return _get_domain(self.E, self.world)

    def step(self, n=1):
        return self.E.step(n)

    def owned(self):
        return _owned(self.E)

    def num_owned(self):
        return self.E.L.wdem_num_owned(self.E.h)

    def step_e2e(self, n, ids, pos, vel, angVel, ids_out, pos_out, vel_out, angVel_out):
        """End-to-end step of this subdomain: host state (by id) -> device, n steps, owned state -> host.  Returns (owned count, Ek, stats);
        the owned set (and its order) may change when spheres migrate, so outputs carry their ids."""
        import os, time
        E = self.E
        t0 = time.perf_counter()
        _set_owned(E, ids, pos, vel, angVel)
        t1 = time.perf_counter()
        st = E.step(n)
        t2 = time.perf_counter()
        m = self.num_owned()
        assert m <= len(ids_out), "output buffers too small for the owned set"
        E._chk(E.L.wdem_download_owned(E.h, ptr(ids_out), ptr(pos_out), None, ptr(vel_out), ptr(angVel_out), None, None))
        t3 = time.perf_counter()
        ek = E.kinetic_energy()
        if os.environ.get("WDEM_E2E_TRACE"):
            import sys
            print("rank %d e2e: set_owned %.2f ms, step %.2f ms (device %.2f, rebuilds %d), download %.2f ms, ek %.2f ms" % (
                self.rank, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st.device_ms, st.n_rebuilds, (t3 - t2) * 1e3, (time.perf_counter() - t3) * 1e3), file=sys.stderr, flush=True)
        return m, ek, st

    def step_e2e_owned(self, ids, pos, vel, angVel, pos_out, vel_out, stats=False):
        """The pipelined end-to-end step of this subdomain (wdem_step_e2e_owned): state of the owned spheres by id in, new pos / vel in the same order out."""
        return self.E.step_e2e_owned(ids, pos, vel, angVel, pos_out, vel_out, stats=stats)

    def domain_stats(self):
        return _domain_stats(self.E)
