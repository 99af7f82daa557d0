"""Host-side handle of the CUDA DEM engine over the C-ABI (array level).  The reference-facing plugin layer
(GpuDemLoop engine + Scene/DemField mirror) sits on top of this in woo_b200/dem.py."""
import ctypes as C
import numpy as np
from . import capi
from .capi import WdemConfig, WdemStats, ptr, f64


class WdemError(RuntimeError):
    pass


class Engine:
    """One DemField on one GPU.  Mirrors the C-ABI 1:1; arrays in, arrays out."""

    def __init__(self, params=None, device=0, variant="", **kw):
        self.L = capi.load(variant=variant)
        self.cfg = WdemConfig()
        self.L.wdem_default_config(C.byref(self.cfg))
        self.cfg.device = device
        if params is not None:
            self._apply(vars(params))
        self._apply(kw)
        self.h = C.c_void_p()
        rc = self.L.wdem_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise WdemError(self.L.wdem_last_error(None).decode())
        self._keep = []

    def _apply(self, d):
        for k, v in d.items():
            if k == "gravity":
                self.cfg.gravity = (C.c_double * 3)(*v)
            elif k == "deterministic":
                continue        # the GPU force reduction is always deterministic (S.deterministic semantics)
            elif hasattr(self.cfg, k):
                setattr(self.cfg, k, v)
            else:
                raise AttributeError("unknown engine attribute %r" % k)

    def _chk(self, rc):
        if rc != 0:
            raise WdemError(self.L.wdem_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.wdem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_config(self, **kw):
        self._apply(kw)
        self._chk(self.L.wdem_set_config(self.h, C.byref(self.cfg)))

    def set_profiling(self, on=True):
        self._chk(self.L.wdem_set_profiling(self.h, C.c_int(int(on))))

    # ---- uploads
    def upload_materials(self, mats):
        a = np.array(mats, dtype=np.float64).reshape(len(mats), 7)
        kind = np.ascontiguousarray(a[:, 0], dtype=np.int32)
        cols = [np.ascontiguousarray(a[:, i]) for i in range(1, 7)]
        self._chk(self.L.wdem_upload_materials(self.h, C.c_int(len(mats)), ptr(kind), *[ptr(c) for c in cols]))

    def upload_spheres(self, ids, pos, radius, mass, inertia, mat=None, mask=None, flags=None, vel=None, angVel=None, ori=None):
        n = len(radius)
        self.n = n
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int32)
        mat = None if mat is None else np.ascontiguousarray(mat, dtype=np.int32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        flags = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint32)
        args = [f64(pos), f64(ori), f64(vel), f64(angVel), f64(mass), f64(inertia), f64(radius)]
        self._chk(self.L.wdem_upload_spheres(self.h, C.c_int64(n), ptr(ids), *[ptr(a) for a in args], ptr(flags), ptr(mask), ptr(mat)))

    def add_spheres(self, ids, pos, radius, mass, inertia, mat=None, mask=None, flags=None, vel=None, angVel=None, ori=None):
        """Append spheres in mid-run (wdem.h: wdem_add_spheres)."""
        m = len(radius)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        mat = None if mat is None else np.ascontiguousarray(mat, dtype=np.int32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        flags = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint32)
        args = [f64(pos), f64(ori), f64(vel), f64(angVel), f64(mass), f64(inertia), f64(radius)]
        self._chk(self.L.wdem_add_spheres(self.h, C.c_int64(m), ptr(ids), *[ptr(a) for a in args], ptr(flags), ptr(mask), ptr(mat)))
        self.n += m

    def remove_spheres(self, ids):
        """Remove spheres by Particle.id in mid-run (wdem.h: wdem_remove_spheres)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._chk(self.L.wdem_remove_spheres(self.h, C.c_int64(len(ids)), ptr(ids)))
        self.n -= len(ids)

    def upload_walls(self, ids, walls):
        n = len(walls)
        if n == 0:
            return
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        pos = f64([w["pos"] for w in walls]); vel = f64([w.get("vel", (0, 0, 0)) for w in walls])
        axis = np.array([w["axis"] for w in walls], dtype=np.int32); sense = np.array([w["sense"] for w in walls], dtype=np.int32)
        mask = np.array([w.get("mask", 3) for w in walls], dtype=np.uint32); mat = np.array([w.get("mat", 0) for w in walls], dtype=np.int32)
        self._chk(self.L.wdem_upload_walls(self.h, C.c_int(n), ptr(ids), ptr(pos), ptr(axis), ptr(sense), ptr(vel), ptr(mask), ptr(mat)))

    def upload_facets(self, ids, facets):
        n = len(facets)
        if n == 0:
            return
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        v = f64([np.asarray(f["v"]).reshape(9) for f in facets]); vv = f64([np.asarray(f.get("vvel", np.zeros(9))).reshape(9) for f in facets])
        fk = f64([f.get("fakeVel", (0, 0, 0)) for f in facets]); ht = f64([f.get("halfThick", 0.) for f in facets])
        nl = f64([f.get("n21lim", (np.nan,) * 3) for f in facets])
        mask = np.array([f.get("mask", 3) for f in facets], dtype=np.uint32); mat = np.array([f.get("mat", 0) for f in facets], dtype=np.int32)
        self._chk(self.L.wdem_upload_facets(self.h, C.c_int(n), ptr(ids), ptr(v), ptr(vv), ptr(fk), ptr(ht), ptr(nl), ptr(mask), ptr(mat)))

    def upload_facet_groups(self, n_facets, groups):
        """Kinematic groups of facets (wdem.h: wdem_upload_facet_groups): dicts with pos, ori (x,y,z,w), vel, angVel, facets (indices)."""
        if not groups:
            return
        of = np.full(n_facets, -1, np.int32)
        for gi, g in enumerate(groups):
            of[np.asarray(g["facets"], dtype=np.int64)] = gi
        arr = lambda k, w: f64([g[k] for g in groups], (len(groups), w))
        self._chk(self.L.wdem_upload_facet_groups(self.h, C.c_int(len(groups)), ptr(of), ptr(arr("pos", 3)), ptr(arr("ori", 4)), ptr(arr("vel", 3)), ptr(arr("angVel", 3))))

    def upload_imposes(self, S):
        """DemData.impose of spheres, walls and facet groups of a SceneSpec (wdem.h: wdem_upload_imposes)."""
        from .capi import WdemImpose
        imps = getattr(S, "imposes", [])
        if not imps:
            return
        keep = []      # arrays referenced by pointer from the descriptors
        arr = (WdemImpose * len(imps))()
        for a, d in zip(arr, imps):
            a.kind, a.flags = int(d["kind"]), int(d["flags"])
            a.v[:] = [float(x) for x in d["v"]]; a.ori[:] = [float(x) for x in d["ori"]]
            a.nTimes = len(d["times"])
            for name in ("times", "vals", "vals2"):
                if len(d[name]):
                    b = np.ascontiguousarray(d[name], dtype=np.float64); keep.append(b); setattr(a, name, b.ctypes.data)
        start = np.zeros(len(S.impose_sets) + 1, np.int32); start[1:] = np.cumsum([len(s) for s in S.impose_sets])
        ops = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int32) for s in S.impose_sets]), dtype=np.int32)
        sid = np.ascontiguousarray(S.sphere_ids(), dtype=np.int32)
        sset = None if S.sphere_impose is None else np.ascontiguousarray(S.sphere_impose, dtype=np.int32)
        wset = None if S.wall_impose is None else np.ascontiguousarray(S.wall_impose, dtype=np.int32)
        gset = None if S.group_impose is None else np.ascontiguousarray(S.group_impose, dtype=np.int32)
        self._chk(self.L.wdem_upload_imposes(self.h, C.c_int32(len(imps)), arr, C.c_int32(len(S.impose_sets)), ptr(start), ptr(ops), C.c_int64(len(sid) if sset is not None else 0),
                                             ptr(sid), ptr(sset), ptr(wset), ptr(gset)))
        cset = getattr(S, "clump_impose", None)
        if cset is not None:
            self._chk(self.L.wdem_upload_clump_imposes(self.h, ptr(np.ascontiguousarray(cset, dtype=np.int32))))
        fset = getattr(S, "facet_impose", None)
        if fset is not None:
            self._chk(self.L.wdem_upload_facet_imposes(self.h, ptr(np.ascontiguousarray(fset, dtype=np.int32))))

    def impose_state(self, ix):
        """sumF, dist (VelocityAndReadForce), F, T (ReadForce) of Impose object ix (wdem.h: wdem_get_impose_state)."""
        o = np.zeros(8)
        self._chk(self.L.wdem_get_impose_state(self.h, C.c_int32(ix), ptr(o)))
        return dict(sumF=o[0], dist=o[1], F=o[2:5].copy(), T=o[5:8].copy())

    def upload_clumps(self, sphere_ids, clumps):
        """Clumps of spheres with mass (wdem.h: wdem_upload_clumps): dicts with members (indices into the uploaded spheres), pos, ori (x,y,z,w), vel, angVel, mass,
        inertia (3), flags and optionally angMom, relPos, relOri."""
        if not clumps:
            return
        ids = np.asarray(sphere_ids, dtype=np.int64)
        start = np.zeros(len(clumps) + 1, np.int32); start[1:] = np.cumsum([len(c["members"]) for c in clumps])
        mem = np.ascontiguousarray(np.concatenate([ids[np.asarray(c["members"], dtype=np.int64)] for c in clumps]), dtype=np.int32)
        arr = lambda k, w: f64([c[k] for c in clumps], (len(clumps), w))
        opt = lambda k, w: arr(k, w) if all(k in c for c in clumps) else None
        cat = lambda k, w: f64(np.concatenate([np.asarray(c[k], dtype=np.float64).reshape(-1, w) for c in clumps])) if all(k in c for c in clumps) else None
        fl = np.array([int(c.get("flags", 0)) for c in clumps], dtype=np.uint32)
        self.n_clumps = len(clumps)
        self._chk(self.L.wdem_upload_clumps(self.h, C.c_int32(len(clumps)), ptr(start), ptr(mem), ptr(arr("pos", 3)), ptr(arr("ori", 4)), ptr(arr("vel", 3)), ptr(arr("angVel", 3)),
                                            ptr(opt("angMom", 3)), ptr(f64([c["mass"] for c in clumps])), ptr(arr("inertia", 3)), ptr(fl), ptr(cat("relPos", 3)), ptr(cat("relOri", 4))))

    def clumps(self):
        """State of the clump master nodes (wdem.h: wdem_download_clumps)."""
        n = int(self.L.wdem_num_clumps(self.h))
        o = dict(pos=np.zeros((n, 3)), ori=np.zeros((n, 4)), vel=np.zeros((n, 3)), angVel=np.zeros((n, 3)), angMom=np.zeros((n, 3)), force=np.zeros((n, 3)), torque=np.zeros((n, 3)))
        if n:
            self._chk(self.L.wdem_download_clumps(self.h, *[ptr(o[k]) for k in ("pos", "ori", "vel", "angVel", "angMom", "force", "torque")]))
        return o

    def facets(self):
        """Facet node positions after the last step, (n,3,3)."""
        n = len(self.spec.facets) if getattr(self, "spec", None) is not None else 0
        o = np.zeros((n, 3, 3))
        if n:
            self._chk(self.L.wdem_download_facets(self.h, ptr(o), None, None))
        return o

    def set_cell(self, hSize, gradV=None, nextGradV=None, homoDeform=4):
        self._chk(self.L.wdem_set_cell(self.h, ptr(f64(hSize)), ptr(f64(gradV)), ptr(f64(nextGradV)), C.c_int(homoDeform)))

    def load(self, S):
        """Upload a woo_b200.scenes.SceneSpec."""
        self.spec = S
        self.upload_materials(S.mats)
        m, I = S.mass_inertia()
        self.upload_spheres(S.sphere_ids(), S.pos, S.radius, m, I, S.mat, S.mask, S.flags, S.vel, S.angVel)
        self.upload_walls(S.wall_ids(), S.walls)
        self.upload_facets(S.facet_ids(), S.facets)
        self.upload_clumps(S.sphere_ids(), getattr(S, "clumps", []))
        self.upload_facet_groups(len(S.facets), getattr(S, "facet_groups", []))
        if S.cell is not None:
            self.set_cell(S.cell["hSize"], S.cell.get("gradV"), S.cell.get("nextGradV"), S.cell.get("homoDeform", 4))
        self.upload_imposes(S)
        return self

    # ---- stepping
    def step(self, n=1):
        st = WdemStats()
        self._chk(self.L.wdem_step(self.h, C.c_int64(n), C.byref(st)))
        self.stats = st
        return st

    def step_e2e(self, n, pos, vel, angVel, pos_out, vel_out, stats=True):
        """stats=False: wdem_step_e2e with stats == NULL, i.e. the call returns when the outputs are on the host while the contact pass
        for the next step is still running (include/wdem.h)."""
        st = WdemStats() if stats else None; ek = C.c_double()
        self._chk(self.L.wdem_step_e2e(self.h, C.c_int64(n), ptr(pos), ptr(vel), ptr(angVel), ptr(pos_out), ptr(vel_out), C.byref(ek), C.byref(st) if stats else None))
        self.stats = st
        return ek.value, st

    def step_e2e_owned(self, ids, pos, vel, angVel, pos_out, vel_out, stats=False):
        """wdem_step_e2e_owned: one end-to-end step of the spheres this handle owns, addressed by id; returns (Ek, owned_changed, stats)."""
        st = WdemStats() if stats else None; ek = C.c_double(); ch = C.c_int32()
        self._chk(self.L.wdem_step_e2e_owned(self.h, C.c_int64(len(ids)), ptr(ids), ptr(pos), ptr(vel), ptr(angVel), ptr(pos_out), ptr(vel_out), C.byref(ek), C.byref(ch),
                                             C.byref(st) if stats else None))
        return ek.value, bool(ch.value), st

    # ---- downloads
    def spheres(self):
        n = self.n
        o = {k: np.zeros((n, 3)) for k in ("pos", "vel", "angVel", "force", "torque")}
        o["ori"] = np.zeros((n, 4))
        self._chk(self.L.wdem_download_state(self.h, ptr(o["pos"]), ptr(o["ori"]), ptr(o["vel"]), ptr(o["angVel"])))
        self._chk(self.L.wdem_download_forces(self.h, ptr(o["force"]), ptr(o["torque"])))
        return o

    def set_state(self, first, pos=None, vel=None, angVel=None):
        a = [f64(x) for x in (pos, vel, angVel)]
        n = next(len(x) for x in a if x is not None)
        self._chk(self.L.wdem_set_state(self.h, C.c_int64(first), C.c_int64(n), *[ptr(x) for x in a]))

    def num_contacts(self, real_only=False):
        n = self.L.wdem_num_contacts(self.h, C.c_int(int(real_only)))
        if n < 0:
            self._chk(1)
        return n

    def contacts(self, real_only=False):
        n = self.L.wdem_download_contacts(self.h, C.c_int(int(real_only)), C.c_int64(0), *([None] * 11))
        if n < 0:
            self._chk(1)
        o = dict(ids=np.zeros((n, 2), np.int32), cellDist=np.zeros((n, 3), np.int32), real=np.zeros(n, np.int32), uN=np.zeros(n),
                 force=np.zeros((n, 3)), torque=np.zeros((n, 3)), trsf=np.zeros((n, 9)), contPt=np.zeros((n, 3)), phys=np.zeros((n, 3)),
                 stepCreated=np.zeros(n, np.int32), physRaw=np.zeros((n, 5)))
        if n:
            m = self.L.wdem_download_contacts(self.h, C.c_int(int(real_only)), C.c_int64(n), ptr(o["ids"]), ptr(o["cellDist"]), ptr(o["real"]),
                                              ptr(o["uN"]), ptr(o["force"]), ptr(o["torque"]), ptr(o["trsf"]), ptr(o["contPt"]), ptr(o["phys"]),
                                              ptr(o["stepCreated"]), ptr(o["physRaw"]))
            if m < 0:
                self._chk(1)
        return o

    def upload_contacts(self, c):
        """Import the history of real contacts as returned by contacts(True) (resume / re-upload); before the first step."""
        m = np.asarray(c["real"]).astype(bool) if "real" in c else np.ones(len(c["ids"]), bool)
        g = lambda k, dt: np.ascontiguousarray(np.asarray(c[k])[m], dtype=dt)
        ids, cd, sc = g("ids", np.int32), g("cellDist", np.int32), g("stepCreated", np.int32)
        uN, f, t, tr, cp, ph = (g(k, np.float64) for k in ("uN", "force", "torque", "trsf", "contPt", "physRaw"))
        self._chk(self.L.wdem_upload_contacts(self.h, C.c_int64(len(ids)), ptr(ids), ptr(cd), ptr(sc), ptr(uN), ptr(f), ptr(t), ptr(tr), ptr(cp), ptr(ph)))

    def set_scene(self, step, time):
        self._chk(self.L.wdem_set_scene(self.h, C.c_int64(int(step)), C.c_double(float(time))))

    def periodic_state(self):
        """(stepLast, stepPrev) of DynDt's PeriodicEngine counters."""
        a, b = C.c_int64(), C.c_int64()
        self._chk(self.L.wdem_get_periodic_state(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_periodic_state(self, step_last, step_prev):
        self._chk(self.L.wdem_set_periodic_state(self.h, C.c_int64(int(step_last)), C.c_int64(int(step_prev))))

    def walls(self):
        """pos / force / torque of the walls after the last step (force: the last step's contacts)."""
        n = len(self.spec.walls) if getattr(self, "spec", None) is not None else 0
        o = {k: np.zeros((n, 3)) for k in ("pos", "force", "torque")}
        if n:
            self._chk(self.L.wdem_download_walls(self.h, ptr(o["pos"]), ptr(o["force"]), ptr(o["torque"])))
        return o

    ENERGY_NAMES = ("kinetic", "grav", "nonviscDamp", "elast", "plast", "broken", "viscN", "viscT")

    def track_energy(self, on=True):
        self._chk(self.L.wdem_track_energy(self.h, C.c_int(int(on))))

    def energy(self):
        """Scene.energy entries of this path as a dict (wdem.h: wdem_get_energy)."""
        o = np.zeros(8)
        self._chk(self.L.wdem_get_energy(self.h, ptr(o)))
        return dict(zip(self.ENERGY_NAMES, o.tolist()))

    def gradv_work(self):
        """EnergyTracker entry "gradV" (wdem.h: wdem_get_gradv_work)."""
        w = C.c_double()
        self._chk(self.L.wdem_get_gradv_work(self.h, C.byref(w)))
        return w.value

    def stress(self):
        """ContactLoop.stress (evalStress) of the current state, 3x3."""
        o = np.zeros((3, 3))
        self._chk(self.L.wdem_get_stress(self.h, ptr(o)))
        return o

    def kinetic_energy(self):
        e = C.c_double()
        self._chk(self.L.wdem_kinetic_energy(self.h, C.byref(e)))
        return e.value

    def cell_hsize(self):
        h = np.zeros((3, 3))
        self._chk(self.L.wdem_get_cell(self.h, ptr(h)))
        return h

    def crit_dt(self):
        e = C.c_double()
        self._chk(self.L.wdem_crit_dt(self.h, C.byref(e)))
        return e.value

    def scalars(self):
        dt, t, vd = C.c_double(), C.c_double(), C.c_double(); st, nf = C.c_int64(), C.c_int64()
        self._chk(self.L.wdem_get_scalars(self.h, C.byref(dt), C.byref(t), C.byref(st), C.byref(vd), C.byref(nf)))
        return dict(dt=dt.value, time=t.value, step=st.value, verletDist=vd.value, nFullRuns=nf.value)

    def device_info(self):
        b = C.c_int64(); s = C.c_void_p()
        self._chk(self.L.wdem_device_info(self.h, C.byref(b), C.byref(s)))
        return dict(bytes=b.value, stream=s.value)
