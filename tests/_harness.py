"""Test harness: ctypes wrapper of the CPU oracle (oracle/libdem_oracle.so) and helpers that feed one
SceneSpec to both the oracle and the CUDA engine.  The oracle is the checker only; nothing in woo_b200/ imports this."""
import ctypes as C
import os, subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libdem_oracle.so")


def build_oracle(variant=""):
    src = os.path.join(ORACLE_DIR, "dem_oracle.cpp")
    so = os.path.join(ORACLE_DIR, "libdem_oracle%s.so" % ("_" + variant if variant else ""))
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("dt", "dtSafety")] + [("gravity", C.c_double * 3)] + \
        [(n, C.c_double) for n in ("distFactor", "damping", "verletDist", "wallHalfThick", "relRollStiff", "relTwistStiff",
                                   "rollTanPhi", "poisson", "en", "maxRelInc")] + [("loneMask", C.c_uint)] + \
        [(n, C.c_int) for n in ("deterministic", "reset", "noRatch", "iniLensTouch", "trsfRenorm", "applyForces", "law", "noSlip",
                                "noBreak", "noFrict", "noAttraction", "dynDtPeriod", "periodic")]


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


_libs = {}


def lib(variant=""):
    """variant "alt": the oracle built with every 3-term reduction order flipped (oracle/Makefile), for tests/test_order_sensitivity.py."""
    if variant not in _libs:
        L = C.CDLL(build_oracle(variant))
        L.orc_create.restype = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        for f in ("orc_get_dt", "orc_get_time", "orc_get_verlet_dist", "orc_pwave_dt", "orc_kinetic_energy", "orc_get_gradv_work"):
            getattr(L, f).restype = C.c_double
        for f in ("orc_get_step", "orc_num_contacts", "orc_get_contacts"):
            getattr(L, f).restype = C.c_long
        _libs[variant] = L
    return _libs[variant]


class Oracle:
    def __init__(self, spec=None, variant=""):
        self.L = lib(variant)
        self.h = C.c_void_p(self.L.orc_create())
        if spec is not None:
            self.load(spec)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h); self.h = None

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())

    def load(self, S):
        L, h = self.L, self.h
        self.spec = S
        for m in S.mats:
            L.orc_add_material(h, C.c_int(m[0]), *[C.c_double(x) for x in m[1:]])
        mass, inertia = S.mass_inertia()

        def add_boundaries():
            for w in S.walls:
                pos = (C.c_double * 3)(*w["pos"]); vel = (C.c_double * 3)(*w.get("vel", (0, 0, 0)))
                L.orc_add_wall(h, pos, C.c_int(w["axis"]), C.c_int(w["sense"]), C.c_int(w.get("mat", 0)), C.c_uint(w.get("mask", 3)), vel)
            for f in S.facets:
                v = np.ascontiguousarray(f["v"], dtype=np.float64).reshape(9)
                fv = np.ascontiguousarray(f.get("fakeVel", (0, 0, 0)), dtype=np.float64)
                nl = np.ascontiguousarray(f.get("n21lim", (np.nan,) * 3), dtype=np.float64)
                vv = np.ascontiguousarray(f.get("vvel", np.zeros(9)), dtype=np.float64).reshape(9)
                L.orc_add_facet(h, _p(v, C.c_double), C.c_double(f.get("halfThick", 0.)), _p(fv, C.c_double), _p(nl, C.c_double),
                                _p(vv, C.c_double), C.c_int(f.get("mat", 0)), C.c_uint(f.get("mask", 3)))
            for g in getattr(S, "facet_groups", []):
                a = lambda k: np.ascontiguousarray(g[k], dtype=np.float64)
                ix = np.ascontiguousarray(g["facets"], dtype=np.int32)
                self._chk(L.orc_add_facet_group(h, _p(a("pos"), C.c_double), _p(a("ori"), C.c_double), _p(a("vel"), C.c_double), _p(a("angVel"), C.c_double),
                                                C.c_int(len(ix)), _p(ix, C.c_int)))

        def add_spheres():
            if S.n:
                L.orc_add_spheres(h, C.c_int(S.n), _p(S.pos, C.c_double), _p(S.radius, C.c_double), _p(mass, C.c_double),
                                  _p(inertia, C.c_double), _p(S.mat, C.c_int), _p(S.mask, C.c_uint), _p(S.flags, C.c_uint),
                                  _p(S.vel, C.c_double), _p(S.angVel, C.c_double), None)
        if S.boundaries_first:
            add_boundaries(); add_spheres()
        else:
            add_spheres(); add_boundaries()
        ids = S.sphere_ids()
        for d in getattr(S, "imposes", []):
            nt = len(d["times"])
            L.orc_add_impose(h, C.c_int(d["kind"]), C.c_int(d["flags"]), _p(np.ascontiguousarray(d["v"]), C.c_double), _p(np.ascontiguousarray(d["ori"]), C.c_double), C.c_int(nt),
                             _p(np.ascontiguousarray(d["times"]), C.c_double) if nt else None, _p(np.ascontiguousarray(d["vals"]), C.c_double) if nt else None,
                             _p(np.ascontiguousarray(d["vals2"]), C.c_double) if len(d["vals2"]) else None)
        for st in getattr(S, "impose_sets", []):
            a = np.ascontiguousarray(st, dtype=np.int32); L.orc_add_impose_set(h, C.c_int(len(a)), _p(a, C.c_int))
        for kind, arr, idof in ((0, getattr(S, "sphere_impose", None), lambda i: int(ids[i])), (1, getattr(S, "wall_impose", None), lambda i: int(S.wall_ids()[i])), (2, getattr(S, "group_impose", None), lambda i: i),
                               (3, getattr(S, "facet_impose", None), lambda i: int(S.facet_ids()[i]))):
            if arr is None:
                continue
            for i, k in enumerate(arr):
                if k >= 0:
                    self._chk(L.orc_attach_impose(h, C.c_int(kind), C.c_int(idof(i)), C.c_int(int(k))))
        for c in getattr(S, "clumps", []):
            a = lambda k: np.ascontiguousarray(c[k], dtype=np.float64)
            mem = np.ascontiguousarray(ids[np.asarray(c["members"], dtype=np.int64)], dtype=np.int32)
            self._chk(L.orc_add_clump(h, C.c_int(len(mem)), _p(mem, C.c_int), _p(a("pos"), C.c_double), _p(a("ori"), C.c_double), _p(a("vel"), C.c_double), _p(a("angVel"), C.c_double),
                                      C.c_double(c["mass"]), _p(a("inertia"), C.c_double), C.c_uint(int(c.get("flags", 0)))))
        if getattr(S, "clump_impose", None) is not None:      # impositions on clump master nodes: after the clumps exist
            for i, k in enumerate(S.clump_impose):
                if k >= 0:
                    self._chk(L.orc_attach_impose(h, C.c_int(4), C.c_int(i), C.c_int(int(k))))
        P = OrcParams(); L.orc_default_params(C.byref(P))
        for k, v in vars(S.params).items():
            if k == "gravity":
                P.gravity = (C.c_double * 3)(*v)
            else:
                setattr(P, k, v)
        L.orc_set_params(h, C.byref(P))
        if S.cell is not None:
            hs = np.ascontiguousarray(S.cell["hSize"], dtype=np.float64).reshape(9)
            gv = np.ascontiguousarray(S.cell.get("gradV", np.zeros(9)), dtype=np.float64).reshape(9)
            ng = np.ascontiguousarray(S.cell.get("nextGradV", gv), dtype=np.float64).reshape(9)
            L.orc_set_cell(h, _p(hs, C.c_double), _p(gv, C.c_double), _p(ng, C.c_double), C.c_int(S.cell.get("homoDeform", 4)))

    def step(self, n=1):
        self._chk(self.L.orc_step(self.h, C.c_int(n)))

    def collider_run(self):
        self._chk(self.L.orc_collider_run(self.h))

    @property
    def dt(self): return self.L.orc_get_dt(self.h)
    @property
    def nstep(self): return self.L.orc_get_step(self.h)
    @property
    def time(self): return self.L.orc_get_time(self.h)
    @property
    def verletDist(self): return self.L.orc_get_verlet_dist(self.h)
    @property
    def nFullRuns(self): return self.L.orc_get_nfullruns(self.h)
    def pWaveDt(self): return self.L.orc_pwave_dt(self.h)
    def kineticEnergy(self): return self.L.orc_kinetic_energy(self.h)

    def walls(self):
        n = len(self.spec.walls)
        o = {k: np.zeros((n, 3)) for k in ("pos", "force", "torque")}
        if n:
            self.L.orc_get_walls(self.h, _p(o["pos"], C.c_double), _p(o["force"], C.c_double), _p(o["torque"], C.c_double))
        return o

    def add_spheres(self, pos, radius, mass, inertia, mat, mask=None, flags=None, vel=None, angVel=None):
        """ParticleContainer.add in mid-run; returns the id of the first new particle (ids continue after everything added so far)."""
        a = lambda x, dt: None if x is None else np.ascontiguousarray(x, dtype=dt)
        pos, radius, mass, inertia, vel, angVel = (a(x, np.float64) for x in (pos, radius, mass, inertia, vel, angVel))
        mat = a(mat, np.int32); mask = a(mask, np.uint32); flags = a(flags, np.uint32)
        return self.L.orc_add_spheres(self.h, C.c_int(len(radius)), _p(pos, C.c_double), _p(radius, C.c_double), _p(mass, C.c_double), _p(inertia, C.c_double),
                                      _p(mat, C.c_int), _p(mask, C.c_uint), _p(flags, C.c_uint), _p(vel, C.c_double), _p(angVel, C.c_double), None)

    def last_added_ids(self, n):
        """Ids ParticleContainer.add handed out in the last add_spheres (freed ids are reused oldest-first, ParticleContainer.cpp:23-36)."""
        o = np.zeros(n, np.int32)
        assert self.L.orc_last_added_ids(self.h, _p(o, C.c_int)) == n
        return o

    def sphere_ids(self):
        o = np.zeros(self.L.orc_num_spheres(self.h), np.int32)
        self.L.orc_get_sphere_ids(self.h, _p(o, C.c_int))
        return o

    def remove_spheres(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._chk(self.L.orc_remove_spheres(self.h, C.c_int(len(ids)), _p(ids, C.c_int)))

    def impose_state(self, ix):
        """sumF, dist (VelocityAndReadForce), F, T (ReadForce) of impose descriptor ix"""
        o = np.zeros(8); self.L.orc_get_impose_state(self.h, C.c_int(ix), _p(o, C.c_double))
        return dict(sumF=o[0], dist=o[1], F=o[2:5].copy(), T=o[5:8].copy())

    def clumps(self):
        n = len(getattr(self.spec, "clumps", []))
        o = dict(pos=np.zeros((n, 3)), ori=np.zeros((n, 4)), vel=np.zeros((n, 3)), angVel=np.zeros((n, 3)))
        if n:
            self.L.orc_get_clumps(self.h, _p(o["pos"], C.c_double), _p(o["ori"], C.c_double), _p(o["vel"], C.c_double), _p(o["angVel"], C.c_double))
        return o

    def facets(self):
        o = np.zeros((len(self.spec.facets), 3, 3))
        if len(o):
            self.L.orc_get_facets(self.h, _p(o, C.c_double))
        return o

    def track_energy(self, on=True):
        self.L.orc_set_track_energy(self.h, C.c_int(int(on)))

    def energy(self):
        o = np.zeros(8)
        self.L.orc_get_energy(self.h, _p(o, C.c_double))
        return dict(zip(("kinetic", "grav", "nonviscDamp", "elast", "plast", "broken", "viscN", "viscT"), o.tolist()))

    def eval_stress(self, on=True):
        """ContactLoop.evalStress: with track_energy in a periodic scene the gradV work is booked every step (ContactLoop.cpp:243-254)."""
        self.L.orc_set_eval_stress(self.h, C.c_int(int(on)))

    def gradv_work(self):
        return self.L.orc_get_gradv_work(self.h)

    def stress(self):
        o = np.zeros((3, 3))
        self._chk(self.L.orc_get_stress(self.h, _p(o, C.c_double)))
        return o

    def cell_hsize(self):
        h = np.zeros((3, 3))
        self.L.orc_get_cell(self.h, _p(h, C.c_double))
        return h

    def spheres(self):
        n = self.L.orc_num_spheres(self.h)
        out = {k: np.zeros((n, 3)) for k in ("pos", "vel", "angVel", "force", "torque")}
        out["ori"] = np.zeros((n, 4))
        self.L.orc_get_spheres(self.h, _p(out["pos"], C.c_double), _p(out["ori"], C.c_double), _p(out["vel"], C.c_double),
                               _p(out["angVel"], C.c_double), _p(out["force"], C.c_double), _p(out["torque"], C.c_double))
        return out

    def set_sphere(self, idx, pos=None, vel=None, angVel=None):
        a = lambda v: None if v is None else (C.c_double * 3)(*v)
        self.L.orc_set_sphere_state(self.h, C.c_int(idx), a(pos), a(vel), a(angVel))

    def wall_forces(self):
        out = np.zeros((len(self.spec.walls), 3))
        if len(out):
            self.L.orc_get_wall_forces(self.h, _p(out, C.c_double))
        return out

    def num_contacts(self, real_only=False):
        return self.L.orc_num_contacts(self.h, C.c_int(int(real_only)))

    def contacts(self, real_only=False):
        n = self.num_contacts(real_only)
        o = dict(ids=np.zeros((n, 2), np.int32), cellDist=np.zeros((n, 3), np.int32), real=np.zeros(n, np.int32), uN=np.zeros(n),
                 force=np.zeros((n, 3)), torque=np.zeros((n, 3)), trsf=np.zeros((n, 9)), contPt=np.zeros((n, 3)), phys=np.zeros((n, 3)),
                 vel=np.zeros((n, 3)), stepCreated=np.zeros(n, np.int32))
        if n:
            self.L.orc_get_contacts(self.h, C.c_int(int(real_only)), _p(o["ids"], C.c_int), _p(o["cellDist"], C.c_int), _p(o["real"], C.c_int),
                                    _p(o["uN"], C.c_double), _p(o["force"], C.c_double), _p(o["torque"], C.c_double), _p(o["trsf"], C.c_double),
                                    _p(o["contPt"], C.c_double), _p(o["phys"], C.c_double), _p(o["vel"], C.c_double), _p(o["stepCreated"], C.c_int))
        return o
