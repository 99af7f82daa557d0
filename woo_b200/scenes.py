"""Synthetic scene generators and the plain-array scene description shared by the
product host layer, the tests and bench.py.  Our own seeded generator, not Woo's (py/pack.py is out of scope).

Mass/inertia follow woo::Sphere::lumpMassInertia (src/dem/Sphere.cpp:26-33):
m = 4/3 π r³ ρ, I = 2/5 m r².  Default masks follow src/dem/Particle.hpp:339-349
(spheres 0b0101, walls/facets 0b0011, loneMask 0b0010).
"""
from __future__ import annotations
import math
from dataclasses import dataclass, field
import numpy as np

MAT_FRICT, MAT_HERTZ = 0, 1
LAW_IDEALELPL, LAW_HERTZ_DMT = 0, 1
HOMO_NONE, HOMO_POS, HOMO_VEL, HOMO_VEL_2ND, HOMO_GRADV2 = 0, 1, 2, 3, 4
MASK_SPHERE, MASK_BOUNDARY, LONE_MASK = 0b0101, 0b0011, 0b0010
# DemData flags (src/dem/Particle.hpp:198-203)
DOF_ALL, GRAVITY_SKIP, DAMPING_SKIP = 63, 512, 2048


@dataclass
class Params:
    """Engine attributes, named as in the reference."""
    dt: float = float("nan")
    dtSafety: float = 0.9
    gravity: tuple = (0.0, 0.0, 0.0)
    loneMask: int = LONE_MASK
    distFactor: float = -1.0
    deterministic: int = 1
    damping: float = 0.2          # Leapfrog.damping
    reset: int = 1                # Leapfrog.reset
    verletDist: float = -0.05     # InsertionSortCollider.verletDist
    wallHalfThick: float = 0.0    # Bo1_Wall_Aabb.halfThick
    noRatch: int = 0              # Cg2_*.noRatch
    iniLensTouch: int = 1
    trsfRenorm: int = 100
    applyForces: int = 1          # ContactLoop.applyForces
    law: int = LAW_IDEALELPL
    noSlip: int = 0
    noBreak: int = 0
    noFrict: int = 0
    relRollStiff: float = 0.0
    relTwistStiff: float = 0.0
    rollTanPhi: float = 0.0
    poisson: float = 0.2          # Cp2_HertzMat_HertzPhys.poisson
    en: float = float("nan")      # Cp2_HertzMat_HertzPhys.en
    noAttraction: int = 1         # Law2_L6Geom_HertzPhys_DMT.noAttraction
    dynDtPeriod: int = 100        # DynDt.stepPeriod (<=0: no DynDt)
    maxRelInc: float = 1e-4
    periodic: int = 0


@dataclass
class SceneSpec:
    """Particles are numbered in the order spheres, walls, facets unless ``order`` says otherwise
    (ids are what S.dem.par.add would hand out)."""
    mats: list = field(default_factory=list)   # (kind, density, young, ktDivKn, tanPhi, surfEnergy, alpha)
    pos: np.ndarray = None
    radius: np.ndarray = None
    vel: np.ndarray = None
    angVel: np.ndarray = None
    mat: np.ndarray = None
    mask: np.ndarray = None
    flags: np.ndarray = None
    walls: list = field(default_factory=list)  # dict(pos, axis, sense, mat, mask, vel)
    facets: list = field(default_factory=list)  # dict(v(3,3), halfThick, fakeVel, n21lim, vvel, mat, mask)
    # Impose (Particle.hpp:130-160, Impose.hpp): descriptors (imp_* helpers below), sets of descriptor indices (CombinedImpose), and per node the index of its set (-1: none)
    imposes: list = field(default_factory=list)
    impose_sets: list = field(default_factory=list)
    sphere_impose: object = None        # int array per sphere (index into impose_sets), or None
    wall_impose: object = None          # int array per wall
    group_impose: object = None         # int array per facet group
    facet_impose: object = None         # int array per facet: the set all three nodes of the facet carry
    clump_impose: object = None         # int array per clump of spheres: the set on its master node
    clumps: list = field(default_factory=list)  # rigid aggregates of spheres with mass: dict(members=[sphere indices], pos, ori (x,y,z,w), vel, angVel, mass, inertia(3), flags); see make_clumps
    facet_groups: list = field(default_factory=list)  # kinematic groups: dict(pos, ori (x,y,z,w), vel, angVel, facets=[indices into ``facets``])
    params: Params = field(default_factory=Params)
    cell: dict = None                           # dict(hSize(3,3), gradV, nextGradV, homoDeform)
    boundaries_first: bool = False              # ids: walls+facets first, then spheres
    ids: np.ndarray = None                      # explicit sphere ids (a part of a larger scene; needs boundaries_first)
    global_max_id: int = None                   # largest Particle.id of the whole scene when this spec holds only a part

    @property
    def n(self):
        return 0 if self.pos is None else len(self.pos)

    def finalize(self):
        n = self.n
        self.pos = np.ascontiguousarray(self.pos, dtype=np.float64).reshape(n, 3)
        self.radius = np.ascontiguousarray(self.radius, dtype=np.float64).reshape(n)
        self.vel = np.zeros((n, 3)) if self.vel is None else np.ascontiguousarray(self.vel, dtype=np.float64).reshape(n, 3)
        self.angVel = np.zeros((n, 3)) if self.angVel is None else np.ascontiguousarray(self.angVel, dtype=np.float64).reshape(n, 3)
        self.mat = np.zeros(n, dtype=np.int32) if self.mat is None else np.ascontiguousarray(self.mat, dtype=np.int32)
        self.mask = np.full(n, MASK_SPHERE, dtype=np.uint32) if self.mask is None else np.ascontiguousarray(self.mask, dtype=np.uint32)
        self.flags = np.zeros(n, dtype=np.uint32) if self.flags is None else np.ascontiguousarray(self.flags, dtype=np.uint32)
        return self

    def mass_inertia(self):
        dens = np.array([m[1] for m in self.mats], dtype=np.float64)[self.mat]
        r = self.radius
        m = (4 / 3.) * math.pi * (r * r * r) * dens     # pow3(x)=x*x*x (Math.hpp:162), not pow()
        I = (2. / 5.) * m * (r * r)
        return m, I

    def sphere_ids(self):
        if self.ids is not None:
            return np.ascontiguousarray(self.ids, dtype=np.int32)
        nb = len(self.walls) + len(self.facets)
        first = nb if self.boundaries_first else 0
        return np.arange(first, first + self.n, dtype=np.int32)

    def wall_ids(self):
        first = 0 if self.boundaries_first else self.n
        return np.arange(first, first + len(self.walls), dtype=np.int32)

    def facet_ids(self):
        first = (0 if self.boundaries_first else self.n) + len(self.walls)
        return np.arange(first, first + len(self.facets), dtype=np.int32)


def frict_mat(density=1e3, young=1e7, ktDivKn=.2, tanPhi=math.tan(.5)):
    """woo.utils.defaultMaterial (py/utils.py:88-92)."""
    return (MAT_FRICT, density, young, ktDivKn, tanPhi, 0.0, 0.0)


def hertz_mat(density=1e3, young=1e8, ktDivKn=.2, tanPhi=.5, surfEnergy=0.0, alpha=0.0):
    return (MAT_HERTZ, density, young, ktDivKn, tanPhi, surfEnergy, alpha)


# ---- Impose descriptors (wdem.h: wdem_impose; kinds and the meaning of v[16] follow include/wdem.h)
IMP_HARMONIC, IMP_ALIGNED_HARMONIC, IMP_CONSTANT_FORCE, IMP_RADIAL_FORCE, IMP_LOCAL6, IMP_VEL_READ_FORCE, IMP_READ_FORCE, IMP_VARIABLE_VEL3D, IMP_VARIABLE_ALIGNED_ROTATION = range(1, 10)
IMP_CIRCULAR_ORBIT, IMP_STABLE_CIRCULAR_ORBIT, IMP_INTERPOLATED_MOTION = 10, 11, 12


def _imp(kind, flags=0, v=(), ori=(0, 0, 0, 1), times=(), vals=(), vals2=()):
    vv = np.zeros(16); vv[:len(v)] = v
    return dict(kind=kind, flags=int(flags), v=vv, ori=np.asarray(ori, dtype=np.float64), times=np.asarray(times, dtype=np.float64),
                vals=np.asarray(vals, dtype=np.float64).reshape(-1), vals2=np.asarray(vals2, dtype=np.float64).reshape(-1))


def _unit(d):
    d = np.asarray(d, dtype=np.float64); return d / math.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])       # Eigen normalize()


def imp_harmonic(freq, amp, dir=(1, 0, 0), t0=0., perpFree=False):
    """HarmonicOscillation (Impose.hpp:5-25)"""
    return _imp(IMP_HARMONIC, 1 if perpFree else 0, [freq, amp, *_unit(dir), t0])


def imp_aligned_harmonic(freqs, amps):
    """AlignedHarmonicOscillations (Impose.hpp:49-65); NaN frequency: axis untouched"""
    return _imp(IMP_ALIGNED_HARMONIC, 0, [*freqs, *amps])


def imp_constant_force(F):
    return _imp(IMP_CONSTANT_FORCE, 0, F)


def imp_radial_force(A, B, F):
    return _imp(IMP_RADIAL_FORCE, 0, [*A, *B, F])


def imp_local6(values, whats, ori=(0, 0, 0, 1)):
    """Local6Dofs (Impose.hpp:90-110): whats 0 nothing, 1 velocity, 2 force per local DoF"""
    return _imp(IMP_LOCAL6, 0, [*values, *whats], ori)


def imp_velocity_and_read_force(dir, vel, latBlock=True, invF=False):
    return _imp(IMP_VEL_READ_FORCE, (1 if latBlock else 0) | (2 if invF else 0), [*_unit(dir), vel])


def imp_read_force(node_pos=None, node_ori=(0, 0, 0, 1)):
    return _imp(IMP_READ_FORCE, 0 if node_pos is None else 1, (0, 0, 0) if node_pos is None else node_pos, node_ori)


def imp_variable_velocity3d(times, vels, angVels=(), ori=(0, 0, 0, 1), diff=False, wrap=False, t0=0.):
    return _imp(IMP_VARIABLE_VEL3D, (1 if diff else 0) | (2 if wrap else 0), [t0], ori, times, vels, angVels)


def imp_variable_aligned_rotation(axis, timeAngVel, wrap=False):
    ta = np.asarray(timeAngVel, dtype=np.float64).reshape(-1, 2)
    return _imp(IMP_VARIABLE_ALIGNED_ROTATION, 1 if wrap else 0, [axis], (0, 0, 0, 1), ta[:, 0], ta[:, 1])


def imp_circular_orbit(node_pos, omega, node_ori=(0, 0, 0, 1), rotate=False):
    """CircularOrbit (Impose.hpp, Impose.cpp:22-37): orbiting around the local z axis of the node (pos, ori)"""
    return _imp(IMP_CIRCULAR_ORBIT, 1 if rotate else 0, [*node_pos, omega], node_ori)


def imp_stable_circular_orbit(node_pos, omega, radius, node_ori=(0, 0, 0, 1), rotate=False):
    """StableCircularOrbit (Impose.cpp:38-48): as CircularOrbit, with the radius enforced (what woo.triangulated.cylinder(angVel=..., realRot=True) puts on every node)"""
    return _imp(IMP_STABLE_CIRCULAR_ORBIT, 1 if rotate else 0, [*node_pos, omega, radius], node_ori)


def imp_interpolated_motion(times, poss, oris, t0=0.):
    """InterpolatedMotion (Impose.hpp, Impose.cpp:111-133): reach poss[k], oris[k] (x,y,z,w) at times[k]+t0; on master nodes of facet groups"""
    return _imp(IMP_INTERPOLATED_MOTION, 0, [t0], (0, 0, 0, 1), times, poss, oris)


def attach_impose(S, ops, spheres=(), walls=(), groups=(), facets=(), clumps=()):
    """Put the impositions `ops` (descriptor dicts; several = CombinedImpose) on the given sphere / wall / facet-group indices."""
    ix = []
    for o in ops:
        S.imposes.append(o); ix.append(len(S.imposes) - 1)
    S.impose_sets.append(ix); k = len(S.impose_sets) - 1
    if len(spheres):
        if S.sphere_impose is None: S.sphere_impose = np.full(len(S.pos), -1, np.int32)
        S.sphere_impose[np.asarray(spheres, dtype=np.int64)] = k
    if len(walls):
        if S.wall_impose is None: S.wall_impose = np.full(len(S.walls), -1, np.int32)
        S.wall_impose[np.asarray(walls, dtype=np.int64)] = k
    if len(groups):
        if S.group_impose is None: S.group_impose = np.full(len(S.facet_groups), -1, np.int32)
        S.group_impose[np.asarray(groups, dtype=np.int64)] = k
    if len(facets):
        if S.facet_impose is None: S.facet_impose = np.full(len(S.facets), -1, np.int32)
        S.facet_impose[np.asarray(facets, dtype=np.int64)] = k
    if len(clumps):
        if S.clump_impose is None: S.clump_impose = np.full(len(S.clumps), -1, np.int32)
        S.clump_impose[np.asarray(clumps, dtype=np.int64)] = k
    return k


CLUMP_CLUMPED = 64      # DemData flag of a member node (src/dem/Particle.hpp:200)


def make_clumps(S, groups, vel=None, angVel=None):
    """Turn groups of (non-overlapping) spheres of a finalized SceneSpec into rigid clumps: mass, centroid, principal inertia and orientation as
    ClumpData::makeClump derives them (Clump.cpp:187-261; numpy's eigh stands in for Woo's computePrincipalAxes — the numbers are inputs to both the oracle and the engine).
    Members get the CLUMPED flag, the clump's velocity field and are listed in S.clumps."""
    m, I = S.mass_inertia()
    for gi, g in enumerate(groups):
        g = np.asarray(g, dtype=np.int64)
        M = m[g].sum(); c = (m[g, None] * S.pos[g]).sum(axis=0) / M
        T = np.zeros((3, 3))
        for k in g:
            d = S.pos[k] - c
            T += I[k] * np.eye(3) + m[k] * (d @ d * np.eye(3) - np.outer(d, d))
        w, V = np.linalg.eigh(T)
        if np.linalg.det(V) < 0:
            V[:, 2] = -V[:, 2]
        # Eigen Quaternion(Matrix3) of the rotation whose columns are the principal axes
        tr = V[0, 0] + V[1, 1] + V[2, 2]
        if tr > 0:
            s4 = math.sqrt(tr + 1.) * 2; q = np.array([(V[2, 1] - V[1, 2]) / s4, (V[0, 2] - V[2, 0]) / s4, (V[1, 0] - V[0, 1]) / s4, .25 * s4])
        else:
            i = int(np.argmax([V[0, 0], V[1, 1], V[2, 2]])); j, k2 = (i + 1) % 3, (i + 2) % 3
            s4 = math.sqrt(V[i, i] - V[j, j] - V[k2, k2] + 1.) * 2
            q = np.zeros(4); q[i] = .25 * s4; q[j] = (V[j, i] + V[i, j]) / s4; q[k2] = (V[k2, i] + V[i, k2]) / s4; q[3] = (V[k2, j] - V[j, k2]) / s4
        q /= np.linalg.norm(q)
        cv = np.zeros(3) if vel is None else np.asarray(vel[gi], dtype=np.float64); cw = np.zeros(3) if angVel is None else np.asarray(angVel[gi], dtype=np.float64)
        S.clumps.append(dict(members=[int(k) for k in g], pos=c, ori=q, vel=cv, angVel=cw, mass=float(M), inertia=w.copy(), flags=0))
        S.flags[g] |= CLUMP_CLUMPED
        S.vel[g] = cv + np.cross(cw, S.pos[g] - c); S.angVel[g] = cw
    return S


def box_walls(lo, hi, mat=0, top=False):
    """5 (or 6) axis-aligned Walls with sense pointing inwards."""
    w = []
    for ax in range(3):
        p = [0., 0., 0.]; p[ax] = lo[ax]
        w.append(dict(pos=tuple(p), axis=ax, sense=1, mat=mat, mask=MASK_BOUNDARY))
        if ax < 2 or top:
            p = [0., 0., 0.]; p[ax] = hi[ax]
            w.append(dict(pos=tuple(p), axis=ax, sense=-1, mat=mat, mask=MASK_BOUNDARY))
    return w


def lattice(nx, ny, nz, spacing, origin, jitter, rng):
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    p = np.asarray(origin) + (g + .5) * spacing
    if jitter > 0:
        p = p + rng.uniform(-jitter, jitter, p.shape)
    return p


def deposition(n, *, r=5e-3, psd_ratio=1.0, seed=1, law=LAW_IDEALELPL, layers=None, damping=.2, young=None,
               density=2500., shuffle=True, spacing_factor=2.2, en=float("nan")):
    """Gravity deposition in an open-top box of 5 Walls (cfg1/cfg2/cfg3 of BASELINE.json).

    Spheres sit on a jittered cubic lattice (spacing = spacing_factor * r_max), radii log-uniform on
    [r, psd_ratio*r].  ``layers`` fixes the lattice height (default: cube)."""
    rng = np.random.default_rng(seed)
    rmax = r * psd_ratio
    sp = spacing_factor * rmax
    if layers is None:
        nz = max(1, round(n ** (1 / 3.)))
    else:
        nz = layers
    nx = max(1, int(math.ceil(math.sqrt(n / nz))))
    ny = max(1, int(math.ceil(n / (nx * nz))))
    p = lattice(nx, ny, nz, sp, (0, 0, 0), 0.05 * r, rng)
    if shuffle:
        p = p[rng.permutation(len(p))]
    p = p[:n]
    rad = np.full(n, r) if psd_ratio == 1.0 else r * np.exp(rng.uniform(0, math.log(psd_ratio), n))
    S = SceneSpec()
    if law == LAW_IDEALELPL:
        S.mats = [frict_mat(density=density, young=young or 1e7, ktDivKn=.2, tanPhi=.5)]
    else:
        S.mats = [hertz_mat(density=density, young=young or 1e8, tanPhi=.5)]
    S.pos, S.radius = p, rad
    S.walls = box_walls((0, 0, 0), (nx * sp, ny * sp, nz * sp), mat=0)
    S.params = Params(gravity=(0, 0, -9.81), damping=damping, law=law, verletDist=-.05, dtSafety=.9, dynDtPeriod=100, en=en)
    return S.finalize()


def clump_deposition(n_clumps, n_free=0, *, r=5e-3, seed=11, law=LAW_IDEALELPL, damping=.2, dynDtPeriod=100, spin=5.0):
    """Gravity deposition of rigid aggregates (dumbbells, L-shaped trimers of unequal spheres, 4-sphere chains; random orientation and spin) mixed with free
    spheres in an open-top box of 5 Walls.  The members of a clump touch or overlap slightly (they never collide with each other)."""
    rng = np.random.default_rng(seed)
    n_sites = n_clumps + n_free
    nz = max(1, round(n_sites ** (1 / 3.))); nx = max(1, int(math.ceil(math.sqrt(n_sites / nz)))); ny = max(1, int(math.ceil(n_sites / (nx * nz))))
    sp = 5.2 * r       # every aggregate fits a ball of radius 2.5 r
    sites = lattice(nx, ny, nz, sp, (0, 0, 0), 0.05 * r, rng)
    sites = sites[rng.permutation(len(sites))][:n_sites]
    pos, rad, groups = [], [], []
    for i in range(n_clumps):
        kind = i % 3
        if kind == 0: loc, rr = np.array([[-.9, 0, 0], [.9, 0, 0]]) * r, [r, r]
        elif kind == 1: loc, rr = np.array([[0, 0, 0], [1.7, 0, 0], [0, 1.4, .3]]) * r, [r, .8 * r, .6 * r]
        else: loc, rr = np.array([[-1.5, 0, 0], [-.5, 0, .2], [.5, 0, -.2], [1.5, 0, 0]]) * r, [.7 * r, .8 * r, .8 * r, .7 * r]
        loc = loc - loc.mean(axis=0)
        q = rng.normal(size=4); q /= np.linalg.norm(q); x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        groups.append(list(range(len(pos), len(pos) + len(rr))))
        pos += list(sites[i] + loc @ R.T); rad += rr
    for i in range(n_free):
        pos.append(sites[n_clumps + i]); rad.append(r * rng.uniform(.7, 1.))
    S = SceneSpec()
    S.mats = [frict_mat(density=2500., young=1e7, ktDivKn=.2, tanPhi=.5)] if law == LAW_IDEALELPL else [hertz_mat(density=2500., young=1e8, tanPhi=.5)]
    S.pos, S.radius = np.array(pos), np.array(rad)
    S.walls = box_walls((0, 0, 0), (nx * sp, ny * sp, nz * sp), mat=0)
    S.params = Params(gravity=(0, 0, -9.81), damping=damping, law=law, verletDist=-.05, dtSafety=.9, dynDtPeriod=dynDtPeriod)
    S.finalize()
    S.vel[len(S.pos) - n_free:] = rng.uniform(-.05, .05, (n_free, 3))
    return make_clumps(S, groups, vel=rng.uniform(-.05, .05, (n_clumps, 3)), angVel=rng.uniform(-spin, spin, (n_clumps, 3)))


def random_gas(n, *, box=1.0, r=0.02, psd_ratio=1.0, seed=0, vmax=1.0, law=LAW_IDEALELPL, overlap=True, walls=True,
               angvel=1.0, en=float("nan")):
    """Dense random cloud with overlaps and random velocities: exercises creation, update, breakage and sliding
    of many contacts in few steps (parity stress test, not a physical config)."""
    rng = np.random.default_rng(seed)
    S = SceneSpec()
    S.mats = [frict_mat(density=2500., young=1e7, tanPhi=.5)] if law == LAW_IDEALELPL else [hertz_mat(density=2500., young=1e8)]
    S.pos = rng.uniform(0.05 * box, 0.95 * box, (n, 3))
    S.radius = np.full(n, r) if psd_ratio == 1.0 else r * np.exp(rng.uniform(0, math.log(psd_ratio), n))
    S.vel = rng.uniform(-vmax, vmax, (n, 3))
    S.angVel = rng.uniform(-angvel, angvel, (n, 3))
    if walls:
        S.walls = box_walls((0, 0, 0), (box, box, box), mat=0, top=True)
    S.params = Params(gravity=(0, 0, -9.81), damping=.2, law=law, verletDist=-.05, dynDtPeriod=10, en=en)
    return S.finalize()


def quad_facets(p0, p1, p2, p3, nu, nv, **kw):
    """Triangulate the parallelogram p0,p1,p2,p3 (p0->p1 = u, p0->p3 = v) into 2*nu*nv facets."""
    p0, p1, p3 = np.asarray(p0, float), np.asarray(p1, float), np.asarray(p3, float)
    u, v = (p1 - p0) / nu, (p3 - p0) / nv
    out = []
    for i in range(nu):
        for j in range(nv):
            a = p0 + i * u + j * v; b = a + u; c = a + u + v; d = a + v
            out.append(dict(v=np.array([a, b, c]), **kw)); out.append(dict(v=np.array([a, c, d]), **kw))
    return out


def facet_box_scene(n=300, *, box=.4, r=.02, psd_ratio=2.0, seed=0, law=LAW_IDEALELPL, div=3, halfThick=0.0, fakeVel=(0, 0, 0),
                    random_n21lim=True, vmax=1.0, en=float("nan")):
    """Spheres thrown into an open box made of triangulated facets (floor + 4 sides): exercises face/edge/vertex regions of
    Cg2_Facet_Sphere_L6Geom, halfThick, fakeVel and the n21lim neighbour-angle adjustment (values are synthetic)."""
    rng = np.random.default_rng(seed)
    S = SceneSpec()
    S.mats = [frict_mat(density=2500., young=1e7, tanPhi=.5)] if law == LAW_IDEALELPL else [hertz_mat(density=2500., young=1e8)]
    b = box
    kw = dict(halfThick=halfThick, fakeVel=fakeVel, mat=0, mask=MASK_BOUNDARY)
    F = quad_facets((0, 0, 0), (b, 0, 0), (b, b, 0), (0, b, 0), div, div, **kw)
    F += quad_facets((0, 0, 0), (0, 0, b), (0, b, b), (0, b, 0), div, div, **kw)
    F += quad_facets((b, 0, 0), (b, b, 0), (b, b, b), (b, 0, b), div, div, **kw)
    F += quad_facets((0, 0, 0), (b, 0, 0), (b, 0, b), (0, 0, b), div, div, **kw)
    F += quad_facets((0, b, 0), (0, b, b), (b, b, b), (b, b, 0), div, div, **kw)
    for f in F:
        if random_n21lim:
            nl = rng.uniform(-1, 1, 3); nl[rng.uniform(size=3) < .3] = np.nan
            f["n21lim"] = nl
        if rng.uniform() < .2:
            f["vvel"] = rng.uniform(-.2, .2, 9)
    S.facets = F
    S.pos = rng.uniform(-.2 * r, b + .2 * r, (n, 3)); S.pos[:, 2] = rng.uniform(-.2 * r, .6 * b, n)
    S.radius = np.full(n, r) if psd_ratio == 1.0 else r * np.exp(rng.uniform(0, math.log(psd_ratio), n))
    S.vel = rng.uniform(-vmax, vmax, (n, 3)); S.angVel = rng.uniform(-1, 1, (n, 3))
    S.params = Params(gravity=(0, 0, -9.81), damping=.2, law=law, verletDist=-.05, dynDtPeriod=10, en=en)
    return S.finalize()


def periodic_gas(n, *, cell=(.5, .5, .5), r=.02, psd_ratio=1.0, seed=0, vmax=1.0, law=LAW_IDEALELPL, gradV=None, homoDeform=HOMO_GRADV2,
                 shear0=0.0, scatter=2, walls=False, en=float("nan"), mean_field=True):
    """Random overlapping cloud in a periodic cell (cfg4 of BASELINE.json in small): positions are NOT wrapped — a fraction is shifted by
    whole cell vectors (``scatter`` periods each way) so that non-zero Contact.cellDist occurs from the first collider run.
    ``shear0`` pre-skews the cell (hSize[0,1]); ``gradV`` (3x3) drives it; initial velocities = fluctuation + gradV·x (mean field)."""
    rng = np.random.default_rng(seed)
    S = SceneSpec()
    S.mats = [frict_mat(density=2500., young=1e7, tanPhi=.5)] if law == LAW_IDEALELPL else [hertz_mat(density=2500., young=1e8)]
    H = np.diag(np.asarray(cell, dtype=np.float64)); H[0, 1] = shear0 * cell[1]
    g = np.zeros((3, 3)) if gradV is None else np.asarray(gradV, dtype=np.float64).reshape(3, 3)
    frac = rng.uniform(0, 1, (n, 3))
    if walls:   # walls at z=0.1 and z=0.9 of the cell height (non-skewed cells only)
        frac[:, 2] = rng.uniform(.12, .88, n)
    shift = rng.integers(-scatter, scatter + 1, (n, 3)) if scatter > 0 else np.zeros((n, 3))
    if walls:
        shift[:, 2] = 0
    S.pos = (frac + shift) @ H.T
    S.radius = np.full(n, r) if psd_ratio == 1.0 else r * np.exp(rng.uniform(0, math.log(psd_ratio), n))
    S.vel = rng.uniform(-vmax, vmax, (n, 3))
    if mean_field:
        S.vel = S.vel + S.pos @ g.T
    S.angVel = rng.uniform(-1, 1, (n, 3))
    if walls:
        S.walls = [dict(pos=(0, 0, .1 * cell[2]), axis=2, sense=1, mat=0, mask=MASK_BOUNDARY), dict(pos=(0, 0, .9 * cell[2]), axis=2, sense=-1, mat=0, mask=MASK_BOUNDARY)]
    S.params = Params(gravity=(0, 0, 0), damping=.1, law=law, verletDist=-.05, dynDtPeriod=10, en=en, periodic=1)
    S.cell = dict(hSize=H, gradV=g, nextGradV=g, homoDeform=homoDeform)
    return S.finalize()


def hertz_deposition(n, *, r=1e-3, seed=3, aspect=(8, 1, 1), en=.5, young=1e8, density=2500., spacing_factor=2.2, vz0=0.0, vfluct=0.0):
    """cfg3 of BASELINE.json in small: monodisperse Hertz/DMT deposition (HertzMat, poisson .2, en .5, noAttraction, damping 0) in a box of
    5 walls with aspect 8:1:1, the long axis x being the slab axis of the multi-GPU runs."""
    rng = np.random.default_rng(seed)
    sp = spacing_factor * r
    k = (n / (aspect[0] * aspect[1] * aspect[2])) ** (1 / 3.)
    nx, ny, nz = (max(1, int(round(k * a))) for a in aspect)
    while nx * ny * nz < n:
        nx += 1
    p = lattice(nx, ny, nz, sp, (0, 0, 0), 0.05 * r, rng)
    p = p[rng.permutation(len(p))][:n]
    S = SceneSpec()
    S.mats = [hertz_mat(density=density, young=young, tanPhi=.5, surfEnergy=0., alpha=0.)]
    S.pos, S.radius = p, np.full(n, r)
    if vz0 or vfluct:
        S.vel = np.tile([0., 0., -vz0], (n, 1)) * rng.uniform(.5, 1.5, (n, 1)) + rng.uniform(-vfluct, vfluct, (n, 3))
    S.walls = box_walls((0, 0, 0), (nx * sp, ny * sp, nz * sp), mat=0)
    S.params = Params(gravity=(0, 0, -9.81), damping=0., law=LAW_HERTZ_DMT, verletDist=-.05, dtSafety=.9, dynDtPeriod=100, en=en, poisson=.2, noAttraction=1)
    return S.finalize()


def periodic_shear(n, *, r=1e-3, seed=4, solid=.55, rate=1.0, vfluct=0.0, young=1e7, density=2500.):
    """cfg4 of BASELINE.json in small: monodisperse spheres at solid fraction ``solid`` in a cubic periodic cell under simple shear
    gradV = [[0,rate,0],[0,0,0],[0,0,0]] (homoDeform gradV2), linear law, no gravity; velocities start on the mean field gradV·x."""
    rng = np.random.default_rng(seed)
    L = (n * (4 / 3.) * math.pi * r ** 3 / solid) ** (1 / 3.)
    m = int(math.ceil(n ** (1 / 3.)))
    sp = L / m
    g = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    p = (g + .5) * sp + rng.uniform(-.05 * r, .05 * r, g.shape)
    p = p[rng.permutation(len(p))][:n]
    G = np.array([[0, rate, 0], [0, 0, 0], [0, 0, 0]], dtype=np.float64)
    S = SceneSpec()
    S.mats = [frict_mat(density=density, young=young, ktDivKn=.2, tanPhi=.5)]
    S.pos, S.radius = p, np.full(n, r)
    S.vel = p @ G.T + rng.uniform(-vfluct, vfluct, p.shape)
    S.params = Params(gravity=(0, 0, 0), damping=.2, law=LAW_IDEALELPL, verletDist=-.05, dtSafety=.9, dynDtPeriod=100, periodic=1)
    S.cell = dict(hSize=np.eye(3) * L, gradV=G, nextGradV=G, homoDeform=HOMO_GRADV2)
    return S.finalize()


def drum(n, *, r=2e-3, seed=5, R=None, length=None, div=32, axial=8, omega=5.0, young=1e7, density=2500., fill=.35, vz0=0.0, vfluct=0.0, rotate=False):
    """cfg5 of BASELINE.json in small: spheres in a horizontal cylinder (axis x) made of static facets whose ``fakeVel = -R*omega*n x ax`` mimics the
    rotation (py/triangulated.py:56-58), closed by two Walls; linear law.  Spheres start on a jittered lattice in the lower part."""
    rng = np.random.default_rng(seed)
    sp = 2.2 * r
    vol = n * sp ** 3 / fill
    if R is None:
        R = (vol / (math.pi * 4)) ** (1 / 3.)          # length = 4 R
    if length is None:
        length = 4 * R
    ax = np.array([1., 0, 0])
    F = []
    th = np.linspace(0, 2 * math.pi, div + 1)
    xs = np.linspace(0, length, axial + 1)
    for i in range(axial):
        for j in range(div):
            a = np.array([xs[i], R * math.cos(th[j]), R * math.sin(th[j])]); b = np.array([xs[i + 1], R * math.cos(th[j]), R * math.sin(th[j])])
            c = np.array([xs[i + 1], R * math.cos(th[j + 1]), R * math.sin(th[j + 1])]); d = np.array([xs[i], R * math.cos(th[j + 1]), R * math.sin(th[j + 1])])
            for tri in ((a, b, c), (a, c, d)):
                nn = np.cross(tri[1] - tri[0], tri[2] - tri[0]); nn /= np.linalg.norm(nn)
                F.append(dict(v=np.array(tri), halfThick=0., fakeVel=tuple(-R * omega * np.cross(nn, ax)), mat=0, mask=MASK_BOUNDARY))
    # lattice points inside the cylinder, lowest first
    m = int(math.ceil(2 * R / sp)); mx = int(length / sp) - 1
    g = np.stack(np.meshgrid(np.arange(mx), np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    p = np.array([sp, -R, -R]) + (g + .5) * sp
    p = p[np.hypot(p[:, 1], p[:, 2]) < R - 1.5 * r]
    p = p[np.argsort(p[:, 2], kind="stable")][:n]
    if len(p) < n:
        raise ValueError("drum too small for %d spheres" % n)
    p = p + rng.uniform(-.05 * r, .05 * r, p.shape)
    S = SceneSpec()
    S.mats = [frict_mat(density=density, young=young, ktDivKn=.2, tanPhi=.5)]
    S.pos, S.radius = p[rng.permutation(n)], np.full(n, r)
    if vz0 or vfluct:
        S.vel = np.tile([0., 0., -vz0], (n, 1)) * rng.uniform(.5, 1.5, (n, 1)) + rng.uniform(-vfluct, vfluct, (n, 3))
    S.facets = F
    if rotate:   # the drum really turns: the facets are clumped to a master node on the axis that carries the angular velocity (examples/mill.py:38-42)
        for f in F:
            f["fakeVel"] = (0., 0., 0.)
        S.facet_groups = [dict(pos=(.5 * length, 0., 0.), ori=(0., 0., 0., 1.), vel=(0., 0., 0.), angVel=(omega, 0., 0.), facets=list(range(len(F))))]
    S.walls = [dict(pos=(0, 0, 0), axis=0, sense=1, mat=0, mask=MASK_BOUNDARY), dict(pos=(length, 0, 0), axis=0, sense=-1, mat=0, mask=MASK_BOUNDARY)]
    S.params = Params(gravity=(0, 0, -9.81), damping=.2, law=LAW_IDEALELPL, verletDist=-.05, dtSafety=.9, dynDtPeriod=100)
    return S.finalize()
