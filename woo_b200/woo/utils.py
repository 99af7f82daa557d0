"""woo.utils of the reference (py/utils.py), the functions that set up the hot path."""
import math
from .dem import (FrictMat, Sphere, Wall, Facet, DemField, Leapfrog, InsertionSortCollider, Bo1_Sphere_Aabb, Bo1_Facet_Aabb, Bo1_Wall_Aabb, ContactLoop,
                  Cg2_Sphere_Sphere_L6Geom, Cg2_Facet_Sphere_L6Geom, Cg2_Wall_Sphere_L6Geom, Cp2_FrictMat_FrictPhys, Law2_L6Geom_FrictPhys_IdealElPl, DynDt,
                  GpuDemLoop)


def spherePWaveDt(radius, density, young):
    """py/utils.py:77-85"""
    return radius / math.sqrt(young / density)


def defaultMaterial():
    """py/utils.py:87-92"""
    return FrictMat(density=1e3, young=1e7, ktDivKn=.2, tanPhi=math.tan(.5))


def defaultEngines(damping=0., gravity=None, verletDist=-.05, noSlip=False, noBreak=False, cp2=None, law=None, dynDtPeriod=100):
    """py/utils.py:94-129 — the reference's engine list, as attribute-carrying objects (they cannot run on their own here)."""
    if gravity:
        raise ValueError("gravity MUST NOT be specified anymore, set DemField.gravity=... instead.")
    cp2 = cp2 if cp2 else Cp2_FrictMat_FrictPhys()
    law = law if law else Law2_L6Geom_FrictPhys_IdealElPl(noSlip=noSlip, noBreak=noBreak, label='contactLaw')
    return [
        Leapfrog(damping=damping, reset=True, label='leapfrog'),
        InsertionSortCollider([Bo1_Sphere_Aabb(), Bo1_Facet_Aabb(), Bo1_Wall_Aabb()], label='collider', verletDist=verletDist),
        ContactLoop([Cg2_Sphere_Sphere_L6Geom(), Cg2_Facet_Sphere_L6Geom(), Cg2_Wall_Sphere_L6Geom()], [cp2], [law], applyForces=True, label='contactLoop'),
    ] + ([DynDt(stepPeriod=dynDtPeriod, label='dynDt')] if dynDtPeriod > 0 else [])


def gpuEngines(device=0, **kw):
    """defaultEngines(**kw) folded into the one engine that runs them on the GPU."""
    e = defaultEngines(**kw)
    return [GpuDemLoop(e[0], e[1], e[2], e[3] if len(e) > 3 else None, device=device, label='gpuDemLoop')]


def sphere(center, radius, mat=defaultMaterial, fixed=False, wire=False, color=None, highlight=False, mask=DemField.defaultMovableMask, vel=None):
    """py/utils.py:153-180"""
    p = Sphere.make(center, radius, mat=mat() if callable(mat) else mat, mask=mask, fixed=fixed)
    if vel is not None:
        p.vel = vel
    return p


def wall(position, axis, sense=0, fixed=True, mat=defaultMaterial, mask=DemField.defaultBoundaryMask, **kw):
    """py/utils.py:245-267"""
    return Wall.make(position, axis=axis, sense=sense, mat=mat() if callable(mat) else mat, mask=mask, fixed=fixed)


def wallBox(box, which=(1, 1, 1, 1, 1, 1), **kw):
    """py/utils.py:269-290: walls at box=((x0,y0,z0),(x1,y1,z1)), senses pointing inside."""
    lo, hi = box
    ret = []
    for ax in range(3):
        if which[ax]:
            ret.append(wall(lo[ax], axis=ax, sense=1, **kw))
        if which[ax + 3]:
            ret.append(wall(hi[ax], axis=ax, sense=-1, **kw))
    return ret


def facet(vertices, fakeVel=None, halfThick=0., fixed=True, mat=defaultMaterial, mask=DemField.defaultBoundaryMask, **kw):
    """py/utils.py:318-345"""
    return Facet.make(vertices, mat=mat() if callable(mat) else mat, halfThick=halfThick, fakeVel=fakeVel, mask=mask, fixed=fixed)
