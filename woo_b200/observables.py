"""Long-run observables of SURVEY.md §8c (the 1 % criterion): kinetic energy, coordination number, packing fraction — computed on the host from
downloaded state; the definitions are the reference's."""
import math
import numpy as np


def coord_number(sphere_ids, pos, real_contact_ids, box=None, skip_free=True):
    """DemFuncs::coordNumber (src/dem/Funcs.cpp:84-102) over the spheres whose centre is inside ``box`` ((lo,hi), None = all):
    (C2-N1)/(N-N0-N1) with C2 = contact count summed per particle (contacts with walls/facets count for the sphere), N0 floaters, N1 rattlers."""
    sphere_ids = np.asarray(sphere_ids)
    cnt = {}
    for a, b in np.asarray(real_contact_ids).reshape(-1, 2):
        cnt[a] = cnt.get(a, 0) + 1; cnt[b] = cnt.get(b, 0) + 1
    n = np.array([cnt.get(i, 0) for i in sphere_ids])
    if box is not None:
        lo, hi = np.asarray(box[0]), np.asarray(box[1])
        inside = np.all((pos >= lo) & (pos <= hi), axis=1)
        n = n[inside]
    N, N0, N1, C2 = len(n), int((n == 0).sum()), int((n == 1).sum()), int(n.sum())
    if skip_free:
        return (C2 - N1) / (N - N0 - N1) if N - N0 - N1 > 0 else float("nan")
    return C2 / N if N else float("nan")


def packing_fraction(pos, radius, box):
    """DemFuncs::porosity (src/dem/Funcs.cpp:104-113; despite the name it returns the solid fraction Vs/V) for centres inside ``box``."""
    lo, hi = np.asarray(box[0], dtype=float), np.asarray(box[1], dtype=float)
    inside = np.all((pos >= lo) & (pos <= hi), axis=1)
    r = np.asarray(radius)[inside]
    return float(((4 / 3.) * math.pi * r * r * r).sum() / np.prod(hi - lo))
