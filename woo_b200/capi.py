"""ctypes binding of the C-ABI in include/wdem.h (libwdem.so).  Thin: no compute here, no CPU fallback — if the
library cannot be built/loaded or there is no sm_100 device, calls raise."""
import ctypes as C
import os
import numpy as np
from . import build as _build

_LIBS = {}


class WdemImpose(C.Structure):
    """include/wdem.h: wdem_impose"""
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("v", C.c_double * 16), ("ori", C.c_double * 4), ("nTimes", C.c_int32), ("pad", C.c_int32),
                ("times", C.c_void_p), ("vals", C.c_void_p), ("vals2", C.c_void_p)]


class WdemConfig(C.Structure):
    _fields_ = [("dt", C.c_double), ("dtSafety", C.c_double), ("gravity", C.c_double * 3), ("distFactor", C.c_double),
                ("damping", C.c_double), ("verletDist", C.c_double), ("wallHalfThick", C.c_double), ("relRollStiff", C.c_double),
                ("relTwistStiff", C.c_double), ("rollTanPhi", C.c_double), ("poisson", C.c_double), ("en", C.c_double),
                ("maxRelInc", C.c_double), ("loneMask", C.c_uint32), ("reset", C.c_int32), ("noRatch", C.c_int32),
                ("iniLensTouch", C.c_int32), ("trsfRenorm", C.c_int32), ("applyForces", C.c_int32), ("law", C.c_int32),
                ("noSlip", C.c_int32), ("noBreak", C.c_int32), ("noFrict", C.c_int32), ("noAttraction", C.c_int32),
                ("dynDtPeriod", C.c_int32), ("periodic", C.c_int32), ("device", C.c_int32), ("evalStress", C.c_int32), ("reserved", C.c_int32 * 6)]


class WdemStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_spheres", "n_contacts", "n_real", "n_rebuilds", "n_full_runs_total", "n_launches", "step")] + \
        [(n, C.c_double) for n in ("dt", "time", "device_ms", "ms_integrate", "ms_contact", "ms_reduce", "ms_rebuild", "ms_dyndt")] + \
        [(n, C.c_int64) for n in ("sum_contacts", "sum_real", "sum_potential_evaluated")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


EXPORTS = ["wdem_default_config", "wdem_create", "wdem_destroy", "wdem_last_error", "wdem_set_config", "wdem_set_profiling",
           "wdem_upload_materials", "wdem_upload_spheres", "wdem_add_spheres", "wdem_remove_spheres", "wdem_upload_forces", "wdem_upload_walls", "wdem_upload_facets", "wdem_set_cell",
           "wdem_set_next_gradv", "wdem_get_cell", "wdem_step", "wdem_step_e2e", "wdem_step_e2e_owned", "wdem_set_periodic_state", "wdem_get_periodic_state", "wdem_download_state", "wdem_download_forces",
           "wdem_set_state", "wdem_num_contacts", "wdem_download_contacts", "wdem_upload_contacts", "wdem_set_scene", "wdem_get_stress", "wdem_download_walls", "wdem_track_energy", "wdem_get_energy", "wdem_get_gradv_work", "wdem_upload_facet_groups", "wdem_download_facets", "wdem_upload_clumps", "wdem_download_clumps", "wdem_num_clumps", "wdem_upload_clump_forces", "wdem_upload_imposes", "wdem_upload_facet_imposes", "wdem_upload_clump_imposes", "wdem_get_impose_state", "wdem_download_boundary_velocities", "wdem_crit_dt", "wdem_kinetic_energy", "wdem_get_scalars",
           "wdem_device_info", "wdem_comm_unique_id", "wdem_comm_init_nccl", "wdem_local_group_create", "wdem_local_group_destroy",
           "wdem_comm_init_local", "wdem_set_domain", "wdem_set_rebalance", "wdem_get_domain", "wdem_num_owned", "wdem_download_owned", "wdem_set_owned", "wdem_domain_stats", "wdem_peer_active"]


def lib_path(variant=""):
    return _build.lib_path(variant)


def load(build=True, variant=""):
    """variant "alt": libwdem_alt.so (flipped reduction orders, woo_b200/build.py), used only by the order-sensitivity test."""
    if variant not in _LIBS:
        path = _build.lib_path(variant)
        if build:
            _build.build(variant=variant)
        if not os.path.exists(path):
            raise RuntimeError("%s is missing: build it with `python -m woo_b200.build --all` (no CPU fallback exists)" % os.path.basename(path))
        L = C.CDLL((os.environ.get("WDEM_LIB_VARIANT") if not variant else None) or path)   # WDEM_LIB_VARIANT: kernel-tuning experiments only
        L.wdem_last_error.restype = C.c_char_p
        L.wdem_last_error.argtypes = [C.c_void_p]
        L.wdem_num_contacts.restype = C.c_int64
        L.wdem_download_contacts.restype = C.c_int64
        L.wdem_num_owned.restype = C.c_int64
        L.wdem_num_clumps.restype = C.c_int64
        L.wdem_local_group_create.restype = C.c_void_p
        _LIBS[variant] = L
    return _LIBS[variant]


def ptr(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)
