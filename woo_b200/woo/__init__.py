"""Reference-shaped namespace over the B200 engine: ``from woo_b200 import woo; import woo_b200.woo.core, woo_b200.woo.dem, woo_b200.woo.utils``.

Scripts written for woodem/woo's DEM hot path (spheres, walls, static facets; linear and Hertz/DMT laws) port by changing the import and
wrapping the engine group:  ``S.engines = woo.utils.gpuEngines(damping=.2)``  (= ``[GpuDemLoop(Leapfrog, InsertionSortCollider, ContactLoop, DynDt)]``)
instead of ``woo.utils.defaultEngines(damping=.2)`` (reference: py/utils.py:94-129).  The classes are the compiled ones of
woo_b200/csrc/woo_plugin.cpp; there is no CPU fallback behind them.
"""
from .. import plugin_build as _pb

_pb.build()
from .. import _woo  # noqa: E402
from . import core, dem, utils  # noqa: E402,F401
