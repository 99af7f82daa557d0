"""Device-time data points for the BASELINE configs that bench.py does not time (it times configs[1]): cfg4 (periodic simple shear) and cfg5 (drum of facets,
static mesh with fakeVel and, as an extra, the same drum really rotating as a kinematic clump).  One JSON line per case on stdout.
  python scripts/bench_configs.py [n=1000000] [settle=1500] [steps=200]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from woo_b200 import scenes as sc
from woo_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
cases = {
    "cfg4_periodic_shear": lambda: sc.periodic_shear(n, solid=.6, rate=1., vfluct=.01),       # BASELINE cfg4: shear rate 1/s
    "cfg5_drum_fakevel": lambda: sc.drum(n, div=256, axial=64, vz0=.5, vfluct=.2),
    "cfg5_drum_rotating_clump": lambda: sc.drum(n, div=256, axial=64, vz0=.5, vfluct=.2, rotate=True),
}
only = sys.argv[4].split(",") if len(sys.argv) > 4 else None
for name, make in cases.items():
    if only and name not in only:
        continue
    t0 = time.time(); S = make(); E = Engine(S.params).load(S)
    E.step(settle); E.set_profiling(True)
    st = E.step(steps)
    print(json.dumps({"case": name, "n_spheres": S.n, "n_facets": len(S.facets), "particle_steps_per_s": S.n * steps / (st.device_ms * 1e-3), "ms_per_step": st.device_ms / steps,
                      "rebuilds": st.n_rebuilds, "real_contacts": st.n_real, "candidates": st.n_contacts,
                      "phase_ms_per_step": {k: getattr(st, k) / steps for k in ("ms_integrate", "ms_rebuild", "ms_contact", "ms_reduce", "ms_dyndt")}, "wall_s_total": round(time.time() - t0, 1)}), flush=True)
    E.close()
