"""Settle the bench workload (untimed), then run a few steps inside a cudaProfilerStart/Stop range (use with ncu --profile-from-start off).
  python scripts/profile_dense.py [n=1000000] [settle=3000] [steps=3] [law=0|1]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv, argv = sys.argv[:1], sys.argv[1:]
import bench
from woo_b200 import scenes as sc
from woo_b200.engine import Engine
n = int(argv[0]) if len(argv) > 0 else 1000000
settle = int(argv[1]) if len(argv) > 1 else 3000
steps = int(argv[2]) if len(argv) > 2 else 3
S, tiling = bench.workload_tiled(n)
E = Engine(S.params).load(S)
E.set_profiling(True)
done = 0
while done < settle:
    k = min(1000, settle - done); st = E.step(k); done += k
    print(done, "rebuilds", st.n_rebuilds, "ms/step %.4f" % (st.device_ms / k), "phases", [round(x / k, 4) for x in (st.ms_integrate, st.ms_rebuild, st.ms_contact, st.ms_reduce, st.ms_dyndt)],
          "real", st.n_real, "cand", st.n_contacts, flush=True)
E.set_profiling(False)
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
st = E.step(steps)
rt.cudaProfilerStop()
print("profiled", steps, "steps; rebuilds", st.n_rebuilds, "launches", st.n_launches, "real", st.n_real, "cand", st.n_contacts)
