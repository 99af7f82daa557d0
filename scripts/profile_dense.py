"""Settle a packing (untimed), then run a few steps inside a cudaProfilerStart/Stop range (use with ncu --profile-from-start off)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from woo_b200 import scenes as sc
from woo_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
S = sc.deposition(n, r=1e-3, psd_ratio=4.0, seed=2, law=sc.LAW_IDEALELPL, layers=40, damping=.2, young=1e7, density=2500.)
E = Engine(S.params).load(S)
done = 0
while done < settle:
    st = E.step(min(2000, settle - done)); done += 2000
    print(done, "rebuilds", st.n_rebuilds, "ms/step", st.device_ms / 2000, "real", st.n_real, "cand", st.n_contacts, "Ek", E.kinetic_energy(), flush=True)
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
st = E.step(steps)
rt.cudaProfilerStop()
print("profiled", steps, "steps; rebuilds", st.n_rebuilds, "launches", st.n_launches, "real", st.n_real, "cand", st.n_contacts)
