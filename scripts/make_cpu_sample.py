"""Generates tests/golden/cfg2_sample_20k.npz: a settled 20k-sphere PSD 1:4 packing (same generator, material and law as
bench.py's 1M workload, 12 lattice layers) used as the bounded sample of the CPU baseline / reference arm.
Run on a GPU box:  python scripts/make_cpu_sample.py gpurun_out/cfg2_sample_20k.npz   (then copy into tests/golden/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from woo_b200 import scenes as sc
from woo_b200.engine import Engine

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cfg2_sample_20k.npz"
n = 20000
S = sc.deposition(n, r=1e-3, psd_ratio=4.0, seed=2, law=sc.LAW_IDEALELPL, layers=12, damping=.2, young=1e7, density=2500.)
E = Engine(S.params).load(S)
for i in range(15):
    st = E.step(2000)
    print(i, "Ek", E.kinetic_energy(), "real", st.n_real, "cand", st.n_contacts, "ms/step", st.device_ms / 2000, flush=True)
s = E.spheres()
lo = np.array([0., 0., 0.]); hi = np.array([w["pos"][w["axis"]] for w in S.walls if w["sense"] == -1] + [1.0])
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
np.savez_compressed(out, pos=s["pos"], vel=s["vel"], angVel=s["angVel"], radius=S.radius, lo=lo, hi=hi, dt=E.scalars()["dt"])
print("saved", out, "coordination", 2 * st.n_real / n)
