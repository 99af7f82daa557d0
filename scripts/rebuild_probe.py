import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from woo_b200 import scenes as sc
from woo_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
S = sc.deposition(n, r=1e-3, psd_ratio=4.0, seed=2, law=sc.LAW_IDEALELPL, layers=40, damping=.2, young=1e7, density=2500.)
E = Engine(S.params).load(S)
E.set_profiling(True)
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 10):
    st = E.step(100)
    sc_ = E.scalars()
    print(i, "rebuilds", st.n_rebuilds, "dt", sc_["dt"], "vd", sc_["verletDist"], "ms/step", st.device_ms / 100, "phases", [round(x / 100, 4) for x in (st.ms_integrate, st.ms_rebuild, st.ms_contact, st.ms_reduce, st.ms_dyndt)], "cand", st.n_contacts, "real", st.n_real, "launches", st.n_launches, flush=True)
