"""cfg1 (10k spheres) stepping-loop diagnostics on the GPU box: WDEM_STEP_TRACE=1 python scripts/small_scene_probe.py [n=10000] [steps=20000]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from woo_b200 import scenes as sc
from woo_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
S = sc.deposition(n, seed=1)
E = Engine(S.params).load(S)
E.step(50)
E = Engine(S.params).load(S)
done = 0
for k in (1000, 4000, 5000, steps - 10000):
    t0 = time.time(); st = E.step(k); w = time.time() - t0; done += k
    print("steps %6d..%6d: device %.1f us/step, wall %.1f us/step, rebuilds %d, launches/step %.1f, real %d" % (done - k, done, st.device_ms / k * 1e3, w / k * 1e6, st.n_rebuilds, st.n_launches / k, st.n_real), flush=True)
