T=${1:-r02s42}
WDEM_STEP_TRACE=1 python scripts/small_scene_probe.py 10000 20000 2>&1 | grep "^steps"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_clumps.py tests/test_gpu_impose.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
python bench.py --no-cpu --e2e-steps 20 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
for l in open("gpurun_out/${T}_bench.json"):
    if l.startswith("{"):
        j=json.loads(l); print("%.4g"%j["value"], "%.4f"%j["ms_per_step"], {k:round(v,4) for k,v in j["roofline"]["phase_ms_per_step"].items()}, "e2e %.4g"%j["e2e"]["value"], "settle_s", j["config"]["settle_wall_s"])
        for k,v in j.get("configs",{}).items(): print("  ",k,"%.4g"%v["value"],"%.4f"%v["ms_per_step"], v.get("wall_s_total"))
PY
