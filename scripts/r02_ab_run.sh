# A/B of the force-reduction variants (run under gpurun on one B200)
T=${1:-r02s28}
python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -30 gpurun_out/${T}_pytest.log
for m in 0 1; do
  WDEM_REDUCE=$m python bench.py --no-cpu --no-configs --e2e-steps 5 > gpurun_out/${T}_bench_red$m.json 2> gpurun_out/${T}_bench_red$m.err
  python - <<PY
import json
for l in open("gpurun_out/${T}_bench_red$m.json"):
    if l.startswith("{"):
        j=json.loads(l); print("reduce mode $m:", "%.4g"%j["value"], "%.4f"%j["ms_per_step"], {k:round(v,4) for k,v in j["roofline"]["phase_ms_per_step"].items()}, "e2e %.4g"%j["e2e"]["value"])
PY
done
