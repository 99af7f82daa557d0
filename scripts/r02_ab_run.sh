T=${1:-r02s38}
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "e2e" > gpurun_out/${T}_pytest.log 2>&1; tail -20 gpurun_out/${T}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 50 --no-configs --no-cpu --e2e-steps 30 --no-parity > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
tail -3 gpurun_out/${T}_bench_2gpu.err
python - <<PY
import json
for l in open("gpurun_out/${T}_bench_2gpu.json"):
    if l.startswith("{"):
        j=json.loads(l); print("%.4g"%j["value"], "%.4f"%j["ms_per_step"], "e2e %.4g"%j["e2e"]["value"], j["e2e"])
PY
