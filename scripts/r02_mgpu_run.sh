# run under gpurun --gpus N: the real multi-process path (NCCL + CUDA-IPC peer mailboxes): pytest + bench with its mgpu_parity block
N=${1:-2}; T=${2:-r02s32}
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/${T}_pytest_mgpu.log 2>&1; tail -5 gpurun_out/${T}_pytest_mgpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
grep -o '"mgpu_parity".*"configs"' gpurun_out/${T}_bench_${N}gpu.json | cut -c1-1800
python - <<PY
import json
for l in open("gpurun_out/${T}_bench_${N}gpu.json"):
    if l.startswith("{"):
        j=json.loads(l); print("%.4g"%j["value"], "%.4f"%j["ms_per_step"], {k:round(v,4) for k,v in j["roofline"]["phase_ms_per_step"].items()}, "e2e %.4g"%j["e2e"]["value"])
        for k,v in j.get("configs",{}).items(): print("  ",k,"%.4g"%v["value"],"%.4f"%v["ms_per_step"])
PY
tail -5 gpurun_out/${T}_bench_${N}gpu.err
