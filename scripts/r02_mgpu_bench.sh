# run under gpurun --gpus N: bench.py on N ranks (NCCL + CUDA-IPC peer mailboxes) with its mgpu_parity and configs blocks
N=${1:-8}; T=${2:-r02s39}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
python - <<PY
import json
for l in open("gpurun_out/${T}_bench_${N}gpu.json"):
    if l.startswith("{"):
        j=json.loads(l); print("%.4g"%j["value"], "%.4f"%j["ms_per_step"], {k:round(v,4) for k,v in j["roofline"]["phase_ms_per_step"].items()}, "e2e %.4g"%j["e2e"]["value"])
        for k,v in j.get("configs",{}).items(): print("  ",k,"%.4g"%v["value"],"%.4f"%v["ms_per_step"])
        print({k:(v["ok"], v.get("owned_per_rank")) for k,v in j.get("mgpu_parity",{}).items()})
PY
tail -4 gpurun_out/${T}_bench_${N}gpu.err
