set -x
T=${1:-r02s46}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; tail -c 600 gpurun_out/${T}_bench_1gpu.json
bash scripts/gpu_profile.sh $T
