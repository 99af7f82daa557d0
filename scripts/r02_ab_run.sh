T=${1:-r02s31}
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -25 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; tail -c 400 gpurun_out/${T}_bench_1gpu.json
