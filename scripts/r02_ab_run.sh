T=${1:-r02s36}
python -m pytest tests/test_gpu_clumps.py -m gpu -q -k "impositions" > gpurun_out/${T}_pytest.log 2>&1; tail -12 gpurun_out/${T}_pytest.log
