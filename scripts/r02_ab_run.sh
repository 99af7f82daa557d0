T=${1:-r02s34}
python -m pytest tests/test_gpu_impose.py tests/test_gpu_kat.py tests/test_plugin_api.py -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -30 gpurun_out/${T}_pytest.log
