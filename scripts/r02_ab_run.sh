T=${1:-r02s48}
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "periodic_with or moving_wall" > gpurun_out/${T}_pytest.log 2>&1; tail -30 gpurun_out/${T}_pytest.log | cut -c1-400
