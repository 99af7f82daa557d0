T=${1:-r02s40}
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rebalance or add_and_remove or e2e" > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
