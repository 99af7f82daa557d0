T=${1:-r02s49}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
