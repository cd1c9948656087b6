# Development aid: one GPU session = the GPU test suite, the smoke check and a pass over every kernel mode.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 200 python scripts/mode_sweep_probe.py 2>&1 | tail -7
