mkdir -p gpurun_out
for v in def1 def0; do
  echo "== $v"
  ABM_LIB=lib_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sample_statistics or validation_experiment or posterior" 2>&1 | tail -2
  ABM_LIB=lib_$v.so timeout 400 python scripts/probe_e2e.py pp32-16 8192 2000 6 2>&1 | grep -v "blocking\|host reduction"
done
