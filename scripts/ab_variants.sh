mkdir -p gpurun_out
for v in lib_unit.so lib_gen.so libabmcmc_b200.so lib_unit.so; do
  ( ABM_LIB=$v timeout 300 python scripts/probe.py pp32-16 8192 2000 4 2>&1 | tail -5 ) > gpurun_out/ab6_$v.log 2>&1
  echo $v; grep rep gpurun_out/ab6_$v.log
done
