mkdir -p gpurun_out
for v in pf2 pf3 pf4; do
  ( ABM_LIB=lib_$v.so timeout 300 python scripts/probe.py pp32-16 8192 2000 4 2>&1 | tail -5 ) > gpurun_out/ab2_$v.log 2>&1
done
for g in 32 64; do
  ( ABM_L2_FETCH=$g ABM_LIB=lib_pf3.so timeout 300 python scripts/probe.py pp32-16 8192 2000 4 2>&1 | tail -5 ) > gpurun_out/ab2_pf3_l2f$g.log 2>&1
done
tail -n 20 gpurun_out/ab2_*.log
