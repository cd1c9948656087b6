#!/bin/bash
# Secondary data point, NOT the headline configuration: chain-steps/s of the step kernel against the number
# of chains per GPU and the lanes per chain, 32x32x16, straight after the feasibility search.  The headline configuration (8192
# chains per GPU) fills exactly 0.99 wave at 16 lanes per chain; this shows what the kernel does when the number of chains is
# not what caps the parallelism.
#   gpurun -- 'bash scripts/chains_sweep.sh > gpurun_out/chains_sweep.log 2>&1'
for w in 16 8; do
  for n in 8192 16384 32768; do
    echo "== lanes per chain $w, chains $n"
    ABM_GROUP_WIDTH=$w python scripts/probe.py pp32-16 $n 1000 3 2>&1 | grep -E "init|rep"
  done
done
