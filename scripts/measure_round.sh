#!/bin/bash
# One GPU session producing the round's evidence (run under gpurun from the repository root):
#   smoke, both bench arms, ncu launch list of a (shortened) bench command, ncu full capture of the step kernel at the stationary
#   state (16x16x8, because ncu replays a kernel ~40x and saves/restores the state it writes), DRAM traffic at the bench
#   configuration at the stationary state.  Every ncu command runs only after the same command has exited 0 without ncu.
set -x
TAG=${1:-r2}
OUT=gpurun_out
python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_bench_reference_line.json 2> $OUT/${TAG}_bench_reference.err; tail -c 600 $OUT/${TAG}_bench_reference_line.json
python bench.py > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench.err; tail -c 2500 $OUT/${TAG}_bench_line.json
SHORT="bench.py --steps 2 --warmup 3 --no-cpu-baseline --burnin-steps 200000 --ess-samples 20000"
python $SHORT > $OUT/${TAG}_bench_short.json 2> $OUT/${TAG}_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_bench_launches.csv \
    python $SHORT > $OUT/${TAG}_ncu_bench.log 2>&1
python scripts/probe.py pp16-8 8192 100 2 --burnin=1500000 > $OUT/${TAG}_plain16.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:stepKernelILi0E -s 7 -c 1 \
    -f -o $OUT/prof_${TAG}_step python scripts/probe.py pp16-8 8192 100 2 --burnin=1500000 > $OUT/${TAG}_ncu_full.log 2>&1
python scripts/probe.py pp32-16 8192 100 2 --burnin=6000000 > $OUT/${TAG}_plain32.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none --kernel-name-base mangled -k regex:stepKernelILi0E -s 25 -c 1 --csv \
    --log-file $OUT/${TAG}_pp32-16_8192x100_dram_traffic.csv python scripts/probe.py pp32-16 8192 100 2 --burnin=6000000 > $OUT/${TAG}_ncu_traffic.log 2>&1
tail -8 $OUT/${TAG}_pp32-16_8192x100_dram_traffic.csv | cut -c1-40,250-
