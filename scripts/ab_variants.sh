mkdir -p gpurun_out
( timeout 300 python scripts/probe_sample.py pp32-16 8192 2000 2>&1 | tail -6 ) > gpurun_out/ab4_sample.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none --kernel-name-base mangled -k regex:stepKernelILi[03]E --csv --log-file gpurun_out/ab4_modes.csv python scripts/probe_sample.py pp32-16 8192 200 > gpurun_out/ab4_ncu.log 2>&1
tail -n 8 gpurun_out/ab4_sample.log gpurun_out/ab4_ncu.log
