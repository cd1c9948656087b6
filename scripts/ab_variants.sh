mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:stepKernelILi3E -s 1 -c 1 \
    -f -o gpurun_out/prof_r1d_sample python scripts/probe_sample_only.py pp16-8 8192 100 > gpurun_out/r1d_ncu_sample.log 2>&1
tail -2 gpurun_out/r1d_ncu_sample.log
