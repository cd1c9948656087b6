#!/bin/bash
# Development aid: A/B throughput of library variants in one GPU session.
#   scripts/ab_libs.sh <tag> <workload> <chains> <mh-steps> <reps> lib[:ENV=VAL[,ENV=VAL]] ...
# Variants are built next to the product library (python -m agentbasedmcmc_b200.build -DX=Y --out=libabm_x.so) and selected
# with ABM_LIB (agentbasedmcmc_b200/_capi.py); "lib:ENV=VAL" sets measurement switches of the library for that run.
tag=$1; wl=$2; chains=$3; mh=$4; reps=$5; shift 5
mkdir -p gpurun_out
for spec in "$@"; do
  lib=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  name=$(echo "$spec" | tr ':=,/' '____')
  echo "== $spec" | tee -a gpurun_out/${tag}_ab.log
  env ABM_LIB=$lib $envs python scripts/probe.py $wl $chains $mh $reps 2>&1 | tail -$((reps + 1)) | tee -a gpurun_out/${tag}_ab.log
done
