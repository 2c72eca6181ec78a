#!/bin/bash
# ncu --set full with source for a selection of kernels of scripts/profile_target.py:  gpu_prof_sel.sh <regex> <count> <out>
mkdir -p gpurun_out
timeout -k 10 800 ncu --set full --clock-control none --import-source on -k regex:"$1" -c $2 -f -o gpurun_out/$3 python scripts/profile_target.py > gpurun_out/ncu_$3.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_$3.log
