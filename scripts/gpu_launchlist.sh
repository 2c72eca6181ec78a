#!/bin/bash
mkdir -p gpurun_out
echo "== launch list default config (ncu, 1 warm-up + 1 step, eager)"; timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python bench.py --profile --steps 1 --warmup 1 --batch 256 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_launch.log | cut -c1-200
