#!/bin/bash
# launch list (gpu__time_duration) of ONE measured eager training step of bench.py's default workload; the warm-up step
# and the set-up kernels are outside the profiled range (cudaProfilerStart/Stop in --profile mode)
mkdir -p gpurun_out
OUT=${1:-launches_r1e}
echo "== launch list default config (ncu, 1 measured eager step)"; timeout -k 10 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/$OUT.csv python bench.py --profile --steps 1 --warmup 1 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_launch.log | cut -c1-200
