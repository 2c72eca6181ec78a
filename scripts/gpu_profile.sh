#!/bin/bash
# ncu evidence for profiles/: (1) full-set capture of every hand-written kernel on a fixed small workload,
# (2) launch list (gpu__time_duration) of one eager training step of bench.py
mkdir -p gpurun_out
echo "== ncu full"; timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"sample_kl|bayes_gemm|bayes_wgrad|wgrad_reduce|layernorm|bias_grad" -c 34 -f -o gpurun_out/prof_r1 python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_full.log
echo "== launch list B=128 (ncu, 1 step)"; timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --profile --steps 1 --warmup 1 --batch 256 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_launch.log | cut -c1-200
