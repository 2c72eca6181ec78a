#!/bin/bash
mkdir -p gpurun_out
echo "== ncu full"; timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"sample_kl|bayes_gemm|bayes_wgrad|wgrad_reduce|layernorm|bias_grad|clip_adamw|grad_sumsq" -c 46 -f -o gpurun_out/prof_r1c python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_full.log
