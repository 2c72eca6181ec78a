#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu -k gelu"; timeout -k 5 300 python -m pytest tests -m gpu -q --timeout 150 -k "gelu" > gpurun_out/pytest_sel.log 2>&1; echo "exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_sel.log | head -10 | cut -c1-300
echo "== ncu full (final kernel set)"; timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"sample_kl|bayes_gemm|bayes_wgrad|wgrad_reduce|layernorm|bias_grad|resln|clip_adamw|grad_sumsq" -c 60 -f -o gpurun_out/prof_r1e python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_full.log
