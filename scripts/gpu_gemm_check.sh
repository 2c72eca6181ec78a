#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests"; timeout -k 5 300 python -m pytest tests -m gpu -q --timeout 60 -x -k "tc_contractions or linear_bf16 or tiny_bert_bf16 or fused_wgrad or gelu" > gpurun_out/pytest_tc.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_tc.log | cut -c1-300
echo "== gemm timing"; timeout -k 10 200 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm.log 2>&1; echo "exit $?"; grep -E "TF|rc [1-9]" gpurun_out/debug_gemm.log | cut -c1-150
