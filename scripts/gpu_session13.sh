#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests with CTA pairs"; BF_GEMM_2CTA=2 timeout -k 5 300 python -m pytest tests -m gpu -q --timeout 60 -x -k "tc_contractions or linear_bf16 or tiny_bert_bf16" > gpurun_out/pytest_2cta.log 2>&1; echo "exit $?"; tail -12 gpurun_out/pytest_2cta.log | cut -c1-300
echo "== gemm timing 1cta"; BF_GEMM_2CTA=0 timeout -k 10 200 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm_1cta.log 2>&1; echo "exit $?"; grep -E "TF|rc [1-9]" gpurun_out/debug_gemm_1cta.log | head -20
echo "== gemm timing 2cta"; BF_GEMM_2CTA=1 timeout -k 10 200 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm_2cta.log 2>&1; echo "exit $?"; grep -E "TF|rc [1-9]|rel_err" gpurun_out/debug_gemm_2cta.log | head -40
