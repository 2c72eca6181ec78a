#!/bin/bash
mkdir -p gpurun_out
echo "== wgrad tests, CTA pairs forced"; BF_WGRAD_2CTA=2 timeout -k 5 300 python -m pytest tests -m gpu -q --timeout 60 -x -k "fused_wgrad or linear_bf16 or tiny_bert_bf16 or presample" > gpurun_out/pytest_wg2.log 2>&1; echo "exit $?"; tail -6 gpurun_out/pytest_wg2.log | cut -c1-300
echo "== gemm timing wgrad single"; BF_WGRAD_2CTA=0 timeout -k 10 200 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm_1cta.log 2>&1; echo "exit $?"; grep -E "TF|rc [1-9]" gpurun_out/debug_gemm_1cta.log | cut -c1-150
echo "== gemm timing wgrad pair"; BF_WGRAD_2CTA=2 timeout -k 10 200 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm_2cta.log 2>&1; echo "exit $?"; grep -E "TF|rc [1-9]" gpurun_out/debug_gemm_2cta.log | cut -c1-150
echo "== gemm timing wgrad auto"; timeout -k 10 200 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm_auto.log 2>&1; echo "exit $?"; grep -E "TF|rc [1-9]" gpurun_out/debug_gemm_auto.log | cut -c1-150
echo "== linear microbench (config 2)"; timeout -k 10 400 python scripts/gpu_linear_microbench.py > gpurun_out/linear_microbench.log 2>&1; echo "exit $?"; cut -c1-330 gpurun_out/linear_microbench.log | tail -12
