#!/bin/bash
# first GPU session: non-tensor-core parity first, then the tcgen05 diagnostics, then everything
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "== elementwise/fp32 parity"; timeout -k 10 600 python -m pytest tests -m gpu -q -x --timeout 300 \
   -k "not tc_ and not fused and not bf16 and not full_size" > gpurun_out/pytest_a.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_a.log
echo "== gemm debug"; timeout -k 10 300 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm.log 2>&1; echo "exit $?"; tail -60 gpurun_out/debug_gemm.log
echo "== tc tests"; timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 300 \
   -k "tc_ or fused or bf16 or full_size" > gpurun_out/pytest_b.log 2>&1; echo "exit $?"; tail -40 gpurun_out/pytest_b.log
echo "== microbench"; timeout -k 10 300 python scripts/gpu_microbench.py > gpurun_out/microbench.log 2>&1; echo "exit $?"; tail -40 gpurun_out/microbench.log
echo "== smoke"; timeout -k 10 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -5 gpurun_out/smoke.log
